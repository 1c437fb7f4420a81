#!/bin/bash
timeout 600 python tools/e2e_probe.py 2>&1 | tail -5
