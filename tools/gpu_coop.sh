#!/bin/bash
for d in 0 65536 0 65536; do echo "== ARTB_DRAIN_RAYS=$d"; ARTB_DRAIN_RAYS=$d timeout 300 python tools/e2e_probe.py 2>&1 | tail -3 | cut -c1-120; done
