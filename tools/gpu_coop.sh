#!/bin/bash
# every BASELINE configuration at its full size through the Engine mirror, final build
mkdir -p gpurun_out/full_configs
timeout 1200 python tools/full_configs.py gpurun_out/full_configs 2>&1 | tail -12
rm -f gpurun_out/full_configs/heightfield*.png
ls -la gpurun_out/full_configs
