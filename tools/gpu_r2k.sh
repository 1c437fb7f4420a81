#!/bin/bash
# GPU session r2k: compute-sanitizer memcheck over every device code path (small sizes)
mkdir -p gpurun_out
python tools/sanity_small.py > gpurun_out/r2k_plain.log 2>&1 && echo "plain run ok" &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanity_small.py > gpurun_out/r2k_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r2k_memcheck.log; tail -3 gpurun_out/r2k_plain.log
