#!/bin/bash
# GPU session r2f: same-address atomic probe, batched-append variants on the final scene, the -m gpu suite
mkdir -p gpurun_out
tools/bin/atomic_probe > gpurun_out/r2f_atomic_probe.json 2>&1; cat gpurun_out/r2f_atomic_probe.json
tools/bench_variants.sh > gpurun_out/r2f_variants.txt 2>&1
cat gpurun_out/r2f_variants.txt
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r2f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
tail -8 gpurun_out/r2f_tests.log
