#!/bin/bash
# GPU session r2j: cp.async record staging in k_shade (on / off), then the -m gpu suite
mkdir -p gpurun_out
tools/bench_variants.sh > gpurun_out/r2j_variants.txt 2>&1
tools/bench_variants.sh --workload cornell >> gpurun_out/r2j_variants.txt 2>&1
tools/bench_variants.sh --workload cow >> gpurun_out/r2j_variants.txt 2>&1
cat gpurun_out/r2j_variants.txt
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
tail -6 gpurun_out/r2j_tests.log
