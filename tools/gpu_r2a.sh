#!/bin/bash
# GPU session r2a: the whole -m gpu suite, one default bench line, then the kernel-variant table
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -25 gpurun_out/r2a_tests.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
tools/bench_variants.sh > gpurun_out/r2a_variants.txt 2>&1
cat gpurun_out/r2a_variants.txt
