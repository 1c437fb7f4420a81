#!/bin/bash
# GPU session r2t (1 GPU): the -m gpu suite, smoke, the default bench line and the reference arm with the batched queue appends,
# then a k_finish capture of the cow scene for profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/r2t_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2t_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
python bench.py --steps 8 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2t_reference.json 2> gpurun_out/r2t_reference.err; echo "reference rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_finish -s 100 -c 2 -o gpurun_out/r2t_cow_finish -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0 --spp-per-step 64 --workload cow > gpurun_out/r2t_cow_ncu.log 2>&1; echo "ncu rc=$?"
