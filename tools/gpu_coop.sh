#!/bin/bash
# cooperative primitive step of k_walk: parity tests on the default build, then A/B on the height field and the cow forced through k_walk
timeout 900 python -m pytest tests/test_gpu_walk_parity.py tests/test_gpu_bvh_build.py tests/test_gpu_trace_parity.py -x -q -m gpu 2>&1 | tail -3
echo "== heightfield"; timeout 900 bash tools/bench_variants.sh --workload heightfield --spp-per-step 32
echo "== cow through k_walk"; ARTB_TWO_PHASE=0 timeout 600 bash tools/bench_variants.sh --workload cow --spp-per-step 64
