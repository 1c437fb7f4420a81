#!/bin/bash
# lazy surface coordinates in k_shade: parity tests, then A/B bench on three workloads
timeout 900 python -m pytest tests/test_gpu_shade_parity.py tests/test_gpu_render_parity.py tests/test_gpu_trace_parity.py -x -q -m gpu 2>&1 | tail -3
for w in final random cornell; do echo "== $w"; bash tools/bench_variants.sh --workload $w; done
