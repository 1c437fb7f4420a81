#!/bin/bash
# k_pairs: CAS checked one row later, operands parked in shared memory: parity, then A/B (16 / 14 / 12 blocks) against the blocking merge
timeout 900 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -3
for w in final cornell cow; do echo "== $w"; timeout 600 bash tools/bench_variants.sh --workload $w; done
