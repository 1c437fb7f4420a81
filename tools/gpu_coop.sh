#!/bin/bash
# float screen of the global rectangles in k_finish: parity suites, then A/B on Cornell, cornell-like scenes, final
timeout 900 python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_upload_flat.py -x -q -m gpu 2>&1 | tail -3
for w in cornell final random cow; do echo "== $w"; timeout 600 bash tools/bench_variants.sh --workload $w; done
