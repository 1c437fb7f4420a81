#!/bin/bash
# sampled per-wave events (every 8th wave): the default bench line
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["render_ms_per_step"], {k: r[k] for k in ("achieved", "frac", "avg_launch_ms", "launches", "launches_timed", "rays_per_launch", "extend_share_of_step", "traverse_ms_per_launch", "pairs_ms_per_launch", "finish_ms_per_launch")}, "config5", d["config5"]["value"])
PY
timeout 600 python -m pytest tests/test_gpu_render_parity.py -x -q -m gpu 2>&1 | tail -2
