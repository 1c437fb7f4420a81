#!/bin/bash
# usage: tools/bench_variants.sh [bench args...]   — runs bench.py once per another_raytracer_b200/libartb*.so (ARTB_LIB)
# and prints samples/s, extend ms per launch, extend share, image checksum. Experiment helper for gpurun calls.
for lib in another_raytracer_b200/libartb*.so; do
  printf "%-45s " "$lib"
  ARTB_LIB=$PWD/$lib python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps ${C5:-0} "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
    print(f\"{d['value']:8.1f} Msamples/s  step {d['ms_per_step']:7.2f} ms  extend {r.get('avg_launch_ms',0):.4f} ms/launch  share {r.get('extend_share_of_step',0):.3f}  checksum {d['image_checksum']}  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} ovf {r.get('overflow_rays',0)} mrays/s {d.get('mrays_per_sec',0):.0f} c5 {(d.get('config5') or {}).get('value',0):.1f}\")
except Exception as e:
    print('FAILED', e)
"
done
