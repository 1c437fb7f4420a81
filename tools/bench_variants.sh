#!/bin/bash
# usage: tools/bench_variants.sh [bench args...]   — runs bench.py once per another_raytracer_b200/libartb*.so (ARTB_LIB)
# and prints samples/s, extend ms per launch, extend share, image checksum. Experiment helper for gpurun calls.
for lib in another_raytracer_b200/libartb*.so; do
  printf "%-45s " "$lib"
  ARTB_LIB=$PWD/$lib python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0 "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print(f\"{d['value']:8.1f} Msamples/s  step {d['ms_per_step']:7.2f} ms  extend {r['avg_launch_ms']:.4f} ms/launch  share {r['extend_share_of_step']:.3f}  checksum {d['image_checksum']}  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} ovf {r.get('overflow_rays',0)}\")
except Exception as e:
    print('FAILED', e)
"
done
