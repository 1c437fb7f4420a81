#!/bin/bash
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} checksum {d['image_checksum']}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for wl in final cornell cow dino random; do printf "%-8s " $wl; python bench.py --workload $wl $B 2>&1 | summ; done
  printf "heightfield "; python bench.py --workload heightfield --spp-per-step 32 $B 2>&1 | summ ) > gpurun_out/r2n_state.txt 2>&1
cat gpurun_out/r2n_state.txt
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -4
