#!/bin/bash
# GPU session r2h: state of the final scene, k_walk occupancy variants on the height field, ncu capture of k_walk
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} checksum {d['image_checksum']}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for wl in final cornell cow random; do printf "%-8s " $wl; python bench.py --workload $wl $B 2>&1 | summ; done
  for lib in libartb libartb_walk3 libartb_walk5 libartb_walk6; do printf "heightfield %-14s " $lib; ARTB_LIB=$PWD/another_raytracer_b200/$lib.so python bench.py --workload heightfield --spp-per-step 32 $B 2>&1 | summ; done
  printf "heightfield two-phase      "; ARTB_TWO_PHASE=1 python bench.py --workload heightfield --spp-per-step 32 $B 2>&1 | summ
) > gpurun_out/r2h_state.txt 2>&1
cat gpurun_out/r2h_state.txt
H="--workload heightfield --spp-per-step 8 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
python bench.py $H > gpurun_out/r2h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_finish|k_shade' -s 60 -c 6 -o gpurun_out/r2h_walk python bench.py $H > gpurun_out/r2h_ncu.log 2>&1
ls -la gpurun_out | tail -4
