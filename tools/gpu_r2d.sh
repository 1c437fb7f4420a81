#!/bin/bash
# GPU session r2d (1 GPU): k_finish variants on the scenes without media, then the ncu captures of the final kernels
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} rays/launch {r.get('rays_per_launch',0):.0f}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for wl in cornell random cow final; do for lib in libartb libartb_light4 libartb_light5 libartb_light8; do
    printf "%-8s %-16s " $wl $lib
    ARTB_LIB=$PWD/another_raytracer_b200/$lib.so python bench.py --workload $wl $B 2>&1 | summ
  done; done ) > gpurun_out/r2d_light.txt 2>&1
cat gpurun_out/r2d_light.txt
python bench.py $B > gpurun_out/r2d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2d_launches.csv python bench.py $B > gpurun_out/r2d_ncu1.log 2>&1
python bench.py $B > gpurun_out/r2d_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse|k_pairs|k_finish' -s 800 -c 8 -o gpurun_out/r2d_prof python bench.py $B > gpurun_out/r2d_ncu2.log 2>&1
ls -la gpurun_out/ | tail -8
