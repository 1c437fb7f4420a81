#!/bin/bash
# ncu --set full of two steady-state waves of one workload -> gpurun_out/traffic_<workload>.ncu-rep (then: tools/ncu_summary.py traffic <report> <workload>)
B="--steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0 --spp-per-step 64"
for wl in "$@"; do
  python bench.py $B --workload $wl > gpurun_out/traffic_${wl}_plain.log 2>&1 &&
  timeout 500 ncu --set full --clock-control none -k regex:'k_shade|k_traverse|k_pairs|k_finish|k_bin|k_walk' -s 400 -c 4 \
      -o gpurun_out/traffic_$wl -f python bench.py $B --workload $wl > gpurun_out/traffic_${wl}_ncu.log 2>&1
  echo "$wl rc=$?"
done
