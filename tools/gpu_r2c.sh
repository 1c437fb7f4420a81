#!/bin/bash
# GPU session r2c: full -m gpu suite, occupancy variants, global-primitive knob on final / cornell / random / cow
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
tail -15 gpurun_out/r2c_tests.log
tools/bench_variants.sh > gpurun_out/r2c_variants.txt 2>&1
cat gpurun_out/r2c_variants.txt
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} rays/launch {r.get('rays_per_launch',0):.0f}\")
"; }
( for wl in final cornell random cow; do for g in default 0; do
    printf "%-8s globals=%-8s " $wl $g
    if [ $g = default ]; then python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0 2>&1 | summ
    else ARTB_MAX_GLOBAL_PRIMS=$g python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0 2>&1 | summ; fi
  done; done ) > gpurun_out/r2c_globals.txt 2>&1
cat gpurun_out/r2c_globals.txt
