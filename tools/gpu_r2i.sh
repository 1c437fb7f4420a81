#!/bin/bash
# GPU session r2i: lean k_bin on the height field, the whole -m gpu suite, one default bench line
mkdir -p gpurun_out
summ() { python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print(f\"{d['value']:8.1f} Msamples/s  {d['mrays_per_sec']:7.0f} Mrays/s step {d['ms_per_step']:7.2f} ms  trav {r.get('traverse_ms_per_launch',0):.4f} pairs {r.get('pairs_ms_per_launch',0):.4f} fin {r.get('finish_ms_per_launch',0):.4f} checksum {d['image_checksum']}\")
"; }
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
( for lib in libartb libartb_binb4 libartb_binb6; do printf "heightfield %-14s " $lib; ARTB_LIB=$PWD/another_raytracer_b200/$lib.so python bench.py --workload heightfield --spp-per-step 32 $B 2>&1 | summ; done
  for wl in final dino; do printf "%-8s " $wl; python bench.py --workload $wl $B 2>&1 | summ; done
) > gpurun_out/r2i_state.txt 2>&1
cat gpurun_out/r2i_state.txt
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
tail -8 gpurun_out/r2i_tests.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "trav/pairs/fin", r["traverse_ms_per_launch"], r["pairs_ms_per_launch"], r["finish_ms_per_launch"], "frac", r["frac"], "cpu", d["cpu_baseline"]["value"], "config5", d["config5"]["value"], d["config5"]["mrays_per_sec"])
PY
