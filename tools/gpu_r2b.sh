#!/bin/bash
# GPU session r2b: the whole -m gpu suite, variant table (feature attribution), dual-context experiment, one bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -30 gpurun_out/r2b_tests.log
tools/bench_variants.sh > gpurun_out/r2b_variants.txt 2>&1
cat gpurun_out/r2b_variants.txt
( echo "== base lib"; python tools/bench_dual.py 256 21; echo "== t512 lib"; ARTB_LIB=$PWD/another_raytracer_b200/libartb_t512.so python tools/bench_dual.py 256 21 ) > gpurun_out/r2b_dual.txt 2>&1
cat gpurun_out/r2b_dual.txt
python bench.py --steps 4 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "trav/pairs/fin", r["traverse_ms_per_launch"], r["pairs_ms_per_launch"], r["finish_ms_per_launch"], "frac", r["frac"], "cpu", d["cpu_baseline"]["value"], "config5", d["config5"]["value"])
PY
tail -3 gpurun_out/r2b_bench.err
