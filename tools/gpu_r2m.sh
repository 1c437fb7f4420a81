#!/bin/bash
# GPU session r2m (1 GPU): the -m gpu suite, the default bench line, the reference arm, then the ncu captures for profiles/
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r2m_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_tests.log
tail -8 gpurun_out/r2m_tests.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_reference.json 2> gpurun_out/r2m_reference.err
echo "reference rc=$?"; tail -c 400 gpurun_out/r2m_reference.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "trav/pairs/fin", r["traverse_ms_per_launch"], r["pairs_ms_per_launch"], r["finish_ms_per_launch"], "frac", r["frac"], "cpu", d["cpu_baseline"]["value"], [x["value"] for x in d["cpu_baseline"]["rows"]], "config5", d["config5"]["value"], d["config5"]["mrays_per_sec"], "launches", d["gpu_launches"])
PY
B="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 --config5-steps 0"
python bench.py $B > gpurun_out/r2m_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2m_launches.csv python bench.py $B > gpurun_out/r2m_ncu1.log 2>&1
python bench.py $B > gpurun_out/r2m_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_traverse|k_pairs|k_finish' -s 800 -c 8 -o gpurun_out/r2m_prof python bench.py $B > gpurun_out/r2m_ncu2.log 2>&1
ls -la gpurun_out | grep r2m
