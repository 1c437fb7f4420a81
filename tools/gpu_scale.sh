#!/bin/bash
# usage (under gpurun --gpus N): bash tools/gpu_scale.sh N   — both bench arms under torchrun at N ranks, as the driver launches them
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/scale_reference_n$N.json 2> gpurun_out/scale_reference_n$N.err
echo "reference N=$N rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err
echo "bench N=$N rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_bench_n$N.err | tail -5 | cut -c1-200
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/scale_bench_n{n}.json").read().strip().splitlines()[-1])
print(f"N={n} value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["render_ms_per_step"], "config5", d["config5"]["value"], d["config5"]["imbalance"], d["clocks"])
r = json.loads(open(f"gpurun_out/scale_reference_n{n}.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"]["cores"])
PY
