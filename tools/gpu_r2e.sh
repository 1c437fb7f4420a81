#!/bin/bash
# GPU session r2e (2 GPUs): the multi-device tests on real peers, bench.py under torchrun (both arms), exit codes
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_reference_host.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2e_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_reference_n2.json 2> gpurun_out/r2e_reference_n2.err
echo "reference N=2 rc=$?"; tail -c 300 gpurun_out/r2e_reference_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
echo "bench N=2 rc=$?"; tail -4 gpurun_out/r2e_bench_n2.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"], "ms/step", d["ms_per_step"], d["render_ms_per_step"], "config5", d["config5"]["value"], d["config5"]["imbalance"], d["clocks"])
PY
