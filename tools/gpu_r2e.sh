#!/bin/bash
# GPU session r2e (2 GPUs): the multi-device context on real peers, then bench.py under torchrun (both partitions) and the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_reference_host.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2e_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
echo "bench N=2 rc=$?"; tail -c 2500 gpurun_out/r2e_bench_n2.json; tail -3 gpurun_out/r2e_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --partition tiles --no-cpu-baseline --config5-steps 0 > gpurun_out/r2e_bench_n2_tiles.json 2> gpurun_out/r2e_bench_n2_tiles.err
echo "bench N=2 tiles rc=$?"; tail -c 600 gpurun_out/r2e_bench_n2_tiles.json
python tools/bench_dual.py 256 21 2>&1 | tail -3
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import another_raytracer_b200 as A
scene = A.Scene.build(8); cam = scene.camera(800, 800)
for part in ("samples", "tiles"):
    ctx = A.Context(devices=[0, 1], partition=part); ctx.upload(scene)
    ctx.render(cam, A.RenderOptions(800, 800, 16))
    t0 = time.perf_counter(); _, st, _, _ = ctx.render(cam, A.RenderOptions(800, 800, 512, seed=3)); dt = time.perf_counter() - t0
    print(f"in-library group [0,1] {part}: {800*800*512/dt/1e6:.1f} Msamples/s wall, device ms {st['device_ms']}, reduce {st['reduce_ms']:.3f} ms")
    ctx.close()
PY
