#!/usr/bin/env python3
"""Experiment helper (GPU box): one context against two contexts sharing ONE GPU (artb_config devices = [0, 0]: two path
pools, two streams, two host threads) on the final scene — does co-scheduling the issue-bound tree walk of one pool with
the latency-bound kernels of the other fill the SMs better? usage: tools/bench_dual.py [spp] [pool_log2]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import another_raytracer_b200 as A  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
plog = int(sys.argv[2]) if len(sys.argv) > 2 else 21
W = H = 800
scene = A.Scene.build(8)
cam = scene.camera(W, H)
for label, kw in (("one context ", dict(device=0, pool_paths=1 << plog)),
                  ("two contexts", dict(devices=[0, 0], pool_paths=1 << (plog - 1))),
                  ("two ctx full", dict(devices=[0, 0], pool_paths=1 << plog))):
    ctx = A.Context(**kw)
    ctx.upload(scene)
    ctx.render(cam, A.RenderOptions(W, H, 16, seed=1))
    best = 1e9
    for k in range(3):
        t0 = time.perf_counter()
        _, st, _, _ = ctx.render(cam, A.RenderOptions(W, H, spp, seed=1, sample_begin=k * spp))
        best = min(best, time.perf_counter() - t0)
    print(f"{label}: {W * H * spp / best / 1e6:8.1f} Msamples/s wall  (device ms {st['device_ms']}, reduce {st['reduce_ms']:.3f} ms)", flush=True)
    ctx.close()
