#!/usr/bin/env python3
"""Where the end-to-end call spends its time outside the render kernels: artb_upload_scene (final scene), and artb_render with
host buffers at 1 spp (fixed costs: pool start, resolve, device->host copies). usage: tools/e2e_probe.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import another_raytracer_b200 as A            # noqa: E402

W = H = 800
scene = A.Scene.build(8)
ctx = A.Context(0, 1 << 21)
ctx.upload(scene)
cam = scene.camera(W, H)
ctx.render(cam, A.RenderOptions(W, H, 1, seed=1), want_accum=True)
for _ in range(3):
    t0 = time.perf_counter(); ctx.upload(scene); t1 = time.perf_counter()
    img, st, acc, _ = ctx.render(cam, A.RenderOptions(W, H, 1, seed=1), want_accum=True); t2 = time.perf_counter()
    print(f"upload {1e3 * (t1 - t0):.3f} ms   render(1 spp, host buffers) wall {1e3 * (t2 - t1):.3f} ms, device {st['elapsed_ms']:.3f} ms, "
          f"h2d {st['h2d_bytes']} d2h {st['d2h_bytes']}")

# the drain regime: 1 spp = one wave of camera rays, then only ever fewer paths; every wave timed
import torch                                   # noqa: E402
accum = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
for spp in (1, 1, 4):
    st = ctx.render_device(cam, A.RenderOptions(W, H, spp, seed=1, profile=1), accum.data_ptr(), 0)
    n = max(st["extend_launches"], 1)
    print(f"spp {spp}: device {st['elapsed_ms']:.3f} ms, waves {st['waves']} (timed {st['extend_launches']}), rays {st['rays']}, per wave: "
          f"traverse {1e3 * st['traverse_ms'] / n:.1f} us, pairs {1e3 * st['pairs_ms'] / n:.1f} us, finish {1e3 * st['finish_ms'] / n:.1f} us, "
          f"stage {1e3 * st['extend_ms'] / n:.1f} us, everything {1e3 * st['elapsed_ms'] / max(st['waves'], 1):.1f} us")
