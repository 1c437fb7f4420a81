import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import another_raytracer_b200 as A
scene = A.Scene.build(8)
ctx = A.Context(0, 1 << 21)
ctx.upload(scene)
cam = scene.camera(800, 800)
acc = torch.zeros((800, 800, 3), dtype=torch.float32, device='cuda')
for k in range(3):
    t0 = time.perf_counter(); st = ctx.render_device(cam, A.RenderOptions(800, 800, 64, sample_begin=k * 64, profile=True), acc.data_ptr()); t1 = time.perf_counter()
    print('render_device wall %.1f ms dev %.1f ms waves %d' % ((t1 - t0) * 1e3, st['elapsed_ms'], st['waves']))
for k in range(3):
    t0 = time.perf_counter(); ctx.upload(scene); t1 = time.perf_counter()
    rgb, st, a, _ = ctx.render(cam, A.RenderOptions(800, 800, 64, sample_begin=(1000 + k) * 64), want_accum=True); t2 = time.perf_counter()
    print('upload %.1f ms  render wall %.1f ms dev %.1f ms waves %d' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st['elapsed_ms'], st['waves']))
