import sys, time
sys.path.insert(0, '.')
import numpy as np
import another_raytracer_b200 as A
scene = A.Scene.build(1)
ctx = A.Context(0, 1 << 21)
ctx.upload(scene)
cam = scene.camera(400, 225)
for k in range(4):
    t0 = time.perf_counter(); ctx.upload(scene); t1 = time.perf_counter()
    rgb, st, a, _ = ctx.render(cam, A.RenderOptions(400, 225, 128, sample_begin=k * 128), want_accum=True); t2 = time.perf_counter()
    print('upload %.2f ms  render wall %.2f ms  device %.2f ms  waves %d launches %d' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st['elapsed_ms'], st['waves'], st['launches']))
