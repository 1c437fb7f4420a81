"""experiment: do two independent wavefront pipelines (two contexts, two streams, two host threads) overlap on one GPU?"""
import sys, time, threading
sys.path.insert(0, '.')
import torch
import another_raytracer_b200 as A
W = H = 800
scene = A.Scene.build(8)
cam = scene.camera(W, H)
def make(pool):
    ctx = A.Context(0, pool); ctx.upload(scene)
    acc = torch.zeros((H, W, 3), dtype=torch.float32, device='cuda')
    return ctx, acc
def run(ctx, acc, spp, begin, out, k):
    st = ctx.render_device(cam, A.RenderOptions(W, H, spp, sample_begin=begin), acc.data_ptr())
    out[k] = st
for pool in (1 << 20, 1 << 21):
    c1, a1 = make(pool); c2, a2 = make(pool)
    out = {}
    run(c1, a1, 64, 0, out, 0); run(c2, a2, 64, 64, out, 1)      # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter(); run(c1, a1, 256, 0, out, 0); torch.cuda.synchronize(); t1 = time.perf_counter()
    single = W * H * 256 / (t1 - t0) / 1e6
    th = [threading.Thread(target=run, args=(c, a, 128, b, out, k)) for k, (c, a, b) in enumerate(((c1, a1, 0), (c2, a2, 128)))]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; torch.cuda.synchronize(); t1 = time.perf_counter()
    dual = W * H * 256 / (t1 - t0) / 1e6
    print(f"pool {pool}: one context 256 spp: {single:.1f} Msamples/s; two contexts x 128 spp concurrently: {dual:.1f} Msamples/s")
    c1.close(); c2.close()
