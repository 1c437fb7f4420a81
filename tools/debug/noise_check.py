import sys, numpy as np
sys.path.insert(0, '.')
import another_raytracer_b200 as A
alias, w, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = A.Context(0, 1 << 20)
s = A.Scene.build(alias); ctx.upload(s)
cl = lambda x: np.clip(x, 0, 4)
for n in (256, 1024, 4096, 32768):
    imgs = []
    for seed in (3, 4):
        rgb, st, acc, _ = ctx.render(s.camera(w, h), A.RenderOptions(w, h, n, seed=seed), want_accum=True)
        imgs.append(acc.astype(np.float64) / n)
    print(n, 'rmse(G1,G2)', np.sqrt(np.mean((cl(imgs[0]) - cl(imgs[1])) ** 2)), 'mean', imgs[0].mean(), imgs[1].mean())
# pool size dependence
for pool in (1 << 14, 1 << 17):
    c2 = A.Context(0, pool); c2.upload(s)
    rgb, st, acc, _ = c2.render(s.camera(w, h), A.RenderOptions(w, h, 4096, seed=3), want_accum=True)
    print('pool', pool, 'rmse vs seed3', np.sqrt(np.mean((cl(acc.astype(np.float64) / 4096) - cl(imgs[0])) ** 2)))
    c2.close()
