import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import another_raytracer_b200 as A
ctx = A.Context(0, 1 << 16)
for alias in range(1, 12):
    g = np.load(f"tests/golden/trace_{alias}.npz")
    sc = A.Scene.build(alias); ctx.upload(sc)
    out = ctx.trace_batch(g["rays"], 0.001, float("inf"), g["xi"] if int(g["n_media"]) else None)
    h = g["hit"].astype(bool)
    dt = np.abs(out["t"][h] - g["t"][h]) / np.abs(g["t"][h])
    dn = np.abs(out["normal"][h] - g["normal"][h]).max()
    dp = np.abs(out["p"][h] - g["p"][h]).max()
    print(f"alias {alias}: hits {h.sum()} ids equal {np.array_equal(out['prim_id'][h], g['prim'][h])} t bit-identical {(out['t'][h] == g['t'][h]).mean():.4f} max rel {dt.max():.2e} normal max abs {dn:.2e} p max abs {dp:.2e}")
