import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import another_raytracer_b200 as A
ctx = A.Context(0, 1 << 16)
for alias in (2, 6):
    g = np.load(f"tests/golden/trace_{alias}.npz")
    sc = A.Scene.build(alias); ctx.upload(sc)
    out = ctx.trace_batch(g["rays"], 0.001, float("inf"), None)
    h = g["hit"].astype(bool)
    bad = np.nonzero(h & (out["t"] != g["t"]))[0]
    print("alias", alias, "differing", len(bad), "of", h.sum())
    for i in bad[:12]:
        a, b = out["t"][i], g["t"][i]
        ulps = int(np.float64(a).view(np.int64) - np.float64(b).view(np.int64))
        print(f"  ray {i}: prim {g['prim'][i]} gpu {a!r} ref {b!r} ulps {ulps} o {g['rays'][i][:3]} d {g['rays'][i][3:6]}")
