"""small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)"""
import sys
sys.path.insert(0, '.')
import numpy as np
import another_raytracer_b200 as A
ctx = A.Context(0, 1 << 14)
for alias, kw in ((8, {}), (6, {}), (12, dict(mesh_as_soup=True, heightfield_side=40, device_bvh=True)), (10, dict(device_bvh=True))):
    sc = A.Scene.build(alias, **kw)
    ctx.upload(sc)
    cam = sc.camera(48, 48)
    rgb, st, acc, cnt = ctx.render(cam, A.RenderOptions(48, 48, 4), want_accum=True, want_count=True)
    rgb2, st2, _, cnt2 = ctx.render(cam, A.RenderOptions(48, 48, 16, mode=1, adaptive_min_spp=4, adaptive_batch_spp=4), want_count=True)
    rgb3, st3, _, _ = ctx.render(cam, A.RenderOptions(48, 48, 4, mode=2))
    g = np.load(f"tests/golden/trace_{alias}.npz") if alias != 12 else None
    if g is not None:
        out = ctx.trace_batch(g["rays"][:512], 0.001, float("inf"), g["xi"][:512] if int(g["n_media"]) else None)
        h = g["hit"][:512].astype(bool)
        assert np.array_equal(out["prim_id"][h], g["prim"][:512][h])
    print(alias, st["samples"], st["rays"], st2["samples"], st3["samples"], int(rgb.sum()), flush=True)
ctx.close()
print("done")
