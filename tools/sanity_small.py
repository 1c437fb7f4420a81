#!/usr/bin/env python3
"""One small pass through every device code path (for `compute-sanitizer --tool memcheck python tools/sanity_small.py` on the
GPU box): trace / shade / camera hooks, fixed / tiled / variance-mask / adaptive-squares renders, the walk path and k_bin on
a small height field, the device BVH build, a two-context device group on one GPU, artb_upload_flat."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import another_raytracer_b200 as A  # noqa: E402
import flat_lib as F                 # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def main():
    ctx = A.Context(0, 1 << 14)
    for alias in (8, 6, 10):
        scene = A.Scene.build(alias)
        ctx.upload(scene)
        g = np.load(GOLDEN / f"trace_{alias}.npz")
        xi = g["xi"] if int(g["n_media"]) else None
        out = ctx.trace_batch(g["rays"][:600], 0.001, float("inf"), None if xi is None else xi[:600])
        h = out["hit"] == 1
        assert np.array_equal(out["prim_id"][h], g["prim"][:600][h])
        ctx.shade_batch(g["rays"][:600][h], out["prim_id"][h], out["t"][h], np.random.default_rng(1).uniform(0, 1, (int(h.sum()), 4)))
        cam = scene.camera(48, 36)
        ctx.camera_rays(cam, np.random.default_rng(2).uniform(0, 1, (64, 2)), np.random.default_rng(3).uniform(0, 1, (64, 3)))
        ctx.render(cam, A.RenderOptions(48, 36, 4), want_accum=True, want_count=True)
        ctx.render(cam, A.RenderOptions(50, 34, 4, tile_size=16, tile_parts=3, tile_part=1), want_count=True)
        ctx.render(cam, A.RenderOptions(48, 36, 24, mode=1, adaptive_min_spp=8, adaptive_batch_spp=8), want_count=True)
        ctx.render(cam, A.RenderOptions(48, 36, 4, mode=2))
        ctx.build_bvh_device(1)
        ctx.trace_batch(g["rays"][:200], 0.001, float("inf"), None if xi is None else xi[:200])
    hf = A.Scene.build(12, heightfield_side=40, device_bvh=True)
    ctx.upload(hf)
    import os
    os.environ["ARTB_TWO_PHASE"] = "0"
    rays = np.zeros((300, 7)); rays[:, 1] = 6; rays[:, 0] = np.linspace(-10, 10, 300); rays[:, 4] = -1; rays[:, 5] = 0.1
    ctx.trace_batch(rays)
    ctx.render(hf.camera(48, 36), A.RenderOptions(48, 36, 4))
    os.environ.pop("ARTB_TWO_PHASE")
    arrays = F.flat_arrays(A.Scene.build(7))
    arrays["nodes"] = np.zeros(0, F.NODE)
    ctx.upload_flat(F.flat_view_from_arrays(A, arrays, bvh_depth=0))
    ctx.render(A.Scene.build(7).camera(40, 40), A.RenderOptions(40, 40, 4))
    ctx.close()
    scene = A.Scene.build(8)
    for part in ("samples", "tiles"):
        grp = A.Context(pool_paths=1 << 13, devices=[0, 0], partition=part)
        grp.upload(scene)
        grp.render(scene.camera(48, 36), A.RenderOptions(48, 36, 5), want_accum=True, want_count=True)
        grp.render(scene.camera(48, 36), A.RenderOptions(48, 36, 16, mode=1, adaptive_min_spp=8, adaptive_batch_spp=8), want_count=True)
        grp.render(scene.camera(48, 36), A.RenderOptions(48, 36, 4, mode=2))
        grp.close()
    print("sanity_small: ok")


if __name__ == "__main__":
    main()
