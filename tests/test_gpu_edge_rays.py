"""Edge-case rays, GPU against the oracle port computed live on the box's host cores (the port is pinned bit-identical
to the reference, tests/test_oracle_port.py): axis-aligned directions (1/d = inf in the reference's aabb::hit, aabb.h:16-29),
origins exactly on surfaces and on box planes, origins inside spheres and media, very short and very long direction
vectors, far-away origins. Same bar as the golden batches: ids bit-exact, surface t bit-identical."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def edge_rays(alias, lookfrom, lookat, lo, hi, rng):
    lookfrom, lookat, lo, hi = (np.asarray(v, dtype=np.float64) for v in (lookfrom, lookat, lo, hi))
    lo, hi = np.maximum(lo, -2000.0), np.minimum(hi, 2000.0)        # the radius-5000 mist sphere would dominate otherwise
    rays = []
    centre, ext = 0.5 * (lo + hi), hi - lo
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 1, 0], [0, 1, -1], [1, 0, 1]], dtype=np.float64)
    for k in range(600):                      # axis-aligned and planar directions from random points in and around the scene
        o = centre + (rng.random(3) - 0.5) * ext * 1.4
        rays.append(np.r_[o, axes[k % len(axes)] * rng.choice([1e-3, 1.0, 37.0, 1e3]), rng.random()])
    for k in range(300):                      # from the camera, scaled directions
        d = (lookat - lookfrom) + (rng.random(3) - 0.5) * np.linalg.norm(lookat - lookfrom) * 0.8
        rays.append(np.r_[lookfrom, d * rng.choice([1e-4, 1.0, 1e4]), rng.random()])
    for k in range(300):                      # far away, aimed at the scene
        o = centre + rng.normal(size=3) * ext * 20
        rays.append(np.r_[o, centre + (rng.random(3) - 0.5) * ext - o, rng.random()])
    for k in range(200):                      # origins snapped to the scene's bounding planes, direction within the plane
        o = lo + rng.random(3) * ext
        a = k % 3
        o[a] = (lo if k % 2 else hi)[a]
        d = rng.normal(size=3); d[a] = 0.0
        rays.append(np.r_[o, d, rng.random()])
    return np.array(rays)


@pytest.mark.parametrize("alias", [1, 5, 6, 7, 8, 10])
def test_edge_rays_match_the_oracle(artb, gpu_ctx, alias):
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    oracle = O.OracleScene(alias)
    i = scene.info
    rng = np.random.default_rng(100 + alias)
    rays = edge_rays(alias, i["lookfrom"], i["lookat"], i["world_lo"], i["world_hi"], rng)
    ref, xi = oracle.trace(rays, 0.001, float("inf"), want_xi=True)
    # second generation: rays leaving the hit points (origins exactly on surfaces, like every bounce of a path)
    h = ref["hit"].astype(bool)
    bounce = np.c_[ref["p"][h], rng.normal(size=(h.sum(), 3)), rays[h, 6]]
    ref2, xi2 = oracle.trace(bounce, 0.001, float("inf"), want_xi=True)
    for batch, want, used in ((rays, ref, xi), (bounce, ref2, xi2)):
        got = gpu_ctx.trace_batch(batch, 0.001, float("inf"), used if oracle.n_media else None)
        assert gpu_ctx.trace_overflow_rays <= 0.02 * len(batch)      # exact ties (coincident box faces) take the one-phase pass
        hit = want["hit"].astype(bool)
        assert np.array_equal(got["hit"].astype(bool), hit)
        assert np.array_equal(got["prim_id"], want["prim_id"])
        surface = hit & (want["material_type"] != 4)
        assert np.array_equal(got["t"][surface], want["t"][surface])
        assert np.array_equal(got["normal"][surface], want["normal"][surface])
        assert np.array_equal(got["front_face"][hit], want["front_face"][hit])
        medium = hit & (want["material_type"] == 4)
        if medium.any():
            assert np.max(np.abs(got["t"][medium] - want["t"][medium]) / np.abs(want["t"][medium])) <= 1e-14
    assert h.sum() > 200
    oracle.close()
