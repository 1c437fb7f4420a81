"""Device BVH builder (csrc/device/bvh_build.cuh; SURVEY 8f rank 1: replaces the bvh_node constructor, bvh.cpp:3-42).

The closest hit must not depend on the acceleration structure: with the tree built on the GPU (Morton order + binary
radix tree + collapse to 8-wide quantised nodes) every golden query must still return the reference's leaf id, and t
must be bit-identical to what the host-built SAH tree gives (same double-precision primitive tests, different boxes).
"""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _trace_golden(ctx, alias):
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    out = ctx.trace_batch(g["rays"], 0.001, float("inf"), g["xi"] if int(g["n_media"]) else None)
    return g, out


@pytest.mark.parametrize("alias", [1, 6, 7, 8, 9, 10, 11])
@pytest.mark.parametrize("leaf", [1, 3])
def test_device_tree_returns_the_reference_hits(artb, gpu_ctx, alias, leaf):
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    g, host_tree = _trace_golden(gpu_ctx, alias)
    st = gpu_ctx.build_bvh_device(leaf)
    n_tree = scene.info["n_world_prims"] - gpu_ctx.last_global_prims(scene)
    assert st["n_prims"] == n_tree
    if n_tree:
        assert st["n_nodes"] >= 1 and 1 <= st["depth"] <= 14
    _, dev_tree = _trace_golden(gpu_ctx, alias)
    h = g["hit"].astype(bool)
    assert np.array_equal(dev_tree["hit"].astype(bool), h)
    assert np.array_equal(dev_tree["prim_id"][h], g["prim"][h]), "leaf ids must be the reference's with any tree"
    assert np.array_equal(dev_tree["t"], host_tree["t"]) and np.array_equal(dev_tree["normal"], host_tree["normal"])


@pytest.mark.parametrize("alias", [8, 10])
def test_scene_without_host_tree_is_built_on_upload(artb, gpu_ctx, alias):
    scene = artb.Scene.build(alias, device_bvh=True)
    assert scene.info["n_nodes"] == 0
    gpu_ctx.upload(scene)
    st = gpu_ctx.last_bvh_build
    assert st["n_nodes"] >= 1 and st["build_ms"] > 0
    g, out = _trace_golden(gpu_ctx, alias)
    h = g["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), h)
    assert np.array_equal(out["prim_id"][h], g["prim"][h])
    assert np.array_equal(out["t"][h & (g["mat"] != 204)], g["t"][h & (g["mat"] != 204)])


def test_heightfield_host_and_device_trees_agree(artb, gpu_ctx):
    """79 k triangles (a 200 x 200 grid of the config-5 height field): random rays from above and from the camera."""
    host = artb.Scene.build(12, mesh_as_soup=True, heightfield_side=200)
    dev = artb.Scene.build(12, mesh_as_soup=True, heightfield_side=200, device_bvh=True)
    rng = np.random.default_rng(7)
    n = 20000
    rays = np.zeros((n, 7))
    rays[:, 0] = rng.uniform(-12, 12, n); rays[:, 1] = rng.uniform(4, 12, n); rays[:, 2] = rng.uniform(-12, 12, n)
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1]) - 0.05
    rays[:, 3:6] = d
    rays[:, 6] = rng.uniform(0, 1, n)
    gpu_ctx.upload(host)
    a = gpu_ctx.trace_batch(rays)
    gpu_ctx.upload(dev)
    st = gpu_ctx.last_bvh_build
    assert st["n_prims"] >= 79000 and st["depth"] <= 14
    b = gpu_ctx.trace_batch(rays)
    assert a["hit"].sum() > n // 4        # triangles are one-sided (triangle.h:60-76), many rays leave through the back
    assert np.array_equal(a["hit"], b["hit"]) and np.array_equal(a["prim_id"], b["prim_id"]) and np.array_equal(a["t"], b["t"])


def test_render_with_device_tree_matches_host_tree(artb):
    cam = None
    out = {}
    for dev in (False, True):
        scene = artb.Scene.build(8, device_bvh=dev)
        cam = scene.camera(64, 64)
        ctx = artb.Context(0, 1 << 16)
        ctx.upload(scene)
        _, st, acc, _ = ctx.render(cam, artb.RenderOptions(64, 64, 16), want_accum=True)
        out[dev] = (acc, st)
        ctx.close()
    assert out[True][1]["rays"] == out[False][1]["rays"]
    assert np.allclose(out[True][0], out[False][0], rtol=1e-4, atol=1e-4)
