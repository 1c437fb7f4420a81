"""GPU parity for the paths big meshes take (VERDICT r01 missing #1, f2): the pruned walk `k_walk` (trees beyond shared
memory), the 1 M-triangle height field scene (alias 12, BASELINE configs[4]) at sizes the oracle traces in seconds, and
the direct triangle_soup mesh path — all against the CPU oracle (`oracle/port`, pinned to the unmodified reference), ids
bit-exact and t / p / normal bit-identical (reference: src/primitives/triangle.h:22-88, src/engine/hittable_list.cpp:5-19,
src/primitives/bvh.cpp:44-52). Calls go through the C ABI (artb_trace_batch / artb_render)."""
import numpy as np
import pytest

from conftest import GOLDEN
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _assert_same_hits(out, ref, surface_only_exact=True):
    h = ref["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), h), "hit / miss decisions differ from the oracle"
    assert h.sum() > 0
    bad = np.nonzero(out["prim_id"][h] != ref["prim_id"][h])[0]
    assert bad.size == 0, f"{bad.size} of {h.sum()} leaf ids differ from the oracle, first at hit #{bad[:5]}"
    surface = h & (ref["material_type"] != 4)
    for k in ("t", "p", "normal"):
        assert np.array_equal(out[k][surface], ref[k][surface]), f"{k} of surface hits must be bit-identical to the oracle"
    medium = h & ~surface
    if medium.any():
        assert np.max(np.abs(out["t"][medium] - ref["t"][medium]) / np.abs(ref["t"][medium])) <= 1e-14
    assert np.array_equal(out["front_face"][h], ref["front_face"][h])
    assert np.array_equal(out["material_type"][h], ref["material_type"][h])
    assert np.max(np.abs(out["u"][surface] - ref["u"][surface]), initial=0) <= 1e-7
    assert np.max(np.abs(out["v"][surface] - ref["v"][surface]), initial=0) <= 1e-7


def _heightfield_rays(oracle, side):
    """all bounces of a small render + rays raining down on the field + grazing rays along it"""
    w, h = (48, 27) if side <= 64 else (32, 18)
    rays, _ = oracle.logged_rays(w, h, 1, reseed=4242)
    rng = np.random.default_rng(side)
    n = 1500 if side <= 64 else 600
    extra = np.zeros((n, 7))
    extra[:, 0] = rng.uniform(-12, 12, n); extra[:, 1] = rng.uniform(-1, 8, n); extra[:, 2] = rng.uniform(-12, 12, n)
    d = rng.normal(size=(n, 3)); d[: n // 2, 1] = -np.abs(d[: n // 2, 1]) - 0.05; d[n // 2:, 1] *= 0.05
    extra[:, 3:6] = d
    extra[:, 6] = rng.uniform(0, 1, n)
    return np.concatenate([rays, extra])


@pytest.mark.parametrize("side", [64, 200])
@pytest.mark.parametrize("device_bvh", [False, True], ids=["host_tree", "device_tree"])
def test_heightfield_walk_matches_oracle(artb, gpu_ctx, monkeypatch, side, device_bvh):
    """alias 12 (the config-5 scene: one triangle_soup + light + metal sphere): k_walk + k_finish against the oracle"""
    oracle = O.OracleScene(12, heightfield_side=side)
    rays = _heightfield_rays(oracle, side)
    ref = oracle.trace(rays)
    scene = artb.Scene.build(12, heightfield_side=side, device_bvh=device_bvh)
    assert scene.info["n_world_prims"] == 2 * (side - 1) ** 2 + 2
    gpu_ctx.upload(scene)
    monkeypatch.setenv("ARTB_TWO_PHASE", "0")            # the pruned walk, whatever the size of the tree
    walk = gpu_ctx.trace_batch(rays)
    assert gpu_ctx.trace_overflow_rays == 0
    _assert_same_hits(walk, ref)
    monkeypatch.setenv("ARTB_TWO_PHASE", "1")            # the pair kernels on the same tree: same numbers
    pairs = gpu_ctx.trace_batch(rays)
    assert np.array_equal(pairs["prim_id"], walk["prim_id"]) and np.array_equal(pairs["t"], walk["t"])


def test_heightfield_default_path_is_the_walk(artb, gpu_ctx, monkeypatch):
    """without the knob a 200 x 200 field (79 k triangles, ~0.8 MB of nodes) does not fit shared memory: trace_batch and
    the renderer both take k_walk. Checked through the result only (the walk is the one path that has no overflow pass
    and still answers when the pair buffer is shrunk to nothing)."""
    monkeypatch.delenv("ARTB_TWO_PHASE", raising=False)
    oracle = O.OracleScene(12, heightfield_side=200)
    rays = _heightfield_rays(oracle, 200)[:1200]
    scene = artb.Scene.build(12, heightfield_side=200, device_bvh=True)
    gpu_ctx.upload(scene)
    monkeypatch.setenv("ARTB_PAIR_CAPACITY", "1")
    out = gpu_ctx.trace_batch(rays)
    assert gpu_ctx.trace_overflow_rays == 0
    _assert_same_hits(out, oracle.trace(rays))


@pytest.mark.parametrize("alias", [8, 10, 13])
def test_reference_scenes_forced_down_the_walk(artb, gpu_ctx, monkeypatch, alias):
    """the golden batches of the unmodified reference through k_walk: transform chains, media, 5 804 triangles"""
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    xi = g["xi"] if int(g["n_media"]) else None
    gpu_ctx.upload(artb.Scene.build(alias))
    monkeypatch.setenv("ARTB_TWO_PHASE", "0")
    out = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), xi)
    assert gpu_ctx.trace_overflow_rays == 0
    h = g["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), h)
    assert np.array_equal(out["prim_id"][h], g["prim"][h])
    surface = h & (g["mat"] != 204)
    assert np.array_equal(out["t"][surface], g["t"][surface]) and np.array_equal(out["normal"][surface], g["normal"][surface])
    assert np.array_equal(out["p"][surface], g["p"][surface])


@pytest.mark.parametrize("alias", [9, 10, 11, 13])
@pytest.mark.parametrize("device_bvh", [False, True], ids=["host_tree", "device_tree"])
def test_direct_soup_path_matches_oracle(artb, gpu_ctx, alias, device_bvh):
    """mesh_as_soup: the .obj goes to the device as ONE triangle_soup with the per-face materials mesh::build would give
    (alias 9: map_Kd image texture with per-face texcoords; alias 10: random colours; 13: metal). Rays: the reference's
    own golden batch for the per-triangle version of the scene (same geometry, so the same rays are interesting)."""
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    oracle = O.OracleScene(alias, mesh_as_soup=True)
    xi = g["xi"] if int(g["n_media"]) else None
    ref = oracle.trace(g["rays"], xi=xi)
    scene = artb.Scene.build(alias, mesh_as_soup=True, device_bvh=device_bvh)
    gpu_ctx.upload(scene)
    out = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), xi)
    _assert_same_hits(out, ref)
    # same geometry as the per-triangle scene: every hit has the reference's t (ids are numbered in file order here)
    h = g["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), h)
    surface = h & (g["mat"] != 204)
    assert np.array_equal(out["t"][surface], g["t"][surface])


def _rmse(a, b, clip=4.0):
    return float(np.sqrt(np.mean((np.clip(a, 0, clip) - np.clip(b, 0, clip)) ** 2)))


@pytest.mark.parametrize("alias,kw,w,h,m,n", [(9, dict(mesh_as_soup=True), 96, 54, 8, 2048),
                                             (12, dict(heightfield_side=64), 96, 54, 8, 2048),
                                             (13, dict(mesh_as_soup=True), 96, 54, 8, 2048)],
                         ids=["capsule_soup_textured", "heightfield64", "cow_metal_soup"])
def test_soup_and_heightfield_images_within_noise_floor(artb, gpu_ctx, alias, kw, w, h, m, n):
    """noise-floor protocol of test_gpu_render_parity.py for the direct-mesh scenes (textured soup: map_Kd atlas)"""
    okw = dict(kw)
    a = O.OracleScene(alias, **okw).render(w, h, m, reseed=1234567)[1]
    b = O.OracleScene(alias, **okw).render(w, h, m, reseed=7654321)[1]
    scene = artb.Scene.build(alias, **kw)
    gpu_ctx.upload(scene)
    _, stats, accum, _ = gpu_ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, n, seed=7), want_accum=True)
    g = accum.astype(np.float64) / n
    floor = _rmse(a, b)
    assert floor > 0
    assert max(_rmse(g, a), _rmse(g, b)) <= 0.85 * floor
    mean_ref = 0.5 * (np.clip(a, 0, 4).mean() + np.clip(b, 0, 4).mean())
    assert abs(np.clip(g, 0, 4).mean() - mean_ref) <= 0.02 * mean_ref + 1e-4
