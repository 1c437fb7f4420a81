"""GPU parity, first gate (north_star): on fixed ray batches the BVH traversal must return the same hit primitive id
as the reference (bit-exact) with t / normal within 1e-5 relative.

Fixtures: tests/golden/trace_<alias>.npz — every closest-hit query the UNMODIFIED reference made while path tracing a
sparse pixel grid of each scene (primary rays and all bounces, rays inside glass, rays through media), with the winning
leaf's depth-first id and the random number each constant_medium consumed (tools/make_golden.py). Calls go through the
C ABI (artb_trace_batch) with host buffers.
"""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # north_star tolerance for t and the normal
ALIASES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14]
MAT_CODE = {200: 0, 201: 1, 202: 2, 203: 3, 204: 4}


@pytest.fixture(scope="module")
def scenes(artb):
    cache = {}

    def get(alias):
        if alias not in cache:
            cache[alias] = artb.Scene.build(alias)
        return cache[alias]
    return get


@pytest.mark.parametrize("alias", ALIASES)
def test_closest_hit_matches_reference(artb, gpu_ctx, scenes, alias):
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    scene = scenes(alias)
    assert scene.info["n_leaves_total"] == int(g["n_leaves"])
    assert scene.info["n_media"] == int(g["n_media"])
    gpu_ctx.upload(scene)
    out = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), g["xi"] if int(g["n_media"]) else None)
    assert gpu_ctx.trace_overflow_rays == 0, "golden rays must go through k_traverse / k_pairs / k_finish, not the overflow pass"

    ref_hit = g["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), ref_hit), "hit / miss decisions differ"
    h = ref_hit
    # primitive id: bit-exact
    bad = np.nonzero(out["prim_id"][h] != g["prim"][h])[0]
    assert bad.size == 0, f"{bad.size} of {h.sum()} primitive ids differ, first at hit #{bad[:5]}"
    assert np.all(out["prim_id"][~h] == -1)
    # t, p, normal: 1e-5 relative
    t_ref, t_gpu = g["t"][h], out["t"][h]
    assert np.max(np.abs(t_gpu - t_ref) / np.abs(t_ref)) <= REL_TOL
    n_ref, n_gpu = g["normal"][h], out["normal"][h]
    scale = np.maximum(np.linalg.norm(n_ref, axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(n_gpu - n_ref) / scale) <= REL_TOL
    p_ref, p_gpu = g["p"][h], out["p"][h]
    pscale = np.maximum(np.linalg.norm(p_ref, axis=1, keepdims=True), 1.0)
    assert np.max(np.abs(p_gpu - p_ref) / pscale) <= REL_TOL
    # Beyond the north-star tolerance: the kernels are compiled without FMA contraction (-fmad=false) and follow the
    # reference's operation order, so every SURFACE hit is bit-identical (t, p, normal). Only scatter events inside a
    # constant_medium go through log(): CUDA's and glibc's differ in the last place there.
    surface = g["mat"][h] != 204
    assert np.array_equal(t_gpu[surface], t_ref[surface]), "surface hit t must be bit-identical to the reference"
    assert np.array_equal(n_gpu[surface], n_ref[surface]) and np.array_equal(p_gpu[surface], p_ref[surface])
    if (~surface).any():
        assert np.max(np.abs(t_gpu[~surface] - t_ref[~surface]) / np.abs(t_ref[~surface])) <= 1e-14
    # face orientation and material class
    assert np.array_equal(out["front_face"][h], g["ff"][h])
    assert np.array_equal(out["material_type"][h], np.vectorize(MAT_CODE.get)(g["mat"][h]))
    # u,v wherever the reference defines them (moving spheres and media leave them untouched)
    has_uv = h.copy()
    has_uv[h] = np.isin(g["mat"][h], [200, 201, 202, 203]) & ~_is_moving_or_medium(scene, g, h)
    if has_uv.any():
        assert np.max(np.abs(out["u"][has_uv] - g["uv"][has_uv, 0])) <= 1e-7
        assert np.max(np.abs(out["v"][has_uv] - g["uv"][has_uv, 1])) <= 1e-7


def _is_moving_or_medium(scene, g, h):
    """leaf ids whose record has no u,v of its own: found from the graph token stream"""
    stream = scene.dump_graph()
    no_uv = set(_leaf_ids_without_uv(stream))
    return np.array([int(i) in no_uv for i in g["prim"][h]])


def _leaf_ids_without_uv(s):
    """walk the pre-order token stream (include/artb_tokens.h) and number leaves like the flattener does"""
    pos, next_id, out = 0, 0, []

    def tex():
        nonlocal pos
        t = int(s[pos]); pos += 1
        if t == 300: pos += 3
        elif t == 301: tex(); tex()
        elif t == 302: pos += 3
        elif t == 303: pos += 5
        elif t == 304: pos += 6; tex()
        elif t == 305: pos += 9
        else: raise AssertionError(f"texture token {t}")

    def mat():
        nonlocal pos
        t = int(s[pos]); pos += 1
        if t in (200, 203, 204): tex()
        elif t == 201: pos += 4
        elif t == 202: pos += 1
        else: raise AssertionError(f"material token {t}")

    def node():
        nonlocal pos, next_id
        t = int(s[pos]); pos += 1
        if t == 100:
            n = int(s[pos]); pos += 1
            for _ in range(n): node()
        elif t == 101:
            same = int(s[pos + 6]); pos += 7
            node()
            if not same: node()
        elif t == 102: pos += 3; node()
        elif t == 103: pos += 9; node()
        elif t == 104: pos += 6; node()
        elif t == 110: pos += 4; mat(); next_id += 1
        elif t == 111: pos += 9; mat(); out.append(next_id); next_id += 1
        elif t in (112, 113, 114): pos += 5; mat(); next_id += 1
        elif t == 115: pos += 9; mat(); next_id += 1
        elif t == 116:
            pos += 1; mat(); out.append(next_id); next_id += 1; node()
        else: raise AssertionError(f"hittable token {t}")

    node()
    assert pos == len(s)
    return out


def test_tmax_and_tmin_window(artb, gpu_ctx, scenes):
    """t_min/t_max are honoured inclusively like the reference's `root < t_min || t_max < root` tests."""
    g = np.load(GOLDEN / "trace_6.npz")
    gpu_ctx.upload(scenes(6))
    h = g["hit"].astype(bool)
    rays = g["rays"][h][:512:37]
    full = gpu_ctx.trace_batch(rays, 0.001, float("inf"))
    assert np.all(full["hit"] == 1)
    for i in range(len(rays)):
        t = float(full["t"][i])
        # a window ending exactly at the hit still reports it ...
        a = gpu_ctx.trace_batch(rays[i:i + 1], 0.001, t)
        assert a["hit"][0] == 1 and a["t"][0] == t and a["prim_id"][0] == full["prim_id"][i]
        # ... one ending one ulp before it does not, and one starting one ulp after it finds something farther
        b = gpu_ctx.trace_batch(rays[i:i + 1], 0.001, float(np.nextafter(t, 0)))
        assert b["hit"][0] == 0
        c = gpu_ctx.trace_batch(rays[i:i + 1], float(np.nextafter(t, np.inf)), float("inf"))
        assert c["hit"][0] == 0 or c["t"][0] > t


def test_empty_batch_and_no_scene(artb):
    ctx = artb.Context(0, 1 << 12)
    with pytest.raises(artb.ArtbError) as e:
        ctx.trace_batch(np.zeros((1, 7)))
    assert e.value.code == -5 and "Invalid input scene" in str(e.value)
    ctx.upload(artb.Scene.build(2))
    assert ctx.trace_batch(np.zeros((0, 7))).shape == (0,)
    ctx.close()


@pytest.mark.parametrize("alias", [1, 6, 8, 10])
def test_overflow_pass_gives_the_same_hits(artb, gpu_ctx, scenes, alias, monkeypatch):
    """Rays whose (ray, primitive) pairs do not fit the pair buffer are traced by the one-phase walk (k_trace_batch /
    k_extend). With the buffer shrunk to nothing every ray that meets a leaf takes that path: ids must still be the
    reference's, and t bit-identical to what the pair kernels produce."""
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    xi = g["xi"] if int(g["n_media"]) else None
    gpu_ctx.upload(scenes(alias))
    two_phase = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), xi)
    assert gpu_ctx.trace_overflow_rays == 0
    monkeypatch.setenv("ARTB_PAIR_CAPACITY", "1")
    one_phase = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), xi)
    assert gpu_ctx.trace_overflow_rays > 0
    h = g["hit"].astype(bool)
    assert np.array_equal(one_phase["hit"].astype(bool), h)
    assert np.array_equal(one_phase["prim_id"][h], g["prim"][h])
    assert np.array_equal(one_phase["prim_id"], two_phase["prim_id"])
    assert np.array_equal(one_phase["t"], two_phase["t"]), "both paths evaluate the same double-precision expressions"


def test_render_one_phase_and_two_phase_agree(artb, monkeypatch):
    """ARTB_TWO_PHASE=0 renders with the pruned one-phase walk (the path big meshes take): same samples, same hits, so
    the image can differ only by the order of the float atomics."""
    scene = artb.Scene.build(8)
    cam = scene.camera(64, 64)
    opt = artb.RenderOptions(64, 64, 16)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("ARTB_TWO_PHASE", mode)
        ctx = artb.Context(0, 1 << 16)
        ctx.upload(scene)
        _, st, acc, _ = ctx.render(cam, opt, want_accum=True)
        out[mode] = (acc, st)
        ctx.close()
    assert out["1"][1]["rays"] == out["0"][1]["rays"], "identical paths => identical ray counts"
    assert np.allclose(out["1"][0], out["0"][0], rtol=1e-4, atol=1e-4)
