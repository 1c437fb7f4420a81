"""Host-side mirror vs the reference, on CPU: RNG stream, scene graphs, flattener invariants.

Ground truth: tests/golden/scene_info.json, written by tools/make_golden.py from the UNMODIFIED reference."""
import ctypes as C
import json

import numpy as np
import pytest

from conftest import GOLDEN

INFO = json.loads((GOLDEN / "scene_info.json").read_text())
import hashlib


def test_rng_first_values(artb):
    """tracer_utils.h:27-31: default-seeded mt19937 + uniform_real_distribution."""
    lib = artb.load()
    lib.artb_host_rng_reset(5489)
    got = [lib.artb_host_rng_next() for _ in range(8)]
    assert got == INFO["rng_first"]
    assert got[:4] == [0.1354770042967805, 0.8350085899945795, 0.96886777112423139, 0.2210340429827049]


def test_rng_matches_libstdcxx_over_a_long_run(artb):
    """the restated generate_canonical mapping against numpy's MT19937 raw words (same generator, same seed)"""
    bitgen = np.random.MT19937()
    bitgen._legacy_seeding(5489)
    words = bitgen.random_raw(20000).astype(np.float64)
    expect = (words[0::2] + words[1::2] * 4294967296.0) / 18446744073709551616.0
    lib = artb.load()
    lib.artb_host_rng_reset(5489)
    got = np.array([lib.artb_host_rng_next() for _ in range(10000)])
    assert np.array_equal(got, np.minimum(expect, np.nextafter(1.0, 0.0)))


@pytest.mark.parametrize("alias", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14])
def test_scene_graph_identical_to_reference(artb, alias):
    """Same objects, same order, same doubles, same BVH shapes, same Perlin tables and texels as the reference, and
    the global RNG left at the same position."""
    ref = INFO["scenes"][str(alias)]
    s = artb.Scene.build(alias)
    assert s.info["rng_next"] == ref["rng_next"]
    assert s.info["top_level_objects"] == ref["top_level"]
    stream = s.dump_graph()
    assert len(stream) == ref["stream_len"]
    assert hashlib.sha256(stream.tobytes()).hexdigest() == ref["stream_sha256"]
    assert s.info["lookfrom"] == ref["lookfrom"] and s.info["lookat"] == ref["lookat"]
    assert s.info["vfov"] == ref["vfov"] and s.info["aperture"] == ref["aperture"] and s.info["background"] == ref["background"]
    counts = ref["counts"]
    leaves = sum(counts.get(k, 0) for k in ("sphere", "moving_sphere", "xy_rect", "xz_rect", "yz_rect", "triangle", "constant_medium"))
    assert s.info["n_leaves_total"] == leaves
    assert s.info["n_media"] == counts.get("constant_medium", 0)


def test_known_structure_numbers(artb):
    """SURVEY 8c item 3."""
    r = INFO["scenes"]
    assert r["1"]["counts"] == {"bvh_node": 1023, "moving_sphere": 389, "sphere": 485} and r["1"]["bvh_single_child"] == 150
    assert r["6"]["counts"]["xy_rect"] == 5 and r["6"]["counts"]["xz_rect"] == 7 and r["6"]["counts"]["yz_rect"] == 6
    assert r["8"]["bvh_nodes"] == 1534 and r["8"]["top_level"] == 11
    assert r["9"]["counts"]["triangle"] == 10200 and r["10"]["counts"]["triangle"] == 5804 and r["11"]["counts"]["triangle"] == 394


from flat_lib import NODE, PRIM      # noqa: E402  (record layouts of include/artb_flat.h)


def node_children(n):
    """slots of an 8-wide node -> list of (slot, 'node', index) / (slot, 'leaf', first prim, count)"""
    out, inner = [], 0
    for sl in range(8):
        m = int(n["meta"][sl])
        if m == 0:
            assert not (int(n["imask"]) >> sl) & 1
            continue
        if (m & 0x18) == 0x18:
            assert m == (1 << 5) | (24 + sl) and (int(n["imask"]) >> sl) & 1
            out.append((sl, "node", int(n["child_base"]) + inner))
            inner += 1
        else:
            count = {1: 1, 3: 2, 7: 3}[m >> 5]
            out.append((sl, "leaf", int(n["prim_base"]) + (m & 31), count))
    return out


def child_box(n, sl):
    cell = np.ldexp(1.0, n["e"].astype(int) - 127)
    lo = n["p"].astype(np.float64) + n["qlo"][:, sl].astype(np.float64) * cell
    hi = n["p"].astype(np.float64) + n["qhi"][:, sl].astype(np.float64) * cell
    return lo, hi


def _flat_arrays(scene):
    v = scene.flat_view()
    prims = np.ctypeslib.as_array(C.cast(v.prims, C.POINTER(C.c_uint8)), (v.n_prims * PRIM.itemsize,)).view(PRIM)
    nodes = (np.ctypeslib.as_array(C.cast(v.nodes, C.POINTER(C.c_uint8)), (v.n_nodes * NODE.itemsize,)).view(NODE)
             if v.n_nodes else np.zeros(0, NODE))
    return v, prims.copy(), nodes.copy()


@pytest.mark.parametrize("alias", [1, 6, 7, 8, 11])
def test_flattened_bvh_invariants(artb, alias):
    s = artb.Scene.build(alias)
    v, prims, nodes = _flat_arrays(s)
    assert PRIM.itemsize == 96 and NODE.itemsize == 80
    n_world, n_global = v.n_world_prims, v.n_global_prims
    # every world primitive is either global or referenced by exactly one leaf; every node but the root has one parent
    seen = np.zeros(n_world, dtype=int)
    seen[:n_global] += 1
    visited = np.zeros(len(nodes), dtype=int)
    depth = 0
    if n_world > n_global:
        stack = [(0, 1)]
        while stack:
            idx, d = stack.pop()
            visited[idx] += 1
            depth = max(depth, d)
            for ch in node_children(nodes[idx]):
                if ch[1] == "node":
                    stack.append((ch[2], d + 1))
                else:
                    assert ch[2] >= n_global and ch[3] <= 3
                    seen[ch[2]:ch[2] + ch[3]] += 1
    assert np.all(seen == 1) and np.all(visited == 1)
    assert depth == s.info["bvh_depth"] <= 14
    assert len(np.unique(prims["dfs_id"])) == len(prims)
    assert prims["dfs_id"].max() < s.info["n_leaves_total"]
    # globals: surfaces in the reference's order, then every constant_medium of the scene
    gt = prims["type"][:n_global]
    n_media = int((gt == 6).sum())
    assert n_media == s.info["n_media"] and np.all(gt[n_global - n_media:] == 6)
    assert np.all(np.diff(prims["dfs_id"][:n_global - n_media].astype(np.int64)) > 0)
    assert np.all(prims["type"][n_global:] != 6)
    # media point at boundary ranges behind the world primitives and carry their world box in g[4..9]
    for p in prims[:n_global][gt == 6]:
        first, count = int(p["g"][1]), int(p["g"][2])
        assert n_world <= first and first + count <= len(prims) and count > 0
        assert np.all(p["g"][4:7] < p["g"][7:10])


def test_global_primitive_lists(artb):
    """scene-spanning primitives are tested up front: the ground sphere of scene 1, the media of scene 8; tiny scenes
    (Cornell box: 18 rects) have no tree at all above a handful of primitives"""
    s1 = artb.Scene.build(1)
    v, prims, _ = _flat_arrays(s1)
    assert v.n_global_prims >= 1 and prims["g"][0][3] == 1000.0
    s8 = artb.Scene.build(8)
    v, prims, _ = _flat_arrays(s8)
    g = prims[:v.n_global_prims]
    assert int((g["type"] == 6).sum()) == 2
    s2 = artb.Scene.build(2)
    v, _, nodes = _flat_arrays(s2)
    assert v.n_global_prims == v.n_world_prims == 2 and len(nodes) == 0


@pytest.mark.parametrize("alias,leaf", [(1, 1), (1, 3), (8, 1), (11, 2)])
def test_child_boxes_bound_their_primitives(artb, alias, leaf):
    """quantised child boxes (lo planes rounded down, hi planes up) contain the double-precision boxes of everything
    below them — checked for untransformed spheres / moving spheres / rects / triangles"""
    s = artb.Scene.build(alias, max_leaf_prims=leaf)
    v, prims, nodes = _flat_arrays(s)

    def prim_box(p):
        g, t = p["g"], int(p["type"])
        if p["xform"] != 0:
            return None
        if t == 0:
            return g[:3] - g[3], g[:3] + g[3]
        if t == 1:
            c0, c1, r = g[0:3], g[3:6], g[8]
            return np.minimum(c0, c1) - r, np.maximum(c0, c1) + r
        if t == 5:
            pts = g[:9].reshape(3, 3)
            return pts.min(axis=0), pts.max(axis=0)
        a0, a1, b0, b1, k = g[:5]
        axes = {2: (0, 1, 2), 3: (0, 2, 1), 4: (1, 2, 0)}[t]
        lo = np.zeros(3); hi = np.zeros(3)
        lo[axes[0]], hi[axes[0]], lo[axes[1]], hi[axes[1]] = a0, a1, b0, b1
        lo[axes[2]], hi[axes[2]] = k - 1e-4, k + 1e-4
        return lo, hi

    def bounds(idx):
        n = nodes[idx]
        lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
        for ch in node_children(n):
            blo, bhi = child_box(n, ch[0])
            assert np.all(blo <= bhi)
            if ch[1] == "node":
                l, h = bounds(ch[2])
            else:
                boxes = [prim_box(p) for p in prims[ch[2]:ch[2] + ch[3]]]
                boxes = [b for b in boxes if b is not None]
                if not boxes:
                    continue
                l, h = np.min([b[0] for b in boxes], axis=0), np.max([b[1] for b in boxes], axis=0)
            assert np.all(blo <= l) and np.all(bhi >= h)
            lo = np.minimum(lo, l); hi = np.maximum(hi, h)
        return lo, hi

    bounds(0)


def test_zero_thickness_bvh_boxes_are_culled(artb):
    """SURVEY Q4 is a property of the reference's aabb test; here: the shipped scenes contain no such leaf"""
    for alias in (9, 10, 11):
        assert artb.Scene.build(alias).info["n_culled"] == 0


def test_heightfield_is_one_million_triangles(artb):
    s = artb.Scene.build(12)
    assert s.info["n_world_prims"] == 2 * 707 * 707 + 2
    assert s.info["bvh_depth"] <= 14


def test_soup_and_per_triangle_meshes_flatten_to_the_same_geometry(artb):
    a = artb.Scene.build(11)
    b = artb.Scene.build(11, mesh_as_soup=True)
    _, pa, _ = _flat_arrays(a)
    _, pb, _ = _flat_arrays(b)
    ta = pa[pa["type"] == 5]; tb = pb[pb["type"] == 5]
    # the reference-shaped BVH reorders the per-triangle objects (dfs order), the soup keeps file order: compare as sets
    # g[0..8] = the three vertices; g[9] is the reference-bvh-node link (tie-breaking), which only the per-triangle graph has
    ka = np.array(sorted(map(tuple, ta["g"][:, :9]))); kb = np.array(sorted(map(tuple, tb["g"][:, :9])))
    assert np.array_equal(ka, kb)


def test_bad_alias_and_missing_asset(artb, tmp_path):
    with pytest.raises(artb.ArtbError) as e:
        artb.Scene.build(99)
    assert e.value.code == -4
    with pytest.raises(artb.ArtbError):
        artb.Scene.build(10, assets_root=str(tmp_path))
    artb.Scene.build(6, assets_root=str(GOLDEN.parent.parent / "assets"))   # restore the root for later tests
