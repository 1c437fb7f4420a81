"""The reference-side boundary on CPU (VERDICT r01 missing #5): the UNMODIFIED reference builds its scene, the
adapter a maintainer would add (include/artb_reference_adapter.hpp, here inside oracle/_ref/ref_tool) serialises the
reference's OWN graph (reference: src/engine/hittable_list.h:9-24, src/primitives/bvh.h:9-26, ...), and
artb_scene_from_graph rebuilds and flattens it. The result must be the scene the host mirror builds itself."""
import numpy as np
import pytest

import flat_lib as F

needs_ref_tool = pytest.mark.skipif(not F.REF_TOOL.exists(), reason="oracle/_ref/ref_tool not built (needs /root/reference once)")


def _resolved_materials(a):
    """per primitive: (material type, albedo, param, texture type, texture p, nested ids dropped) — independent of how the
    material / texture arrays were deduplicated"""
    m = a["materials"][a["prims"]["material"]]
    t = a["textures"][np.minimum(m["texture"], max(len(a["textures"]) - 1, 0))] if len(a["textures"]) else None
    uses_tex = np.isin(m["type"], [0, 3, 4])
    cols = [m["type"].astype(np.float64), np.where(m["type"][:, None] == 1, m["albedo"], 0.0), np.where(np.isin(m["type"], [1, 2]), m["param"], 0.0)[:, None]]
    if t is not None:
        cols += [np.where(uses_tex, t["type"], 0)[:, None].astype(np.float64), np.where(uses_tex[:, None], t["p"], 0.0)]
    return np.column_stack([c if c.ndim == 2 else c[:, None] for c in cols])


@needs_ref_tool
@pytest.mark.parametrize("alias", [1, 3, 4, 5, 6, 7, 8, 9, 11, 13])
def test_reference_graph_imports_to_the_same_flat_scene(artb, tmp_path, alias):
    from another_raytracer_b200.engine import read_graph_file
    path = tmp_path / "scene.artg"
    meta = F.export_reference_graph(alias, path)
    tokens, blobs, background = read_graph_file(path)
    assert len(tokens) == meta["tokens"] and len(blobs) == meta["blobs"]
    imported = artb.Scene.from_graph(tokens, blobs, background)
    mirror = artb.Scene.build(alias)
    # 1. the imported graph IS the reference's graph, and the host mirror builds the same one
    assert np.array_equal(imported.dump_graph(), tokens)
    assert np.array_equal(mirror.dump_graph(), tokens)
    # 2. flattening it gives the arrays the mirror's own scene flattens to
    a, b = F.flat_arrays(imported), F.flat_arrays(mirror)
    for k in ("n_world_prims", "n_global_prims", "n_media", "bvh_depth"):
        assert getattr(a["view"], k) == getattr(b["view"], k)
    for f in ("g", "type", "xform", "dfs_id"):
        assert np.array_equal(a["prims"][f], b["prims"][f]), f"prims.{f}"
    assert a["nodes"].tobytes() == b["nodes"].tobytes()
    assert a["ref_nodes"].tobytes() == b["ref_nodes"].tobytes()
    assert a["xform_ops"].tobytes() == b["xform_ops"].tobytes() and a["xforms"].tobytes() == b["xforms"].tobytes()
    assert a["texels"].tobytes() == b["texels"].tobytes() and a["perlins"].tobytes() == b["perlins"].tobytes()
    # materials are shared by CONTENT in an imported graph and by pointer in the mirror: compare what each primitive resolves to
    assert len(a["materials"]) <= len(b["materials"])
    assert np.array_equal(_resolved_materials(a), _resolved_materials(b))
    assert list(imported.info["background"]) == list(mirror.info["background"])


def _tiny_graph():
    # LIST 2: SPHERE (0,0,-1) r 0.5 lambertian(solid .5 .5 .5); XZ rect light
    return np.array([100, 2,
                     110, 0, 0, -1, 0.5, 200, 300, 0.5, 0.5, 0.5,
                     113, -1, 1, -1, 1, 2, 203, 300, 4, 4, 4], dtype=np.float64)


def test_hand_written_graph_and_malformed_streams(artb):
    s = artb.Scene.from_graph(_tiny_graph(), {}, (0.1, 0.2, 0.3))
    assert s.info["n_world_prims"] == 2 and s.info["n_materials"] == 2 and s.info["background"] == [0.1, 0.2, 0.3]
    assert np.array_equal(s.dump_graph(), _tiny_graph())
    good = _tiny_graph()
    bad = {
        "truncated": good[:-3],
        "trailing": np.append(good, 7.0),
        "unknown hittable": np.where(good == 110, 199, good),
        "unknown material": np.where(good == 203, 299, good),
        "list too long": np.concatenate([[100, 3], good[2:]]),
        "missing blob": np.array([100, 1, 110, 0, 0, 0, 1, 200, 302, 4.0, 123, 456], dtype=np.float64),
        "nan token": np.concatenate([[100, np.nan], good[2:]]),
    }
    for name, stream in bad.items():
        with pytest.raises(artb.ArtbError) as e:
            artb.Scene.from_graph(stream, {}, (0, 0, 0))
        assert e.value.code == -4, name
    # a blob whose bytes do not match the hash the stream names
    noise = np.array([100, 1, 110, 0, 0, 0, 1, 200, 302, 4.0, 123, 0], dtype=np.float64)
    with pytest.raises(artb.ArtbError):
        artb.Scene.from_graph(noise, {123: bytes(9216)}, (0, 0, 0))
