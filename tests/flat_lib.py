"""numpy views of the flattened scene records (include/artb_flat.h) and of exported scene graphs — shared test helpers."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_TOOL = ROOT / "oracle" / "_ref" / "ref_tool"

PRIM = np.dtype([("g", "<f8", 10), ("type", "<u4"), ("material", "<u4"), ("xform", "<u4"), ("dfs_id", "<u4")])
NODE = np.dtype([("p", "<f4", 3), ("e", "u1", 3), ("imask", "u1"), ("child_base", "<u4"), ("prim_base", "<u4"),
                 ("meta", "u1", 8), ("qlo", "u1", (3, 8)), ("qhi", "u1", (3, 8))])
REF_NODE = np.dtype([("lo", "<f8", 3), ("hi", "<f8", 3), ("parent", "<i4"), ("xform", "<u4"), ("depth", "<i4"), ("pad", "<u4")])
XFORM_OP = np.dtype([("a", "<f8"), ("b", "<f8"), ("c", "<f8"), ("type", "<u4"), ("pad", "<u4")])
XFORM = np.dtype([("first_op", "<u4"), ("n_ops", "<u4")])
MATERIAL = np.dtype([("albedo", "<f8", 3), ("param", "<f8"), ("type", "<u4"), ("texture", "<u4"), ("pad", "<u4", 2)])
TEXTURE = np.dtype([("p", "<f8", 9), ("type", "<u4"), ("a", "<u4"), ("b", "<u4"), ("pad", "<u4", 3)])
IMAGE = np.dtype([("offset", "<u8"), ("width", "<i4"), ("height", "<i4"), ("bytes_per_pixel", "<i4"), ("pad", "<i4")])
assert (PRIM.itemsize, NODE.itemsize, REF_NODE.itemsize, XFORM_OP.itemsize, XFORM.itemsize, MATERIAL.itemsize, TEXTURE.itemsize,
        IMAGE.itemsize) == (96, 80, 64, 32, 8, 48, 96, 24)


def _view(ptr, n, dtype):
    if not n or not ptr:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (n * dtype.itemsize,)).view(dtype).copy()


def flat_arrays(scene) -> dict:
    """every array of a scene's flat form as a numpy COPY, plus the view struct under "view" """
    v = scene.flat_view()
    return {"view": v, "prims": _view(v.prims, v.n_prims, PRIM), "nodes": _view(v.nodes, v.n_nodes, NODE),
            "ref_nodes": _view(v.ref_nodes, v.n_ref_nodes, REF_NODE), "xform_ops": _view(v.xform_ops, v.n_xform_ops, XFORM_OP),
            "xforms": _view(v.xforms, v.n_xforms, XFORM), "materials": _view(v.materials, v.n_materials, MATERIAL),
            "textures": _view(v.textures, v.n_textures, TEXTURE), "images": _view(v.images, v.n_images, IMAGE),
            "texels": _view(v.texels, v.n_texel_bytes, np.dtype("u1")),
            "perlins": _view(v.perlins, v.n_perlins * 9216, np.dtype("u1"))}


def flat_view_from_arrays(A, arrays: dict, **override):
    """an artb_flat_scene over numpy arrays (kept alive by the caller): what a foreign host passes to artb_upload_flat"""
    from another_raytracer_b200 import _abi
    src = arrays["view"]
    v = _abi.FlatSceneView()
    for name, _ in _abi.FlatSceneView._fields_:
        setattr(v, name, getattr(src, name))
    for key, count in (("prims", "n_prims"), ("nodes", "n_nodes"), ("ref_nodes", "n_ref_nodes"), ("xform_ops", "n_xform_ops"),
                       ("xforms", "n_xforms"), ("materials", "n_materials"), ("textures", "n_textures"), ("images", "n_images")):
        a = arrays[key]
        setattr(v, key, a.ctypes.data if len(a) else None)
        setattr(v, count, len(a))
    v.texels = arrays["texels"].ctypes.data if len(arrays["texels"]) else None
    v.n_texel_bytes = len(arrays["texels"])
    v.perlins = arrays["perlins"].ctypes.data if len(arrays["perlins"]) else None
    v.n_perlins = len(arrays["perlins"]) // 9216
    for k, val in override.items():
        setattr(v, k, val)
    return v


def export_reference_graph(alias: int, path) -> dict:
    """run the UNMODIFIED reference's scene_manager::build(alias) (oracle/_ref/ref_tool, built from /root/reference by
    oracle/Makefile; the binary travels to the GPU box) and serialise ITS graph with include/artb_reference_adapter.hpp"""
    import json
    env = dict(os.environ, ARTB_ASSETS=str(ROOT / "assets"))
    r = subprocess.run([str(REF_TOOL), "export", str(alias), str(path)], capture_output=True, text=True, env=env, check=True)
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
