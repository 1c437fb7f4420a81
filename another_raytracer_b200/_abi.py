"""ctypes binding of the C ABI declared in include/artb.h (libartb.so, built in-tree by csrc/Makefile).

This is the stub a reference maintainer would write for a Python host; the C++ equivalent is in
INTEGRATION.md. Nothing here computes anything: every call goes to the CUDA library, and loading fails
loudly if the library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libartb.so"

# every symbol include/artb.h declares (tests/test_abi_symbols.py checks the header against this list)
SYMBOLS = [
    "artb_abi_version", "artb_device_count", "artb_create", "artb_context_devices", "artb_destroy", "artb_last_error",
    "artb_scene_build", "artb_scene_from_graph", "artb_scene_free", "artb_scene_get_info", "artb_scene_dump_graph", "artb_scene_get_flat",
    "artb_upload_scene", "artb_upload_flat",
    "artb_trace_batch", "artb_trace_batch_device", "artb_trace_overflow_rays",
    "artb_shade_batch", "artb_camera_rays",
    "artb_build_bvh_device", "artb_last_bvh_build",
    "artb_render", "artb_render_device", "artb_resolve_device", "artb_stream",
    "artb_save_png", "artb_host_rng_reset", "artb_host_rng_next",
]

ARTB_OK = 0
ARTB_E_INVALID, ARTB_E_NO_DEVICE, ARTB_E_CUDA, ARTB_E_SCENE, ARTB_E_EMPTY_SCENE, ARTB_E_ADAPTIVE_SIZE = -1, -2, -3, -4, -5, -6
MODE_FIXED, MODE_VARIANCE_MASK, MODE_ADAPTIVE_SQUARES = 0, 1, 2


MAX_DEVICES = 8
PARTITION_SAMPLES, PARTITION_TILES = 0, 1


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("pool_paths", C.c_uint32), ("n_devices", C.c_int32), ("partition", C.c_int32),
                ("devices", C.c_int32 * MAX_DEVICES)]


class SceneOptions(C.Structure):
    _fields_ = [("mesh_as_soup", C.c_int32), ("heightfield_side", C.c_int32), ("max_leaf_prims", C.c_int32),
                ("device_bvh", C.c_int32), ("assets_root", C.c_char_p)]


class GraphBlob(C.Structure):
    _fields_ = [("hash", C.c_uint64), ("data", C.c_void_p), ("bytes", C.c_uint64)]


class SceneInfo(C.Structure):
    _fields_ = [("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3), ("vfov", C.c_double), ("aperture", C.c_double),
                ("background", C.c_double * 3), ("rng_next", C.c_double), ("world_lo", C.c_double * 3),
                ("world_hi", C.c_double * 3), ("rng_draws", C.c_uint64),
                ("n_prims", C.c_uint32), ("n_world_prims", C.c_uint32), ("n_nodes", C.c_uint32), ("n_media", C.c_uint32),
                ("n_materials", C.c_uint32), ("n_textures", C.c_uint32), ("n_images", C.c_uint32), ("n_xforms", C.c_uint32),
                ("n_leaves_total", C.c_uint32), ("n_culled", C.c_uint32), ("bvh_depth", C.c_uint32),
                ("top_level_objects", C.c_uint32), ("flat_bytes", C.c_uint64)]


class FlatSceneView(C.Structure):
    _fields_ = [("prims", C.c_void_p), ("n_prims", C.c_uint32), ("n_world_prims", C.c_uint32),
                ("nodes", C.c_void_p), ("n_nodes", C.c_uint32), ("reserved0", C.c_int32), ("n_global_prims", C.c_uint32),
                ("bvh_depth", C.c_int32),
                ("xform_ops", C.c_void_p), ("n_xform_ops", C.c_uint32),
                ("xforms", C.c_void_p), ("n_xforms", C.c_uint32),
                ("materials", C.c_void_p), ("n_materials", C.c_uint32),
                ("textures", C.c_void_p), ("n_textures", C.c_uint32),
                ("images", C.c_void_p), ("n_images", C.c_uint32),
                ("texels", C.c_void_p), ("n_texel_bytes", C.c_uint64),
                ("perlins", C.c_void_p), ("n_perlins", C.c_uint32),
                ("n_media", C.c_uint32), ("background", C.c_double * 3),
                ("ref_nodes", C.c_void_p), ("n_ref_nodes", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("t", C.c_double), ("p", C.c_double * 3), ("normal", C.c_double * 3), ("u", C.c_double), ("v", C.c_double),
                ("hit", C.c_int32), ("prim_id", C.c_int32), ("front_face", C.c_int32), ("material_type", C.c_int32)]


class ShadeRecord(C.Structure):
    _fields_ = [("emitted", C.c_double * 3), ("attenuation", C.c_double * 3), ("origin", C.c_double * 3),
                ("direction", C.c_double * 3), ("time", C.c_double), ("scattered", C.c_int32), ("material_type", C.c_int32),
                ("front_face", C.c_int32), ("reserved", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("lookfrom", C.c_double * 3), ("lookat", C.c_double * 3), ("vup", C.c_double * 3),
                ("vfov", C.c_double), ("aspect_ratio", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
                ("time0", C.c_double), ("time1", C.c_double)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32),
                ("seed", C.c_uint32), ("stream", C.c_uint32), ("mode", C.c_int32), ("sample_begin", C.c_int32),
                ("tile_size", C.c_int32), ("tile_parts", C.c_int32), ("tile_part", C.c_int32),
                ("adaptive_min_spp", C.c_int32), ("adaptive_batch_spp", C.c_int32), ("adaptive_threshold", C.c_double),
                ("accumulate", C.c_int32), ("profile", C.c_int32)]


class BvhStats(C.Structure):
    _fields_ = [("build_ms", C.c_double), ("n_prims", C.c_uint32), ("n_nodes", C.c_uint32), ("depth", C.c_int32),
                ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("elapsed_ms", C.c_double), ("samples", C.c_uint64), ("rays", C.c_uint64), ("waves", C.c_uint64),
                ("launches", C.c_uint64), ("extend_ms", C.c_double), ("extend_launches", C.c_uint64),
                ("pixels_active", C.c_uint64), ("pixels_converged", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("traverse_ms", C.c_double), ("pairs_ms", C.c_double), ("finish_ms", C.c_double),
                ("overflow_rays", C.c_uint64), ("n_devices", C.c_int32), ("reserved", C.c_int32), ("reduce_ms", C.c_double),
                ("device_ms", C.c_double * MAX_DEVICES), ("device_samples", C.c_uint64 * MAX_DEVICES)]


_lib = None


def load() -> C.CDLL:
    """dlopen libartb.so and declare prototypes. Raises if the library is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ARTB_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C another_raytracer_b200/csrc). This backend has no CPU fallback.")
    lib = C.CDLL(path)
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    lib.artb_abi_version.restype = C.c_int
    lib.artb_device_count.restype = C.c_int
    lib.artb_context_devices.argtypes = [vp]
    lib.artb_context_devices.restype = C.c_int
    lib.artb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.artb_destroy.argtypes = [vp]
    lib.artb_destroy.restype = None
    lib.artb_last_error.argtypes = [vp]
    lib.artb_last_error.restype = C.c_char_p
    lib.artb_scene_build.argtypes = [C.c_int32, C.POINTER(SceneOptions), C.POINTER(vp)]
    lib.artb_scene_from_graph.argtypes = [vp, C.c_int64, C.POINTER(GraphBlob), C.c_int32, C.POINTER(C.c_double * 3),
                                          C.POINTER(SceneOptions), C.POINTER(vp)]
    lib.artb_scene_from_graph.restype = C.c_int
    lib.artb_scene_free.argtypes = [vp]
    lib.artb_scene_free.restype = None
    lib.artb_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    lib.artb_scene_dump_graph.argtypes = [vp, vp, C.c_int64]
    lib.artb_scene_dump_graph.restype = C.c_int64
    lib.artb_scene_get_flat.argtypes = [vp, C.POINTER(FlatSceneView)]
    lib.artb_upload_scene.argtypes = [vp, vp]
    lib.artb_upload_flat.argtypes = [vp, C.POINTER(FlatSceneView)]
    lib.artb_trace_batch.argtypes = [vp, vp, C.c_uint64, C.c_double, C.c_double, vp, vp]
    lib.artb_trace_batch_device.argtypes = [vp, vp, C.c_uint64, C.c_double, C.c_double, vp, vp, C.POINTER(C.c_float)]
    lib.artb_shade_batch.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, vp]
    lib.artb_shade_batch.restype = C.c_int
    lib.artb_camera_rays.argtypes = [vp, C.POINTER(Camera), vp, vp, C.c_uint64, vp]
    lib.artb_camera_rays.restype = C.c_int
    lib.artb_build_bvh_device.argtypes = [vp, C.c_int32, C.POINTER(BvhStats)]
    lib.artb_build_bvh_device.restype = C.c_int
    lib.artb_last_bvh_build.argtypes = [vp, C.POINTER(BvhStats)]
    lib.artb_last_bvh_build.restype = C.c_int
    lib.artb_trace_overflow_rays.argtypes = [vp]
    lib.artb_trace_overflow_rays.restype = C.c_uint64
    lib.artb_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), vp, vp, vp, C.POINTER(Stats)]
    lib.artb_render_device.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), vp, vp, C.POINTER(Stats)]
    lib.artb_resolve_device.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp]
    lib.artb_stream.argtypes = [vp]
    lib.artb_stream.restype = vp
    lib.artb_save_png.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.artb_host_rng_reset.argtypes = [C.c_uint32]
    lib.artb_host_rng_reset.restype = None
    lib.artb_host_rng_next.restype = C.c_double
    for name in ("artb_create", "artb_scene_build", "artb_scene_get_info", "artb_scene_get_flat", "artb_upload_scene",
                 "artb_upload_flat", "artb_trace_batch", "artb_trace_batch_device", "artb_render", "artb_render_device",
                 "artb_resolve_device", "artb_save_png"):
        getattr(lib, name).restype = C.c_int
    if lib.artb_abi_version() != 2:
        raise ImportError("libartb.so ABI version mismatch")
    _lib = lib
    return lib
