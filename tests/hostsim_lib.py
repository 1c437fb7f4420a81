"""ctypes access to tests/hostsim (TEST INFRASTRUCTURE): the device traversal source compiled with g++ and replayed
ray by ray on the host. Built on demand next to the source; nothing in the product loads it."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "hostsim" / "hostsim.cpp"
LIB = HERE / "hostsim" / "_build" / "libhostsim.so"
DEVICE_DIR = HERE.parent / "another_raytracer_b200" / "csrc" / "device"

HIT = np.dtype([("t", "<f8"), ("hit", "<i4"), ("prim_id", "<i4"), ("prim_index", "<i4"), ("pad", "<i4")])

_lib = None


def load():
    global _lib
    if _lib is None:
        deps = [SRC] + sorted(DEVICE_DIR.glob("*.cuh")) + sorted(DEVICE_DIR.glob("*.h")) + sorted((HERE.parent / "include").glob("*.h"))
        if not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps):
            LIB.parent.mkdir(exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=gnu++17", "-fPIC", "-shared", f"-I{HERE.parent / 'include'}", "-x", "c++", str(SRC), "-o", str(LIB)], check=True)
        lib = C.CDLL(str(LIB))
        lib.hostsim_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hostsim_trace.restype = C.c_int
        lib.hostsim_set_two_phase.argtypes = [C.c_int]
        _lib = lib
    return _lib


def trace(scene, rays, t_min=0.001, t_max=float("inf"), xi=None, per_ray=False, two_phase=True):
    """closest hits of `rays` over the flattened `scene` (another_raytracer_b200.Scene); returns (hits, stats);
    per_ray=True adds stats["per_ray"]: (n, 3) node visits / primitive tests / medium tests of every ray"""
    lib = load()
    lib.hostsim_set_two_phase(int(two_phase))     # True: collect + confirm (what the kernels run); False: the one-phase walk
    view = scene.flat_view()
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
    n = len(rays)
    out = np.zeros(n, dtype=HIT)
    stats = np.zeros(3, dtype=np.uint64)
    xi_c = np.ascontiguousarray(xi, dtype=np.float64) if xi is not None and view.n_media else None
    pr = np.zeros((n, 3), dtype=np.uint32) if per_ray else None
    rc = lib.hostsim_trace(C.byref(view), rays.ctypes.data, n, t_min, t_max, xi_c.ctypes.data if xi_c is not None else None,
                           out.ctypes.data, stats.ctypes.data, pr.ctypes.data if per_ray else None)
    assert rc == 0
    st = {"nodes": int(stats[0]), "prims": int(stats[1]), "media": int(stats[2])}
    if per_ray:
        st["per_ray"] = pr
    return out, st


def rect_screen(prim_type, g, rays, t_min, t_max):
    """bit 0: the exact rectangle test hits, bit 1: the float screen of k_finish calls the ray a certain miss"""
    lib = load()
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
    g = np.ascontiguousarray(g, dtype=np.float64)
    t_max = np.ascontiguousarray(np.broadcast_to(np.asarray(t_max, dtype=np.float64), (len(rays),)))
    out = np.zeros(len(rays), dtype=np.uint8)
    lib.hostsim_rect_screen.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_void_p, C.c_void_p]
    rc = lib.hostsim_rect_screen(int(prim_type), g.ctypes.data, rays.ctypes.data, len(rays), float(t_min), t_max.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out
