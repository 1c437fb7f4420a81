"""ctypes access to the oracle (oracle/_ref/liboracle_port.so = CPU restatement; oracle/_ref/ref_tool = the
unmodified reference). TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PORT_LIB = ROOT / "oracle" / "_ref" / "liboracle_port.so"
REF_TOOL = ROOT / "oracle" / "_ref" / "ref_tool"

HIT_DTYPE = np.dtype([("t", "<f8"), ("p", "<f8", 3), ("normal", "<f8", 3), ("u", "<f8"), ("v", "<f8"),
                      ("hit", "<i4"), ("prim_id", "<i4"), ("front_face", "<i4"), ("material_type", "<i4")])

SHADE_DTYPE = np.dtype([("emitted", "<f8", 3), ("attenuation", "<f8", 3), ("origin", "<f8", 3), ("direction", "<f8", 3),
                        ("time", "<f8"), ("scattered", "<i4"), ("material_type", "<i4"), ("front_face", "<i4"), ("prim_id", "<i4")])

_lib = None


def load():
    global _lib
    if _lib is None:
        if not PORT_LIB.exists():
            raise ImportError(f"{PORT_LIB} missing: run `make -C oracle port`")
        lib = C.CDLL(str(PORT_LIB))
        lib.oracle_scene_build.argtypes = [C.c_int, C.c_char_p]
        lib.oracle_scene_build.restype = C.c_void_p
        lib.oracle_scene_build_opt.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int]
        lib.oracle_scene_build_opt.restype = C.c_void_p
        lib.oracle_scene_free.argtypes = [C.c_void_p]
        lib.oracle_scene_n_media.argtypes = [C.c_void_p]
        lib.oracle_scene_n_leaves.argtypes = [C.c_void_p]
        lib.oracle_rng_next.restype = C.c_double
        lib.oracle_rng_burn.argtypes = [C.c_long]
        lib.oracle_rng_reseed.argtypes = [C.c_uint]
        lib.oracle_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_camera_rays.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_uint64)]
        lib.oracle_render.restype = C.c_double
        _lib = lib
    return _lib


class OracleScene:
    """A scene built by the host mirror and intersected / shaded on the CPU with the reference's algorithm."""

    def __init__(self, alias: int, mesh_as_soup: bool = False, heightfield_side: int = 0):
        self.lib = load()
        self.h = self.lib.oracle_scene_build_opt(int(alias), str(ROOT / "assets").encode(), int(mesh_as_soup), int(heightfield_side))
        if not self.h:
            raise RuntimeError("oracle scene build failed")
        self.n_media = self.lib.oracle_scene_n_media(self.h)
        self.n_leaves = self.lib.oracle_scene_n_leaves(self.h)

    def trace(self, rays, t_min=0.001, t_max=float("inf"), xi=None, want_xi=False):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = len(rays)
        out = np.zeros(n, dtype=HIT_DTYPE)
        xi_c = np.ascontiguousarray(xi, dtype=np.float64) if xi is not None and self.n_media else None
        used = np.full((n, self.n_media), np.nan) if want_xi else None
        rc = self.lib.oracle_trace(self.h, rays.ctypes.data, n, t_min, t_max, xi_c.ctypes.data if xi_c is not None else None,
                                   used.ctypes.data if used is not None and self.n_media else None, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_trace failed")
        return (out, used) if want_xi else out

    def render(self, w, h, spp, max_depth=50, rows=None, burn=0, reseed=None):
        """single-mode render (engine.h:70-94); returns (rgb8, linear means, elapsed ms, rays). `burn` skips draws of
        the reference stream (what SURVEY 8c does); `reseed` starts an independent MT19937 stream after the scene build."""
        if reseed is not None:
            self.lib.oracle_rng_reseed(int(reseed))
        if burn:
            self.lib.oracle_rng_burn(burn)
        r0, r1 = rows if rows else (0, h)
        rgb = np.zeros((h, w, 3), dtype=np.uint8)
        lin = np.zeros((h, w, 3), dtype=np.float64)
        rays = C.c_uint64(0)
        ms = self.lib.oracle_render(self.h, w, h, spp, max_depth, r0, r1, rgb.ctypes.data, lin.ctypes.data, C.byref(rays))
        if ms < 0:
            raise RuntimeError("oracle_render failed")
        return rgb, lin, ms, rays.value

    def shade(self, rays, xi4, medium_xi=None):
        """closest hit (reference algorithm) + emitted / scatter of that hit with the explicit random numbers xi4 (n x 4)
        through the device's closed-form samplers. Returns (hits, shade records)."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 7)
        n = len(rays)
        xi4 = np.ascontiguousarray(xi4, dtype=np.float64).reshape(n, 4)
        mx = np.ascontiguousarray(medium_xi, dtype=np.float64) if medium_xi is not None and self.n_media else None
        hits = np.zeros(n, dtype=HIT_DTYPE)
        out = np.zeros(n, dtype=SHADE_DTYPE)
        rc = self.lib.oracle_shade(self.h, rays.ctypes.data, n, mx.ctypes.data if mx is not None else None, xi4.ctypes.data,
                                   hits.ctypes.data, out.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_shade failed")
        return hits, out

    def logged_rays(self, w, h, spp, cap=2_000_000, reseed=None):
        """every closest-hit query (primary rays and all bounces) of a w x h x spp render: (n, 7) rays + bounce index"""
        self.lib.oracle_ray_log_begin.argtypes = [C.c_void_p, C.c_uint64]
        self.lib.oracle_ray_log_end.restype = C.c_uint64
        buf = np.zeros((cap, 8))
        self.lib.oracle_ray_log_begin(buf.ctypes.data, cap)
        try:
            self.render(w, h, spp, reseed=reseed)
        finally:
            n = self.lib.oracle_ray_log_end()
        return buf[:n, :7].copy(), (50 - buf[:n, 7]).astype(int)

    def close(self):
        if self.h:
            self.lib.oracle_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_rays(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, time0, time1, st, xi3):
    """camera::camera + camera::get_ray of the oracle for explicit (s, t) and random numbers (n x 3)"""
    lib = load()
    st = np.ascontiguousarray(st, dtype=np.float64).reshape(-1, 2)
    n = len(st)
    xi3 = np.ascontiguousarray(xi3, dtype=np.float64).reshape(n, 3)
    cam9 = np.array(list(lookfrom) + list(lookat) + list(vup), dtype=np.float64)
    out = np.zeros((n, 7))
    if lib.oracle_camera_rays(cam9.ctypes.data, vfov, aspect, aperture, focus_dist, time0, time1, st.ctypes.data, xi3.ctypes.data,
                              n, out.ctypes.data) != 0:
        raise RuntimeError("oracle_camera_rays failed")
    return out
