#!/usr/bin/env python3
"""Realistic ray sets for tuning the traversal on the host (TEST/TOOL infrastructure): renders a scene with the CPU
oracle while logging every closest-hit query, then replays the rays through tests/hostsim and reports, for several
ways of forming 32-ray warps, the SIMT utilisation of node steps = mean(node visits) / mean(per-warp max).
usage: tools/ray_sets.py [alias] [width] [height] [spp]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import another_raytracer_b200 as A      # noqa: E402
import hostsim_lib as H                 # noqa: E402
import oracle_lib as O                  # noqa: E402


def logged_rays(alias, w, h, spp, cap=4_000_000):
    sc = O.OracleScene(alias)
    lib = sc.lib
    lib.oracle_ray_log_begin.argtypes = [C.c_void_p, C.c_uint64]
    lib.oracle_ray_log_end.restype = C.c_uint64
    buf = np.zeros((cap, 8))
    lib.oracle_ray_log_begin(buf.ctypes.data, cap)
    sc.render(w, h, spp)
    n = lib.oracle_ray_log_end()
    return buf[:n].copy()


def utilisation(nv, order):
    n = len(nv) // 32 * 32
    wv = nv[order][:n].reshape(-1, 32)
    return wv.mean() / wv.max(axis=1).mean(), wv.max(axis=1).mean()


def morton3(q, bits):
    out = np.zeros(len(q), dtype=np.uint64)
    for b in range(bits):
        for a in range(3):
            out |= ((q[:, a] >> b) & 1).astype(np.uint64) << np.uint64(3 * b + a)
    return out


def main():
    alias = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 160
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    log = logged_rays(alias, w, h, spp)
    rays, depth = log[:, :7], (50 - log[:, 7]).astype(int)
    scene = A.Scene.build(alias)
    _, st = H.trace(scene, rays, per_ray=True)
    pr = st["per_ray"]
    nv = pr[:, 0].astype(np.float64)
    print(f"alias {alias}: {len(rays)} rays, node visits/ray {nv.mean():.2f}, prim tests/ray {pr[:, 1].mean():.2f}, depth hist {np.bincount(depth)[:8].tolist()}")
    rng = np.random.default_rng(0)
    o, d = rays[:, :3], rays[:, 3:6]
    lo, hi = np.percentile(o, 1, axis=0), np.percentile(o, 99, axis=0)
    octant = ((d[:, 0] < 0) * 4 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0)).astype(np.uint64)
    orders = {"path order (oracle)": np.arange(len(nv)), "random": rng.permutation(len(nv)),
              "by bounce, then pixel order": np.argsort(depth, kind="stable")}
    for bits in (2, 3, 4, 5):
        q = np.clip(((o - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
        key = (octant << np.uint64(3 * bits)) | morton3(q, bits)
        orders[f"sorted: octant + {bits}-bit/axis origin cell"] = np.argsort(key, kind="stable")
        key2 = (morton3(q, bits) << np.uint64(3)) | octant
        orders[f"sorted: {bits}-bit/axis origin cell + octant"] = np.argsort(key2, kind="stable")
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    qd = np.clip(((dn + 1) / 2 * 8).astype(np.int64), 0, 7)
    q = np.clip(((o - lo) / (hi - lo) * 8).astype(np.int64), 0, 7)
    orders["sorted: 3-bit dir + 3-bit origin"] = np.argsort((morton3(qd, 3) << np.uint64(9)) | morton3(q, 3), kind="stable")
    orders["oracle: sorted by node visits (bound)"] = np.argsort(nv, kind="stable")
    for name, order in orders.items():
        u, mx = utilisation(nv, order)
        print(f"  {name:48s} utilisation {u:.3f}   node steps per warp {mx:.2f}")


if __name__ == "__main__":
    main()
