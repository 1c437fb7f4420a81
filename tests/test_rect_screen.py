"""The FP32 screen k_finish puts in front of the exact global-rectangle tests (rect_certain_miss, rt_intersect.cuh) may
only reject rays the exact test (rect_t, the reference's aarect.cpp:3-55) rejects too. Checked on the host build of the
same source (tests/hostsim) with random rays, rays aimed at the edges and corners, axis-parallel rays, zero direction
components, huge and tiny magnitudes, and t_max equal to the exact hit parameter."""
import numpy as np
import pytest

import hostsim_lib

RECT_XY, RECT_XZ, RECT_YZ = 2, 3, 4          # include/artb_flat.h
AXES = {RECT_XY: (2, 0, 1), RECT_XZ: (1, 0, 2), RECT_YZ: (0, 1, 2)}      # plane axis, a axis, b axis

RECTS = [
    (RECT_YZ, [0, 555, 0, 555, 555]), (RECT_YZ, [0, 555, 0, 555, 0]), (RECT_XZ, [213, 343, 227, 332, 554]),
    (RECT_XZ, [0, 555, 0, 555, 0]), (RECT_XZ, [0, 555, 0, 555, 555]), (RECT_XY, [0, 555, 0, 555, 555]),      # scene_manager.cpp:121-126
    (RECT_XY, [3, 5, 1, 3, -2]), (RECT_XZ, [-1e-3, 1e-3, -1e-3, 1e-3, 1e-4]), (RECT_YZ, [-1e6, 1e6, -1e6, 1e6, 123456.789]),
]


def rays_for(prim_type, g, rng, n):
    k, a, b = AXES[prim_type]
    a0, a1, b0, b1, kk = g
    span = max(abs(a0), abs(a1), abs(b0), abs(b1), abs(kk), 1e-3)
    o = rng.uniform(-1.5 * span, 1.5 * span, (n, 3))
    target = np.zeros((n, 3))
    target[:, k] = kk
    # a third of the rays aim inside, a third exactly at an edge / corner, a third around the rectangle
    u, v = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    kind = rng.integers(0, 3, n)
    u = np.where(kind == 1, rng.integers(0, 2, n).astype(float), np.where(kind == 2, rng.uniform(-0.5, 1.5, n), u))
    v = np.where(kind == 1, rng.choice([0.0, 1.0, 0.5], n), np.where(kind == 2, rng.uniform(-0.5, 1.5, n), v))
    target[:, a] = a0 + u * (a1 - a0)
    target[:, b] = b0 + v * (b1 - b0)
    d = target - o
    d *= rng.uniform(0.25, 4.0, (n, 1))                       # any length, as the reference's scattered rays
    # special directions: parallel to the plane, axis-aligned, zero components, tiny components
    m = rng.integers(0, 12, n)
    d[m == 0, k] = 0.0
    d[m == 1, a] = 0.0
    d[m == 2, k] = rng.choice([1e-300, -1e-300, 1e-40, -1e-40, 1e-20], (m == 2).sum())
    o[m == 3, k] = kk                                          # origin on the plane
    rays = np.zeros((n, 7))
    rays[:, 0:3], rays[:, 3:6] = o, d
    return rays


@pytest.mark.parametrize("prim_type,g", RECTS)
def test_screen_never_rejects_a_hit(prim_type, g):
    rng = np.random.default_rng(prim_type * 1000 + int(abs(g[4])) % 997)
    rays = rays_for(prim_type, g, rng, 200_000)
    flags = hostsim_lib.rect_screen(prim_type, g, rays, 0.001, np.inf)
    hits, certain = (flags & 1).astype(bool), (flags & 2).astype(bool)
    assert not np.any(hits & certain)
    assert hits.sum() > 20_000                                 # the batch does exercise hits ...
    assert certain.sum() > 20_000                              # ... and the screen does reject
    # with t_max at the exact hit parameter (a tie with the best hit so far, hittable_list.cpp:11-15) and one ulp around it
    k = AXES[prim_type][0]
    with np.errstate(all="ignore"):
        t_hit = (g[4] - rays[:, k]) / rays[:, 3 + k]
    for t_max in (t_hit, np.nextafter(t_hit, np.inf), np.nextafter(t_hit, -np.inf), t_hit * (1 + 1e-7), t_hit * (1 - 1e-7)):
        t_max = np.where(np.isfinite(t_max), t_max, np.inf)
        flags = hostsim_lib.rect_screen(prim_type, g, rays, 0.001, t_max)
        assert not np.any((flags & 1).astype(bool) & (flags & 2).astype(bool))


def test_screen_rejects_most_misses_of_a_cornell_wall():
    """what the screen is for: of rays that miss a wall, nearly all are rejected without the double-precision test"""
    rng = np.random.default_rng(7)
    prim_type, g = RECTS[0]
    rays = np.zeros((100_000, 7))
    rays[:, 0:3] = rng.uniform(0, 555, (len(rays), 3))          # anywhere in the room, any direction
    rays[:, 3:6] = rng.normal(size=(len(rays), 3))
    flags = hostsim_lib.rect_screen(prim_type, g, rays, 0.001, np.inf)
    miss = (flags & 1) == 0
    assert 0.1 < 1 - miss.mean() < 0.5
    assert ((flags & 2) != 0)[miss].mean() > 0.999
