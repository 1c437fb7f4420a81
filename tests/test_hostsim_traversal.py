"""CPU check of the device traversal SOURCE (csrc/device/rt_intersect.cuh compiled with g++, tests/hostsim): the
8-wide quantised tree the flattener emits, walked by the same code the kernels run, returns the reference's winning
leaf for every recorded query of every scene (same fixtures as the GPU gate, tests/test_gpu_trace_parity.py). This is
how tree/traversal changes are validated on machines without a GPU; the GPU gate itself goes through the C ABI."""
import numpy as np
import pytest

import another_raytracer_b200 as A
import hostsim_lib as H
from conftest import GOLDEN

ALIASES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14]


@pytest.mark.parametrize("alias", ALIASES)
@pytest.mark.parametrize("leaf", [1, 3])
@pytest.mark.parametrize("two_phase", [True, False])
def test_hostsim_matches_reference(alias, leaf, two_phase):
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    scene = A.Scene.build(alias, max_leaf_prims=leaf)
    out, stats = H.trace(scene, g["rays"], 0.001, float("inf"), g["xi"] if int(g["n_media"]) else None, two_phase=two_phase)
    h = g["hit"].astype(bool)
    assert np.array_equal(out["hit"].astype(bool), h), "hit / miss decisions differ"
    bad = np.nonzero(out["prim_id"][h] != g["prim"][h])[0]
    assert bad.size == 0, f"{bad.size} of {h.sum()} primitive ids differ, first at hit #{bad[:5]}"
    assert np.max(np.abs(out["t"][h] - g["t"][h]) / np.abs(g["t"][h])) <= 1e-9
    assert stats["nodes"] >= 0
    scene.close()


def test_window_is_inclusive():
    g = np.load(GOLDEN / "trace_6.npz")
    scene = A.Scene.build(6)
    h = g["hit"].astype(bool)
    rays = g["rays"][h][:512:37]
    full, _ = H.trace(scene, rays)
    for i in range(len(rays)):
        t = float(full["t"][i])
        a, _ = H.trace(scene, rays[i:i + 1], 0.001, t)
        assert a["hit"][0] == 1 and a["t"][0] == t and a["prim_id"][0] == full["prim_id"][i]
        b, _ = H.trace(scene, rays[i:i + 1], 0.001, float(np.nextafter(t, 0)))
        assert b["hit"][0] == 0
