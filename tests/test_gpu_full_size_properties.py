"""Parity at BASELINE.json's FULL image sizes through size-independent properties (the oracle needs minutes per frame
there): results must not depend on how the work is cut — pool size, sample ranges, tiles, the one- or two-phase
closest-hit path, the host- or device-built tree — because every sample is keyed by (pixel, sample index) in Philox and
the closest hit does not depend on the acceleration structure. Sums may differ by the order of the float atomics only.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def close(a, b):
    return np.allclose(a, b, rtol=2e-4, atol=2e-4)


def test_final_scene_800x800_is_independent_of_pool_size_and_sample_ranges(artb):
    scene = artb.Scene.build(8)
    cam = scene.camera(800, 800)
    big = artb.Context(0, 1 << 21); big.upload(scene)
    small = artb.Context(0, 1 << 17); small.upload(scene)
    _, st_all, all16, _ = big.render(cam, artb.RenderOptions(800, 800, 16, seed=5), want_accum=True)
    _, st_small, small16, _ = small.render(cam, artb.RenderOptions(800, 800, 16, seed=5), want_accum=True)
    _, st_a, first8, _ = big.render(cam, artb.RenderOptions(800, 800, 8, seed=5, sample_begin=0), want_accum=True)
    _, st_b, next8, _ = big.render(cam, artb.RenderOptions(800, 800, 8, seed=5, sample_begin=8), want_accum=True)
    assert st_all["samples"] == 800 * 800 * 16
    assert st_all["rays"] == st_small["rays"] == st_a["rays"] + st_b["rays"], "identical paths => identical ray counts"
    assert close(all16, small16), "the image must not depend on the pool size / wave order"
    assert close(all16, first8 + next8), "sample ranges add up (what the multi-GPU sample split relies on)"
    assert 4.0 < st_all["rays"] / st_all["samples"] < 6.0          # 4.85 rays per sample on this scene
    big.close(); small.close()


def test_final_scene_800x800_tiles_compose(artb):
    scene = artb.Scene.build(8)
    cam = scene.camera(800, 800)
    ctx = artb.Context(0, 1 << 20); ctx.upload(scene)
    _, st, whole, _ = ctx.render(cam, artb.RenderOptions(800, 800, 4, seed=9), want_accum=True)
    parts, rays = np.zeros_like(whole), 0
    for part in range(4):
        _, sp, acc, _ = ctx.render(cam, artb.RenderOptions(800, 800, 4, seed=9, tile_parts=4, tile_part=part), want_accum=True)
        assert np.count_nonzero(acc.sum(axis=2)) <= 800 * 800 // 4 + 32 * 800
        parts += acc
        rays += sp["rays"]
    assert rays == st["rays"] and close(parts, whole)
    ctx.close()


def test_cow_1920x1080_one_phase_equals_two_phase(artb, monkeypatch):
    scene = artb.Scene.build(10)
    cam = scene.camera(1920, 1080)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("ARTB_TWO_PHASE", mode)
        ctx = artb.Context(0, 1 << 20); ctx.upload(scene)
        _, st, acc, _ = ctx.render(cam, artb.RenderOptions(1920, 1080, 2, seed=4), want_accum=True)
        out[mode] = (acc, st["rays"])
        ctx.close()
    assert out["1"][1] == out["0"][1] and close(out["1"][0], out["0"][0])


def test_heightfield_3840x2160_device_tree_equals_host_tree(artb):
    out = {}
    for dev in (False, True):
        scene = artb.Scene.build(12, mesh_as_soup=True, device_bvh=dev)
        ctx = artb.Context(0, 1 << 21); ctx.upload(scene)
        _, st, acc, _ = ctx.render(scene.camera(3840, 2160), artb.RenderOptions(3840, 2160, 1, seed=2), want_accum=True)
        out[dev] = (acc, st["rays"])
        ctx.close()
    assert out[True][1] == out[False][1] and close(out[True][0], out[False][0])
