"""Edge cases of the hot path through the C ABI: ragged / tiny batches, tiny images, depth limits, oversized frames.
Reference behaviour: engine.h:58-94 (sample loop), engine.h:447-466 (`depth <= 0` returns black BEFORE the hit test, so
max_depth = 1 sees only emitted light of the first hit), hittable_list.cpp:5-19."""
import numpy as np
import pytest

from conftest import GOLDEN
import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_ragged_batches_equal_slices_of_the_whole(artb, gpu_ctx):
    """batch sizes that are not multiples of the warp / chunk size give the results of the same rays in the big batch"""
    g = np.load(GOLDEN / "trace_8.npz")
    gpu_ctx.upload(artb.Scene.build(8))
    whole = gpu_ctx.trace_batch(g["rays"], 0.001, float("inf"), g["xi"])
    for n in (1, 2, 31, 32, 33, 257, 1000):
        part = gpu_ctx.trace_batch(g["rays"][:n], 0.001, float("inf"), g["xi"][:n])
        assert part.tobytes() == whole[:n].tobytes(), n
    tail = gpu_ctx.trace_batch(g["rays"][-77:], 0.001, float("inf"), g["xi"][-77:])
    assert tail.tobytes() == whole[-77:].tobytes()


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_depth_limit_matches_oracle(artb, gpu_ctx, depth):
    """max_depth small: the depth cut returns black before the hit test (engine.h:451-452) — compared with the oracle at the
    same max_depth by the noise-floor protocol (Cornell box: depth 1 shows only the light itself)"""
    w = h = 64
    a = O.OracleScene(6).render(w, h, 64, max_depth=depth, reseed=11)[1]
    b = O.OracleScene(6).render(w, h, 64, max_depth=depth, reseed=12)[1]
    scene = artb.Scene.build(6)
    gpu_ctx.upload(scene)
    _, st, acc, _ = gpu_ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, 4096, max_depth=depth, seed=3), want_accum=True)
    g = acc.astype(np.float64) / 4096
    assert st["rays"] <= st["samples"] * depth
    rm = lambda x, y: float(np.sqrt(np.mean((np.clip(x, 0, 4) - np.clip(y, 0, 4)) ** 2)))
    floor = rm(a, b)
    if depth == 1:
        # only directly visible emitters: no Monte-Carlo noise beyond pixel coverage of the light's edge
        assert rm(g, a) < 0.08 and abs(g.mean() - a.mean()) < 0.03 * a.mean() + 1e-4
    else:
        assert max(rm(g, a), rm(g, b)) <= 0.85 * floor
        assert abs(np.clip(g, 0, 4).mean() - 0.5 * (np.clip(a, 0, 4).mean() + np.clip(b, 0, 4).mean())) <= 0.03 * np.clip(a, 0, 4).mean() + 1e-4


def test_smallest_image_one_sample_and_bad_sizes(artb, gpu_ctx):
    scene = artb.Scene.build(2)
    gpu_ctx.upload(scene)
    rgb, st, acc, cnt = gpu_ctx.render(scene.camera(2, 2), artb.RenderOptions(2, 2, 1), want_accum=True, want_count=True)
    assert st["samples"] == 4 and np.all(cnt == 1) and rgb.shape == (2, 2, 3)
    for w, h in ((1, 8), (8, 1), (0, 0)):
        with pytest.raises(artb.ArtbError) as e:
            gpu_ctx.render(scene.camera(8, 8), artb.RenderOptions(w, h, 1))
        assert e.value.code == -1
    # 32-bit pixel indices on the device: an oversized frame is refused, not wrapped. Called directly: the
    # check must come before any buffer is touched (the accumulation pointer here is not a real one)
    import ctypes as C
    c, p = scene.camera(30000, 30000).to_c(), artb.RenderOptions(30000, 30000, 1).to_c()
    rc = gpu_ctx._lib.artb_render_device(gpu_ctx._h, C.byref(c), C.byref(p), 16, None, None)
    assert rc == -1 and b"too large" in gpu_ctx._lib.artb_last_error(gpu_ctx._h)
    with pytest.raises(artb.ArtbError):
        gpu_ctx.render(scene.camera(8, 8), artb.RenderOptions(8, 8, 0))


def test_sample_ranges_far_from_zero(artb, gpu_ctx):
    """sample_begin near the top of the 32-bit index space: the Philox counter takes the index as it is"""
    scene = artb.Scene.build(6)
    gpu_ctx.upload(scene)
    cam = scene.camera(48, 48)
    a = gpu_ctx.render(cam, artb.RenderOptions(48, 48, 64, seed=2, sample_begin=2 ** 31 - 100), want_accum=True)[2]
    b = gpu_ctx.render(cam, artb.RenderOptions(48, 48, 64, seed=2, sample_begin=0), want_accum=True)[2]
    assert np.isfinite(a).all() and abs(a.mean() - b.mean()) < 0.25 * b.mean()
    assert not np.allclose(a, b)
