"""engine_mode::adaptive of the reference (engine.h:151-333) on the GPU (csrc/device/adaptive_squares.cuh).

The GPU keys every sample by (pixel, sample index), so a pixel traced by the adaptive mode has exactly the radiance it
has in a full render with the same seed. That makes the mode checkable pixel by pixel: `replay_reference_adaptive`
below restates engine.h:96-149 and :178-305 in numpy (TEST INFRASTRUCTURE) with the full render standing in for
_stochastic_sample; its output and its traced/interpolated map must equal what the device produced.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def write_color_int(lin_sum, spp):
    """color.h:6-22 with T = int"""
    v = np.sqrt(lin_sum.astype(np.float64) * (1.0 / spp))
    return (256 * np.clip(v, 0.0, 0.999)).astype(np.int64)


def corners_differ(work, x, y, size):
    """_compute_corners_heuristic, engine.h:96-137"""
    c1, c2 = work[y, x], work[y, x + size - 1]
    c3, c4 = work[y + size - 1, x], work[y + size - 1, x + size - 1]
    d2 = lambda a, b: int(((a - b) ** 2).sum())
    return d2(c1, c2) > 100 or d2(c2, c4) > 100 or d2(c4, c3) > 100 or d2(c3, c1) > 100


def interpolate_square(work, x, y, size):
    """interpolate_square + _interpolate, engine.h:186-219, 139-149"""
    x1, x2, y1, y2 = x, x + size - 1, y, y + size - 1
    q11, q12 = work[y1, x1].astype(np.float64), work[y2, x1].astype(np.float64)
    q21, q22 = work[y1, x2].astype(np.float64), work[y2, x2].astype(np.float64)
    xd, yd = x2 - x1, y2 - y1
    for l in range(size):
        for k in range(size):
            if work[y + l, x + k, 0] >= 0:
                continue
            xx, yy = x + k, y + l
            r1 = ((x2 - xx) * q11) * (1.0 / xd) + ((xx - x1) * q21) * (1.0 / xd)
            r2 = ((x2 - xx) * q12) * (1.0 / xd) + ((xx - x1) * q22) * (1.0 / xd)
            work[y + l, x + k] = (((y2 - yy) * r1) * (1.0 / yd) + ((yy - y1) * r2) * (1.0 / yd)).astype(np.int64)


def replay_reference_adaptive(full8, w, h):
    """process_square for every 12x12 square (engine.h:236-305); full8[y, x] = write_color<int> of the pixel's samples"""
    work = -np.ones((h, w, 3), dtype=np.int64)
    traced = np.zeros((h, w), dtype=bool)

    def trace(x, y):
        work[y, x] = full8[y, x]
        traced[y, x] = True

    def corners(x, y, s):
        for cx, cy in ((x, y), (x + s - 1, y), (x, y + s - 1), (x + s - 1, y + s - 1)):
            trace(cx, cy)

    for j in range(0, h, 12):
        for i in range(0, w, 12):
            corners(i, j, 12)
            if not corners_differ(work, i, j, 12):
                interpolate_square(work, i, j, 12)
                continue
            for l in range(j, j + 12, 6):
                for k in range(i, i + 12, 6):
                    corners(k, l, 6)
                    if not corners_differ(work, k, l, 6):
                        interpolate_square(work, k, l, 6)
                        continue
                    for n in range(l, l + 6, 3):
                        for m in range(k, k + 6, 3):
                            corners(m, n, 3)
                            if corners_differ(work, m, n, 3):
                                for px, py in ((m + 1, n), (m, n + 1), (m + 1, n + 1), (m + 2, n + 1), (m + 1, n + 2)):
                                    trace(px, py)
                            else:
                                interpolate_square(work, m, n, 3)
    return work, traced


@pytest.mark.parametrize("alias,w,h,spp", [(6, 96, 96, 256), (8, 120, 96, 128), (1, 144, 84, 64)])
def test_adaptive_squares_equal_the_replayed_reference_logic(artb, gpu_ctx, alias, w, h, spp):
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    cam = scene.camera(w, h)
    _, _, full_sum, _ = gpu_ctx.render(cam, artb.RenderOptions(w, h, spp, seed=3), want_accum=True)
    rgb, st, acc, cnt = gpu_ctx.render(cam, artb.RenderOptions(w, h, spp, seed=3, mode=2), want_accum=True, want_count=True)
    traced_gpu = cnt == spp
    assert np.all((cnt == 0) | traced_gpu)
    assert st["samples"] == int(traced_gpu.sum()) * spp
    # a traced pixel is the pixel of the full render (same samples; only the order of the float atomics differs)
    assert np.allclose(acc[traced_gpu], full_sum[traced_gpu], rtol=2e-4, atol=1e-4)
    want, traced_ref = replay_reference_adaptive(write_color_int(full_sum, spp), w, h)
    # float-sum order can move a value across an integer boundary, and with it — rarely — a decision: allow 0.5 %
    assert (traced_gpu != traced_ref).mean() <= 0.005
    same = traced_gpu == traced_ref
    assert (np.abs(rgb.astype(np.int64) - want).max(axis=2)[same] > 1).mean() <= 0.005
    # the strategy must actually save work on these scenes, and trace every corner of every big square
    assert 0.05 < traced_gpu.mean() < 0.99      # noisy images subdivide almost everywhere, as they do in the reference
    assert traced_gpu[0::12, 0::12].all() and traced_gpu[11::12, 11::12].all()


def test_adaptive_squares_flat_image_traces_corners_only(artb, gpu_ctx):
    """simple_light's upper half is black sky: those squares must not subdivide (4 traced pixels of 144)"""
    scene = artb.Scene.build(5)
    gpu_ctx.upload(scene)
    w, h = 96, 48
    rgb, st, _, cnt = gpu_ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, 32, seed=1, mode=2), want_count=True)
    flat = [(x, y) for y in range(0, h, 12) for x in range(0, w, 12) if (cnt[y:y + 12, x:x + 12] > 0).sum() == 4]
    assert len(flat) >= 4
    for x, y in flat:
        sq = rgb[y:y + 12, x:x + 12].astype(int)
        assert np.abs(sq - sq[0, 0]).max() <= 10      # corners within the heuristic's threshold, interior interpolated between them


def test_adaptive_squares_size_contract(artb, gpu_ctx):
    scene = artb.Scene.build(2)
    gpu_ctx.upload(scene)
    with pytest.raises(artb.ArtbError) as e:
        gpu_ctx.render(scene.camera(100, 60), artb.RenderOptions(100, 60, 4, mode=2))
    assert e.value.code == -6 and "big square size" in str(e.value)
