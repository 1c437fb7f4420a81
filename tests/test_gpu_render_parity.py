"""GPU parity, second gate (north_star): rendered images must match the reference's converged render within a stated
per-pixel RMSE / PSNR bound at matched spp.

Two protocols, both through the C ABI with host buffers (artb_render):

 A. noise-floor protocol, oracle computed HERE on the host cores: two independent oracle renders A, B at m spp (independent
    MT19937 streams, re-seeded after the scene build) define the Monte-Carlo noise floor RMSE(A,B) = sqrt(2)*sigma_m. An unbiased
    GPU image G with negligible noise of its own has RMSE(G,A) = sigma_m, i.e. 0.707 of the floor; a bias pushes
    RMSE(G,A) towards or beyond the floor. Bound asserted: RMSE(G,A) <= 0.85 * RMSE(A,B) (linear radiance clipped at 4),
    and the mean radiance of G within 2 % of the oracle's.
 B. converged protocol, oracle from committed fixtures (tests/golden/converged_*.npz, tools/make_converged.py, thousands
    of spp): the fixture's two halves give its own noise limit, PSNR_fixture = PSNR(half_a, half_b) + 6 dB. Bound
    asserted on the 8-bit gamma-2 images: PSNR(G, fixture) >= min(PSNR_fixture, 48 dB) - 3 dB, and >= 40 dB wherever
    the fixture itself is converged to 43 dB or better.
"""
import numpy as np
import pytest

from conftest import GOLDEN
import oracle_lib as O

pytestmark = pytest.mark.gpu

CLIP = 4.0      # linear radiance is clipped before comparing: a single firefly must not decide the RMSE


def rmse(a, b):
    a = np.clip(a, 0, CLIP)
    b = np.clip(b, 0, CLIP)
    return float(np.sqrt(np.mean((a - b) ** 2)))


def to8(lin):
    return (256.0 * np.clip(np.sqrt(np.clip(lin, 0, None)), 0.0, 0.999)).astype(np.uint8)


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse))


# alias, W, H, oracle spp per image, GPU spp
FLOOR_CASES = [(1, 120, 68, 16, 4096), (6, 96, 96, 32, 8192), (7, 64, 64, 32, 8192), (8, 96, 96, 16, 8192),
               (3, 80, 45, 16, 2048), (4, 80, 45, 16, 2048), (5, 80, 45, 64, 8192), (9, 96, 54, 8, 2048),
               (10, 96, 54, 8, 2048), (11, 96, 54, 8, 2048), (2, 80, 45, 16, 2048),
               (13, 96, 54, 8, 2048), (14, 96, 54, 8, 2048)]


@pytest.mark.parametrize("alias,w,h,m,n", FLOOR_CASES, ids=[f"scene{c[0]}" for c in FLOOR_CASES])
def test_gpu_image_within_noise_floor_of_oracle(artb, gpu_ctx, alias, w, h, m, n):
    a = O.OracleScene(alias).render(w, h, m, reseed=1234567)[1]
    b = O.OracleScene(alias).render(w, h, m, reseed=7654321)[1]
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    rgb, stats, accum, _ = gpu_ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, n, seed=7), want_accum=True)
    g = accum.astype(np.float64) / n
    assert stats["samples"] == w * h * n
    floor = rmse(a, b)
    ga, gb = rmse(g, a), rmse(g, b)
    assert floor > 0
    assert max(ga, gb) <= 0.85 * floor, f"RMSE(G,A)={ga:.5f} RMSE(G,B)={gb:.5f} vs noise floor RMSE(A,B)={floor:.5f}"
    mean_ref = 0.5 * (np.clip(a, 0, CLIP).mean() + np.clip(b, 0, CLIP).mean())
    assert abs(np.clip(g, 0, CLIP).mean() - mean_ref) <= 0.02 * mean_ref + 1e-4
    # the 8-bit image the ABI returns is write_color of the sums it returns (float vs double sqrt at bin edges aside)
    assert np.mean(rgb != to8(g)) < 1e-3


CONVERGED = sorted(GOLDEN.glob("converged_*.npz"))


@pytest.mark.parametrize("path", CONVERGED, ids=lambda p: p.stem)
def test_gpu_image_psnr_against_converged_oracle(artb, gpu_ctx, path):
    c = np.load(path)
    alias, spp = int(c["alias"]), int(c["spp"])
    h, w, _ = c["lin"].shape
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    # 4x the oracle's samples where that is cheap (the GPU image's own noise is then minor); the 64 Ki-spp fixtures of the
    # north-star scenes are matched sample for sample (expected: the fixture's limit - 3 dB, still well above 40 dB)
    n = 4 * spp if spp <= 8192 else spp
    rgb, stats, accum, _ = gpu_ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, n, seed=3), want_accum=True)
    g = accum.astype(np.float64) / n
    halves = psnr8(to8(c["half_a"]), to8(c["half_b"]))
    fixture_psnr = halves + 6.0                       # halves differ by 2 sigma_full
    got = psnr8(to8(g), to8(c["lin"]))
    print(f"{path.stem}: PSNR(G, converged oracle) = {got:.2f} dB; fixture noise limit {fixture_psnr:.2f} dB")
    # expected: the fixture's noise plus the GPU image's own (4x the samples: +1 dB) plus 8-bit quantisation, which
    # dominates above ~48 dB; bound = 3 dB under the fixture's limit, the limit itself capped at 48 dB
    # (matched sample counts: the GPU image carries as much noise as the fixture, the expected value is the limit - 3 dB)
    slack = 3.0 if n > spp else 3.8
    assert got >= min(fixture_psnr, 48.0) - slack, f"PSNR(G, converged)={got:.2f} dB, fixture noise limit {fixture_psnr:.2f} dB"
    if fixture_psnr >= 43.0:
        assert got >= 40.0
    if alias in (6, 7, 8):
        # north_star: "PSNR >= 40 dB against the reference's converged image" on the Next Week final scene / Cornell box
        assert fixture_psnr >= 45.0, "the committed fixture must itself be converged well beyond 40 dB"
        assert got >= 40.0, f"north-star bound: PSNR(G, converged reference) = {got:.2f} dB < 40 dB"

    def box(x):      # linear-space bias check on 8x8 box-filtered means (separates bias from noise)
        hh, ww = (x.shape[0] // 8) * 8, (x.shape[1] // 8) * 8
        return np.clip(x[:hh, :ww], 0, CLIP).reshape(hh // 8, 8, ww // 8, 8, 3).mean(axis=(1, 3))
    d = np.abs(box(g) - box(c["lin"]))
    tol = 4.0 * np.abs(box(c["half_a"]) - box(c["half_b"])).mean() + 0.01 * box(c["lin"]).mean()
    assert d.mean() <= tol


def test_render_is_deterministic_given_seed_up_to_float_sum_order(artb, gpu_ctx):
    scene = artb.Scene.build(6)
    gpu_ctx.upload(scene)
    cam = scene.camera(64, 64)
    a = gpu_ctx.render(cam, artb.RenderOptions(64, 64, 64, seed=5), want_accum=True)[2]
    b = gpu_ctx.render(cam, artb.RenderOptions(64, 64, 64, seed=5), want_accum=True)[2]
    c = gpu_ctx.render(cam, artb.RenderOptions(64, 64, 64, seed=6), want_accum=True)[2]
    assert np.allclose(a, b, rtol=1e-4, atol=1e-4)          # same paths; only the float atomic order may differ
    assert not np.allclose(a, c, rtol=1e-3, atol=1e-3)


def test_sample_split_and_tile_split_compose_to_the_whole(artb, gpu_ctx):
    """partition A: two disjoint sample ranges sum to the full render; partition B: two tile parts have disjoint support
    and sum to the full render (what the multi-GPU reduce relies on)."""
    from another_raytracer_b200 import distributed as D
    scene = artb.Scene.build(6)
    gpu_ctx.upload(scene)
    w = h = 96
    cam = scene.camera(w, h)
    full = gpu_ctx.render(cam, artb.RenderOptions(w, h, 32, seed=9), want_accum=True)[2]
    parts = []
    for r in range(2):
        sh = D.sample_share(32, r, 2)
        parts.append(gpu_ctx.render(cam, artb.RenderOptions(w, h, sh.samples, seed=9, sample_begin=sh.sample_begin),
                                    want_accum=True)[2])
    assert np.allclose(parts[0] + parts[1], full, rtol=1e-4, atol=1e-4)
    tiles = []
    for r in range(2):
        sh = D.tile_share(32, r, 2)
        acc = gpu_ctx.render(cam, artb.RenderOptions(w, h, 32, seed=9, tile_size=32, tile_parts=2, tile_part=r), want_accum=True)[2]
        mask = D.owned_pixel_mask(w, h, 32, sh)
        assert np.all(acc[~mask] == 0)
        tiles.append(acc)
    assert np.allclose(tiles[0] + tiles[1], full, rtol=1e-4, atol=1e-4)


def test_variance_mask_mode_saves_samples_and_stays_close(artb, gpu_ctx):
    scene = artb.Scene.build(5)          # simple_light: the black background converges immediately
    gpu_ctx.upload(scene)
    w, h = 120, 72
    cam = scene.camera(w, h)
    fixed = gpu_ctx.render(cam, artb.RenderOptions(w, h, 512, seed=2), want_accum=True)
    rgb, st, acc, cnt = gpu_ctx.render(cam, artb.RenderOptions(w, h, 512, seed=2, mode=1, adaptive_min_spp=32, adaptive_batch_spp=32,
                                                                adaptive_threshold=0.03), want_accum=True, want_count=True)
    assert cnt.min() >= 32 and cnt.max() <= 512
    assert st["samples"] == int(cnt.sum()) < 0.8 * w * h * 512
    mean_fixed = fixed[2].astype(np.float64) / 512
    mean_adapt = acc.astype(np.float64) / np.maximum(cnt, 1)[..., None]
    assert psnr8(to8(mean_adapt), to8(mean_fixed)) > 30.0


@pytest.mark.parametrize("alias", [8, 6])
def test_variance_mask_against_converged_oracle(artb, gpu_ctx, alias):
    """north_star (3): the per-pixel variance mask compared with the ORACLE's converged image (64 Ki spp), not with the
    GPU's own fixed render: at a relative-standard-error target of 5 % and a cap of 32 Ki spp it must retire pixels (mean
    spp well under the cap), reach at least the quality a FIXED render of the same total sample count reaches (within
    1.5 dB: it trades samples between pixels, it does not lose them), and stay above 34 dB."""
    c = np.load(GOLDEN / f"converged_{alias}_128x128.npz")
    h, w, _ = c["lin"].shape
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    cam = scene.camera(w, h)
    cap = 32768
    _, st, acc, cnt = gpu_ctx.render(cam, artb.RenderOptions(w, h, cap, seed=5, mode=1, adaptive_min_spp=256, adaptive_batch_spp=256,
                                                             adaptive_threshold=0.05), want_accum=True, want_count=True)
    adaptive = acc.astype(np.float64) / np.maximum(cnt, 1)[..., None]
    mean_spp = st["samples"] / (w * h)
    fixed_spp = max(int(mean_spp), 1)
    _, _, facc, _ = gpu_ctx.render(cam, artb.RenderOptions(w, h, fixed_spp, seed=6), want_accum=True)
    fixed = facc.astype(np.float64) / fixed_spp
    ref8 = to8(c["lin"])
    p_adaptive, p_fixed = psnr8(to8(adaptive), ref8), psnr8(to8(fixed), ref8)
    print(f"alias {alias}: variance mask {mean_spp:.0f} spp/pixel (cap {cap}, {np.mean(cnt < cap) * 100:.0f} % of the pixels retired early) "
          f"PSNR {p_adaptive:.2f} dB; fixed {fixed_spp} spp PSNR {p_fixed:.2f} dB")
    assert cnt.min() >= 256 and cnt.max() <= cap and st["samples"] == int(cnt.sum())
    assert mean_spp < 0.8 * cap, "the mask must retire converged pixels"
    assert p_adaptive >= 34.0
    assert p_adaptive >= p_fixed - 1.5


def test_fixed_mode_counts_and_device_tile_lists(artb, gpu_ctx):
    """artb_render's count buffer in ARTB_MODE_FIXED (include/artb.h: samples per pixel) and the tile lists built on the
    device: odd image sizes, three parts, edge tiles narrower than the tile size"""
    from another_raytracer_b200 import distributed as D
    scene = artb.Scene.build(6)
    gpu_ctx.upload(scene)
    w, h, spp = 100, 70, 8
    cam = scene.camera(w, h)
    _, st, full, cnt = gpu_ctx.render(cam, artb.RenderOptions(w, h, spp, seed=4), want_accum=True, want_count=True)
    assert np.all(cnt == spp) and st["samples"] == w * h * spp
    total = np.zeros_like(full)
    covered = np.zeros((h, w), dtype=int)
    for parts, ts in ((3, 32), (2, 16), (5, 7)):
        total[:] = 0; covered[:] = 0
        for part in range(parts):
            _, st, acc, c = gpu_ctx.render(cam, artb.RenderOptions(w, h, spp, seed=4, tile_size=ts, tile_parts=parts, tile_part=part),
                                           want_accum=True, want_count=True)
            mask = D.owned_pixel_mask(w, h, ts, D.tile_share(ts, part, parts))
            assert np.array_equal(c, np.where(mask, spp, 0)), f"parts={parts} ts={ts} part={part}"
            assert np.all(acc[~mask] == 0) and st["samples"] == int(mask.sum()) * spp
            total += acc; covered += mask
        assert np.all(covered == 1)
        assert np.allclose(total, full, rtol=1e-4, atol=1e-4)


def test_engine_mirror_run_contract(artb, gpu_ctx, capsys):
    """Engine.set_scene / run: elapsed ms on success, -1 + 'Invalid input scene!' on an empty scene (engine.h:30-36)"""
    scene = artb.Scene.build(2)
    eng = artb.Engine(scene.camera(64, 36), artb.EngineMode.single, width=64, height=36, samples_per_pixel=8, context=gpu_ctx)
    out = np.zeros(64 * 36 * 3, dtype=np.uint8)
    assert eng.run(out) == -1
    assert "Invalid input scene!" in capsys.readouterr().err
    eng.set_scene(scene)
    ms = eng.run(out)
    assert ms >= 0 and out.reshape(36, 64, 3)[0].mean() > 100          # sky at the top of the image (row 0 = top)
    adaptive = artb.Engine(scene.camera(64, 36), artb.EngineMode.adaptive, width=64, height=36, context=gpu_ctx)
    adaptive.set_scene(scene)
    with pytest.raises(ValueError):          # 64 is not a multiple of 12 (engine.h:182-183)
        adaptive.run(out)
