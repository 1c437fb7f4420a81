"""The drop-in proof: the reference's OWN host program — its scene_manager, its classes, its main flow
(src/main.cpp:29-46) — with `engine` replaced by the `cuda_engine` binding of INTEGRATION.md (oracle/ref_glue/cuda_engine.h,
include/artb_reference_adapter.hpp), compiled against the unmodified reference and linked with libartb.so
(oracle/_ref/ref_cuda_engine, built by oracle/Makefile where /root/reference exists; the binary travels to the GPU box).
Its image must be the one this repository's own host mirror renders for the same scene, seed and sample count."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

EXE = ROOT / "oracle" / "_ref" / "ref_cuda_engine"


def _run(alias, w, h, spp, mode, out, n_devices=1):
    env = dict(os.environ, ARTB_ASSETS=str(ROOT / "assets"))
    r = subprocess.run([str(EXE), str(alias), str(w), str(h), str(spp), mode, str(out), str(n_devices)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    meta = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    return np.fromfile(out, dtype=np.uint8).reshape(h, w, 3), meta


@pytest.mark.skipif(not EXE.exists(), reason="oracle/_ref/ref_cuda_engine not built (needs /root/reference once)")
@pytest.mark.parametrize("alias,w,h,spp", [(6, 96, 96, 32), (8, 96, 96, 16), (1, 120, 68, 16), (9, 96, 54, 16)])
def test_reference_host_program_renders_through_the_backend(artb, tmp_path, alias, w, h, spp):
    img, meta = _run(alias, w, h, spp, "single", tmp_path / "o.raw")
    assert meta["samples"] == w * h * spp and meta["elapsed_ms"] >= 0
    scene = artb.Scene.build(alias)
    ctx = artb.Context(0, 1 << 16)
    ctx.upload(scene)
    rgb, st, _, _ = ctx.render(scene.camera(w, h), artb.RenderOptions(w, h, spp, seed=1))
    ctx.close()
    assert st["rays"] == meta["rays"], "the same paths: the reference's graph and the mirror's flatten to the same scene"
    assert np.mean(img != rgb) < 5e-3


@pytest.mark.skipif(not EXE.exists(), reason="oracle/_ref/ref_cuda_engine not built (needs /root/reference once)")
def test_reference_host_adaptive_and_parallel_modes(artb, tmp_path):
    n_dev = artb.load().artb_device_count()
    single, _ = _run(6, 96, 96, 32, "single", tmp_path / "s.raw")
    adaptive, meta = _run(6, 96, 96, 32, "adaptive", tmp_path / "a.raw")
    assert 0 < meta["samples"] < 96 * 96 * 32                   # engine_mode::adaptive traces only part of the pixels
    assert np.mean(np.abs(adaptive.astype(int) - single.astype(int))) < 12
    r = subprocess.run([str(EXE), "6", "100", "100", "8", "adaptive", str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert r.returncode != 0 and "perfectly fit big square size" in r.stderr      # engine.h:182-183
    if n_dev >= 2:
        for mode in ("images", "stripes"):
            img, meta = _run(6, 96, 96, 32, mode, tmp_path / "p.raw", n_devices=min(n_dev, 4))
            assert meta["n_devices"] == min(n_dev, 4)
            assert np.mean(img != single) < 5e-3
