#!/usr/bin/env python3
"""Render BASELINE.json's configurations at their FULL sizes through the Engine mirror (the call a user makes) and write
the images + a JSON summary into gpurun_out/ (copied to profiles/ by hand). usage: tools/full_configs.py [outdir]"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import another_raytracer_b200 as A            # noqa: E402
from another_raytracer_b200 import _abi       # noqa: E402

CONFIGS = [
    ("rtiow_random_400x225_10spp", 1, 400, 225, 10, {}),
    ("cornell_600x600_1000spp", 6, 600, 600, 1000, {}),
    ("final_scene_800x800_10000spp", 8, 800, 800, 10000, {}),
    ("cow_1920x1080_256spp", 10, 1920, 1080, 256, {}),
    ("dino_1920x1080_256spp", 11, 1920, 1080, 256, {}),
    ("heightfield_1M_3840x2160_64spp", 12, 3840, 2160, 64, dict(mesh_as_soup=True, device_bvh=True)),
]


def main():
    out = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out")
    out.mkdir(parents=True, exist_ok=True)
    ctx = A.Context(0, 1 << 21)
    rows = []
    for name, alias, w, h, spp, kw in CONFIGS:
        t0 = time.perf_counter()
        scene = A.Scene.build(alias, **kw)
        t_build = time.perf_counter() - t0
        eng = A.Engine(scene.camera(w, h), A.EngineMode.single, width=w, height=h, samples_per_pixel=spp, context=ctx)
        t0 = time.perf_counter()
        eng.set_scene(scene)
        img = np.zeros(w * h * 3, dtype=np.uint8)
        ms = eng.run(img)
        wall = time.perf_counter() - t0
        st = eng.last_stats
        png = out / f"{name}.png"
        _abi.load().artb_save_png(str(png).encode(), w, h, 3, img.ctypes.data)
        rows.append(dict(config=name, width=w, height=h, spp=spp, engine_run_ms=ms, set_scene_plus_run_wall_s=round(wall, 3),
                         host_scene_build_s=round(t_build, 3), samples=st["samples"], rays=st["rays"],
                         msamples_per_s=round(st["samples"] / st["elapsed_ms"] / 1e3, 1), mrays_per_s=round(st["rays"] / st["elapsed_ms"] / 1e3, 1),
                         device_bvh=ctx.last_bvh_build if kw.get("device_bvh") else None, png=png.name))
        print(json.dumps(rows[-1]), flush=True)
    (out / "full_configs.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
