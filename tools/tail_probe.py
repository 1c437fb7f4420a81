#!/usr/bin/env python3
"""How much of a render call is ramp-up and drain: the final scene at 800x800 rendered with 64 ... 4096 spp in one call each;
the marginal cost per spp comes from the two largest calls, the fixed part of a call is what is left. usage: tools/tail_probe.py"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import another_raytracer_b200 as A            # noqa: E402

W = H = 800
scene = A.Scene.build(8)
ctx = A.Context(0, 1 << 21)
ctx.upload(scene)
cam = scene.camera(W, H)
accum = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
rows = []
for spp in (256, 64, 256, 1024, 4096):
    st = ctx.render_device(cam, A.RenderOptions(W, H, spp, seed=1), accum.data_ptr(), 0)
    rows.append({"spp": spp, "elapsed_ms": st["elapsed_ms"], "launches": st["waves"], "rays": st["rays"]})
    print(json.dumps(rows[-1]))
a, b = rows[-2], rows[-1]
per_spp = (b["elapsed_ms"] - a["elapsed_ms"]) / (b["spp"] - a["spp"])
for r in rows[1:]:
    print(f"spp {r['spp']:5d}: {r['elapsed_ms']:9.2f} ms, fixed part {r['elapsed_ms'] - per_spp * r['spp']:7.2f} ms, waves {r['launches']}")
print("marginal ms per spp", per_spp, "->", W * H / per_spp / 1e3, "Msamples/s asymptotic")
