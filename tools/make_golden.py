#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref/ref_tool).

Run in the build container (needs /root/reference to have built oracle/_ref; `make -C oracle`). The GPU box only
ever sees the committed outputs:

  scene_info.json            per alias: RNG position after the build, structure counts, FNV-1a hash of the graph
                             token stream, camera preset, first RNG values, get_sphere_uv table
  trace_<alias>.npz          closest-hit queries recorded while the reference path-traces a sparse pixel grid:
                             rays, hit flag, winning leaf id (depth-first order), t, p, normal, u, v, front_face,
                             material type, and the xi each constant_medium consumed
  image_<alias>_<W>x<H>x<spp>.npz   single-mode render (bit-reproducible): RGB8 + linear float64 means
"""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
TOOL = ROOT / "oracle" / "_ref" / "ref_tool"
OUT = ROOT / "tests" / "golden"

TRACE = {   # alias: (W, H, stride, spp, max_rays)
    1: (400, 225, 9, 1, 3000), 2: (200, 112, 12, 1, 600), 3: (200, 112, 12, 1, 600), 4: (200, 112, 12, 1, 600),
    5: (200, 112, 8, 1, 1200), 6: (300, 300, 10, 1, 4000), 7: (300, 300, 12, 1, 3000), 8: (400, 400, 11, 1, 5000),
    9: (320, 180, 9, 1, 2500), 10: (320, 180, 9, 1, 2500), 11: (320, 180, 9, 1, 2500),
    13: (320, 180, 9, 1, 2500), 14: (320, 180, 9, 1, 2500),        # cow / dino under metal (BASELINE configs[3])
}
IMAGES = [(1, 120, 68, 4), (6, 96, 96, 8), (7, 64, 64, 8), (8, 96, 96, 4), (9, 96, 54, 2), (10, 96, 54, 2), (11, 96, 54, 2),
          (3, 80, 45, 4), (4, 80, 45, 4), (5, 80, 45, 8), (2, 80, 45, 4), (13, 96, 54, 2), (14, 96, 54, 2)]
ALIASES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14]     # 12 (height field) is not a reference scene: the oracle port covers it


def run(*args):
    env = dict(os.environ, ARTB_ASSETS=str(ROOT / "assets"))
    r = subprocess.run([str(TOOL), *map(str, args)], capture_output=True, text=True, env=env, check=True)
    return r.stdout


def last_json(text):
    return json.loads([l for l in text.splitlines() if l.startswith("{")][-1])


def read_trace(path):
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw, dtype="<i8", count=8)
    assert hdr[0] == 0x54545241 and hdr[1] == 1
    n, n_media, n_leaves = int(hdr[2]), int(hdr[3]), int(hdr[4])
    off = 64

    def take(dtype, count, shape=None):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a.reshape(shape) if shape else a

    d = {"n_media": n_media, "n_leaves": n_leaves, "alias": int(hdr[5]), "width": int(hdr[6]), "height": int(hdr[7])}
    d["rays"] = take("<f8", n * 7, (n, 7))
    d["tmax"] = take("<f8", n)
    for k in ("hit", "prim", "ff", "mat", "bounce"):
        d[k] = take("<i4", n)
    d["t"] = take("<f8", n)
    d["p"] = take("<f8", n * 3, (n, 3))
    d["normal"] = take("<f8", n * 3, (n, 3))
    d["uv"] = take("<f8", n * 2, (n, 2))
    d["xi"] = take("<f8", n * n_media, (n, n_media)) if n_media else np.zeros((n, 0))
    assert off == len(raw)
    return d


def main():
    """usage: make_golden.py [--info-only] [--only ALIAS ...]   (--only: refresh just these aliases, keep the other files)"""
    if not TOOL.exists():
        sys.exit("oracle/_ref/ref_tool missing: run `make -C oracle` where /root/reference exists")
    only = None
    if "--only" in sys.argv:
        only = {int(a) for a in sys.argv[sys.argv.index("--only") + 1:] if a.isdigit()}
    OUT.mkdir(parents=True, exist_ok=True)
    info_path = OUT / "scene_info.json"
    if only and info_path.exists():
        info = json.loads(info_path.read_text())
    else:
        info = {"rng_first": [float(x) for x in run("rng", 8).split()],
                "sphere_uv": [[float(x) for x in l.split()] for l in run("uv").splitlines() if l.strip()], "scenes": {}}
    import hashlib
    with tempfile.TemporaryDirectory() as tmp:
        for alias in ALIASES:
            if only and alias not in only:
                continue
            info["scenes"][str(alias)] = last_json(run("info", alias))
            p = os.path.join(tmp, "g.bin")
            run("graph", alias, p)
            info["scenes"][str(alias)]["stream_sha256"] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    info_path.write_text(json.dumps(info, indent=1))
    if "--info-only" in sys.argv:
        return
    with tempfile.TemporaryDirectory() as tmp:
        for alias, (w, h, stride, spp, max_rays) in TRACE.items():
            if only and alias not in only:
                continue
            p = os.path.join(tmp, "t.bin")
            print("trace", alias, last_json(run("trace", alias, w, h, stride, spp, max_rays, p)))
            d = read_trace(p)
            np.savez_compressed(OUT / f"trace_{alias}.npz", **d)
        for alias, w, h, spp in IMAGES:
            if only and alias not in only:
                continue
            p = os.path.join(tmp, "i.raw")
            print("render", alias, last_json(run("render", alias, w, h, spp, p)))
            rgb = np.fromfile(p, dtype=np.uint8).reshape(h, w, 3)
            lin = np.fromfile(p + ".lin", dtype="<f8").reshape(h, w, 3)
            np.savez_compressed(OUT / f"image_{alias}_{w}x{h}x{spp}.npz", rgb=rgb, lin=lin, alias=alias, spp=spp)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
