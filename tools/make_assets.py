#!/usr/bin/env python3
"""Recipe that produced assets/textures/earthmap.png and assets/models/capsule/capsule.png (run once, in the build
container, where /root/reference exists).

The reference decodes its two JPEG textures with stb_image (src/utils/imageio.cpp:11-15). This repository ships no JPEG
decoder (out of scope); instead the texels are decoded ONCE by the reference's own decoder
(`oracle/_ref/ref_tool decode`, i.e. imageio::load_image of the unmodified reference) and stored losslessly as PNG, so
that every texture lookup of the product and of the oracle reads bit-identical texels. capsule.mtl's map_Kd is pointed at
the PNG accordingly. A user texture that is a JPEG must be converted the same way (any lossless re-encode will do).

    python tools/make_assets.py            # rewrites the two PNGs and verifies them
    python tools/make_assets.py --check    # only verifies: PNG texels == what the reference decodes from its JPEG
"""
import ctypes as C
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
TOOL = ROOT / "oracle" / "_ref" / "ref_tool"
JOBS = [(REF / "textures" / "earthmap.jpg", ROOT / "assets" / "textures" / "earthmap.png"),
        (REF / "models" / "capsule" / "capsule.jpg", ROOT / "assets" / "models" / "capsule" / "capsule.png")]


def decode(path):
    """texels as the reference's decoder produces them: ref_tool decode -> 'w h bpp\\n' + raw bytes"""
    with tempfile.TemporaryDirectory() as tmp:
        out = Path(tmp) / "t.raw"
        subprocess.run([str(TOOL), "decode", str(path), str(out)], check=True, capture_output=True)
        raw = out.read_bytes()
    head, _, body = raw.partition(b"\n")
    w, h, bpp = (int(x) for x in head.split())
    return np.frombuffer(body, dtype=np.uint8, count=w * h * bpp).reshape(h, w, bpp)


def main():
    if not TOOL.exists() or not REF.exists():
        sys.exit("needs /root/reference and oracle/_ref/ref_tool (make -C oracle)")
    sys.path.insert(0, str(ROOT))
    import another_raytracer_b200 as A
    lib = A.load()
    for src, dst in JOBS:
        texels = decode(src)
        h, w, bpp = texels.shape
        if "--check" not in sys.argv:
            buf = np.ascontiguousarray(texels)
            if lib.artb_save_png(str(dst).encode(), w, h, bpp, buf.ctypes.data) != 0:
                sys.exit(f"cannot write {dst}")
        again = decode(dst)              # stb decodes PNG too: the round trip must be exact
        assert again.shape == texels.shape and np.array_equal(again, texels), f"{dst} does not hold the reference's texels"
        print(f"{dst.relative_to(ROOT)}: {w}x{h}x{bpp}, identical to the reference's decode of {src}")


if __name__ == "__main__":
    main()
