"""The C-ABI library loads and exports every symbol include/artb.h declares (no compute calls: no GPU here)."""
import ctypes
import re

from conftest import ROOT


def _declared():
    text = (ROOT / "include" / "artb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(artb_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(artb):
    assert _declared() == sorted(artb.SYMBOLS)


def test_library_exports_every_declared_symbol(artb):
    lib = ctypes.CDLL(str(artb.LIB_PATH))
    for name in _declared():
        assert hasattr(lib, name), f"libartb.so does not export {name}"
    assert lib.artb_abi_version() == 1


def test_struct_sizes_match_header(artb):
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives them."""
    import subprocess, tempfile, os
    from another_raytracer_b200 import _abi
    names = {"artb_config": _abi.Config, "artb_scene_options": _abi.SceneOptions, "artb_scene_info": _abi.SceneInfo,
             "artb_flat_scene": _abi.FlatSceneView, "artb_hit": _abi.Hit, "artb_camera": _abi.Camera,
             "artb_render_params": _abi.RenderParams, "artb_stats": _abi.Stats}
    src = '#include "artb.h"\n#include <stdio.h>\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "s")
        subprocess.run(["gcc", "-I", str(ROOT / "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        n, size = line.split()
        assert ctypes.sizeof(names[n]) == int(size), n


def test_no_device_fails_loudly(artb):
    """Without a CUDA device artb_create must fail with ARTB_E_NO_DEVICE — there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        ctx = artb.Context(0, 1 << 10)
        ctx.close()
        return
    try:
        artb.Context(0)
    except artb.ArtbError as e:
        assert e.code == -2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("artb_create succeeded without a GPU")
