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
    assert lib.artb_abi_version() == 2


def test_struct_sizes_match_header(artb):
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives them."""
    import subprocess, tempfile, os
    from another_raytracer_b200 import _abi
    names = {"artb_config": _abi.Config, "artb_scene_options": _abi.SceneOptions, "artb_scene_info": _abi.SceneInfo,
             "artb_flat_scene": _abi.FlatSceneView, "artb_hit": _abi.Hit, "artb_camera": _abi.Camera,
             "artb_render_params": _abi.RenderParams, "artb_stats": _abi.Stats, "artb_shade_record": _abi.ShadeRecord,
             "artb_graph_blob": _abi.GraphBlob, "artb_bvh_stats": _abi.BvhStats}
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


def test_flat_records_compile_as_c_with_the_documented_sizes(artb):
    """include/artb_flat.h is a C header (a foreign host may be C): sizes and alignments as the kernels expect them"""
    import subprocess, tempfile, os
    import flat_lib as F
    recs = {"artb_flat_prim": (F.PRIM, 16), "artb_flat_node8": (F.NODE, 16), "artb_flat_ref_node": (F.REF_NODE, 16),
            "artb_flat_xform_op": (F.XFORM_OP, 16), "artb_flat_xform": (F.XFORM, 4), "artb_flat_material": (F.MATERIAL, 16),
            "artb_flat_texture": (F.TEXTURE, 16), "artb_flat_image": (F.IMAGE, 8)}
    src = '#include "artb_flat.h"\n#include <stdio.h>\nint main(void){' + "".join(
        f'printf("{n} %zu %zu\\n", sizeof({n}), _Alignof({n}));' for n in recs) + "return 0;}"
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "s")
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    for line in out.splitlines():
        n, size, align = line.split()
        assert recs[n][0].itemsize == int(size) and recs[n][1] == int(align), line


def test_multi_device_config_is_validated(artb):
    """artb_config::n_devices / partition are checked before any device is touched"""
    import ctypes as C
    from another_raytracer_b200 import _abi
    lib = artb.load()
    assert lib.artb_device_count() >= 0
    h = C.c_void_p()
    arr = (C.c_int32 * _abi.MAX_DEVICES)()
    assert lib.artb_create(C.byref(_abi.Config(0, 0, 9, 0, arr)), C.byref(h)) == _abi.ARTB_E_INVALID
    assert lib.artb_create(C.byref(_abi.Config(0, 0, -1, 0, arr)), C.byref(h)) == _abi.ARTB_E_INVALID
    assert lib.artb_create(C.byref(_abi.Config(0, 0, 2, 7, arr)), C.byref(h)) == _abi.ARTB_E_INVALID
    assert not h.value


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
