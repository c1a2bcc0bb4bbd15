"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; no compute
call works without a device (there is no CPU fallback)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for header in ("crt.h", "crt_render.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_every_declared_symbol_is_exported(crt):
    lib = ctypes.CDLL(crt.lib_path())
    names = _declared()
    assert len(names) >= 46 and "crt_render" in names and "crt_world_new" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_record_sizes(crt):
    assert crt.RAY_DTYPE.itemsize == 32 and crt.HIT_DTYPE.itemsize == 32


def test_render_record_sizes_match_the_header(crt):
    from crust_render_b200 import render_types as T
    assert T.MATERIAL_DTYPE.itemsize == 236 and T.LIGHT_DTYPE.itemsize == 116 and T.CAMERA_DTYPE.itemsize == 52
    assert T.SETTINGS_DTYPE.itemsize == 64 and T.STATS_DTYPE.itemsize == 64
    rec = T.OpenPBR()
    out = rec.copy()
    out["base_weight"] = 0
    ctypes.CDLL(crt.lib_path()).crt_material_default(ctypes.c_void_p(out.ctypes.data))
    assert out.tobytes() == rec.tobytes()  # openpbr.rs:130-173 defaults, library vs Python mirror


def test_no_cpu_fallback_without_a_device(crt):
    """Without a CUDA device every compute entry fails with CRT_ERR_NO_DEVICE instead of falling back."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    try:
        crt.init(0)
    except crt.CrtError as e:
        assert e.status == 4
    else:
        raise AssertionError("crt_init succeeded without a device")
