"""The C-ABI library loads without a GPU and exports every symbol include/crt.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "crt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(crt):
    lib = ctypes.CDLL(crt.lib_path())
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_record_sizes(crt):
    assert crt.RAY_DTYPE.itemsize == 32 and crt.HIT_DTYPE.itemsize == 32
