"""CPU tier: the product library exists, loads, and exports every symbol include/*.h declares.
No compute is attempted (there is no GPU in this tier, and there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

from superpang_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "superpang_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert set(declared_symbols()) == set(_capi.EXPORTS)


def test_cuda_library_exports_all_symbols():
    assert os.path.exists(_capi.LIB_PATH), "libsuperpang_b200.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    lib.spb_build_kind.restype = ctypes.c_char_p
    assert lib.spb_build_kind() == b"cuda sm_100a"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.SpbError):
        _capi.Engine(device=0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.load_library(os.path.join(tmp_path, "nope.so"))
