"""The C-ABI library loads and exports exactly what include/vipcup.h declares (no compute calls: CPU-only test)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vipcup.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vip_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = _header_symbols()
    assert "vip_preprocess" in syms and "vip_last_error" in syms


def test_library_exports_every_declared_symbol():
    from vipcup_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "libvipcup.so was not built"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in _header_symbols():
        assert hasattr(L, s), f"{s} declared in vipcup.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in vipcup_b200/_lib.py"
    assert set(_lib.SIGNATURES) <= set(_header_symbols())
    L.vip_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.vip_version()


def test_argument_validation_without_gpu():
    """Errors are reported through return codes + vip_last_error, never exceptions across the ABI."""
    from vipcup_b200 import _lib

    L = _lib.lib()
    rc = L.vip_preprocess(None, 4, 200, 200, None, None, None, 224, 224, None, 0, None)
    assert rc == -1 and b"null" in L.vip_last_error()
    rc = L.vip_preprocess(None, 0, 200, 200, None, None, None, 224, 224, None, 0, None)
    assert rc == 0  # empty batch is a no-op


def test_ops_refuse_cpu_tensors():
    import torch

    from vipcup_b200 import ops

    with pytest.raises(ops.VipError):
        ops.preprocess(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), (8, 8))
