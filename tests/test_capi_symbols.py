"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/b200fem.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from dune_fem_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200fem.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    declared = _header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200fem.h but not exported"
    assert sorted(_capi.SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _capi.lib()
    h = C.c_void_p()
    rc = L.b200fem_ctx_create(0, None, C.byref(h))
    assert rc == _capi.ERR_CUDA
    assert b"no CPU fallback" in L.b200fem_last_error()


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "dune_fem_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "__none__", f"{f} mentions the oracle"
