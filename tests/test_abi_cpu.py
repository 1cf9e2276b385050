"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/totsu_b200.h
declares, and refuses to run without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT
from totsu_b200 import capi, host


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "totsu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 60
    for n in names:
        assert hasattr(L, n), "libtotsu_b200.so does not export %s" % n


def test_python_binding_covers_header():
    assert set(_declared_symbols()) == set(capi.SIGNATURES.keys())


def test_host_library_loads():
    host.hlib()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = capi.lib()
    st = L.tb_init(0)
    assert st != 0, "tb_init must fail without a CUDA device"
    assert b"CUDA" in L.tb_last_error() or b"cuda" in L.tb_last_error() or len(L.tb_last_error()) > 0
    out = C.c_float()
    st = L.tb_norm_f32(capi.View(1, 0, 1), C.byref(out))
    assert st != 0, "compute entry points must fail loudly when the backend is not initialised"
