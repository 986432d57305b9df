"""The C-ABI shared library loads on a CPU-only host, exports every symbol include/lvi_exc_b200.h declares, agrees with the
ctypes mirror on structure sizes, and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from lvi_exc_b200 import _capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "lvi_exc_b200.h").read_text()


def header_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(lvi_[a-z0-9_]+)\s*\(", body)))


def test_every_declared_symbol_is_exported():
    lib = _capi.load()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lvi_exc_b200.h but not exported"
    assert set(_capi.ABI_SYMBOLS) == set(names), set(_capi.ABI_SYMBOLS) ^ set(names)


def test_abi_version_and_struct_sizes():
    lib = _capi.load()
    assert lib.lvi_abi_version() == 1
    lib.lvi_abi_sizeof.restype = C.c_int64
    assert lib.lvi_abi_sizeof(0) == C.sizeof(_capi.ProblemDesc)
    assert lib.lvi_abi_sizeof(1) == C.sizeof(_capi.SolveOptions)
    assert lib.lvi_abi_sizeof(2) == C.sizeof(_capi.SolveSummary)
    assert lib.lvi_abi_sizeof(3) == _capi.RAW_POINT_DTYPE.itemsize == 32
    assert lib.lvi_abi_sizeof(4) == _capi.SURFEL_POINT_DTYPE.itemsize == 64


def test_solve_options_defaults_are_ceres_defaults():
    lib = _capi.load()
    o = _capi.SolveOptions()
    lib.lvi_solve_options_default(C.byref(o))
    d = _capi.SolveOptions.default()
    for f, _ in _capi.SolveOptions._fields_:
        assert getattr(o, f) == getattr(d, f), f
    assert (o.initial_trust_region_radius, o.function_tolerance, o.gradient_tolerance, o.parameter_tolerance) == (1e4, 1e-6, 1e-10, 1e-8)


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point fails loudly"""
    lib = _capi.load()
    if lib.lvi_device_count() > 0:
        pytest.skip("a GPU is visible")
    ctx = C.c_void_p()
    rc = lib.lvi_ctx_create(0, None, 0, 1, C.byref(ctx))
    assert rc == _capi.LVI_ERR_NO_DEVICE and not ctx
    assert b"no CPU fallback" in lib.lvi_last_error()
    from lvi_exc_b200.backend import CudaBackend
    with pytest.raises(_capi.LviError):
        CudaBackend(0)


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setenv("LVI_EXC_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.LibraryMissing):
        _capi.load()
    monkeypatch.delenv("LVI_EXC_B200_LIB")
    monkeypatch.setattr(_capi, "_lib", None)
    _capi.load()
