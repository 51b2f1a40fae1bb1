"""The C-ABI library loads on a GPU-less machine, exports every symbol include/glia_rd.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os

import pytest

from glia_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "glia_rd.h")


def test_header_and_binding_agree():
    declared = _capi.declared_symbols(HEADER)
    assert declared, "no symbols parsed from the header"
    assert sorted(_capi.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(cuda_lib):
    lib = C.CDLL(cuda_lib)
    for name in _capi.declared_symbols(HEADER):
        assert hasattr(lib, name), name
    lib.glia_rd_build_info.restype = C.c_char_p
    assert lib.glia_rd_build_info() == b"cuda-sm_100a"
    assert lib.glia_rd_abi_version() == 2


def test_no_cpu_fallback(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from glia_b200.rd import RDHandle
    with pytest.raises(_capi.GliaRdError, match="no CUDA device"):
        RDHandle(32, "f32", lib_path=cuda_lib)


def test_missing_library_is_loud(tmp_path):
    with pytest.raises(_capi.GliaRdError, match="no CPU fallback"):
        _capi.load_library(str(tmp_path / "libglia_rd.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "glia_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle's", ""), os.path.join(dirpath, f)
