"""CPU checks of the drop-in boundary: libvpz.so loads, exports every symbol include/vpz.h declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from vorbispizza_b200 import _native as N


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vpz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vpz_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(N.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(N.DEFAULT_LIB), "build with __graft_entry__.build()"
    lib = C.CDLL(N.DEFAULT_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_emulated_build_exports_the_same_abi(emu_lib_path):
    lib = C.CDLL(emu_lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_product_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu suite")
    lib = N.load()
    assert b"sm_100a" in lib.vpz_version()
    ctx = C.c_void_p()
    rc = lib.vpz_ctx_create(0, C.byref(ctx))
    assert rc in (N.VPZ_E_NO_DEVICE, N.VPZ_E_CUDA) and not ctx.value
    from vorbispizza_b200 import Context, VpzError
    with pytest.raises(VpzError):
        Context(0)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError):
        N.load(str(tmp_path / "libvpz.so"))


def test_error_strings():
    lib = N.load()
    for code in range(0, -12, -1):
        assert lib.vpz_strerror(code)
