"""The C-ABI boundary: every function include/g753.h declares is exported by the built library and
bound by ffi.py, with no compute call made (runs without a GPU)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = importlib.import_module("ginger-lib_b200")
ffi = G.ffi


def declared_functions():
    src = open(os.path.join(ROOT, "include", "g753.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g753_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert names == sorted(n for n, _, _ in ffi.SYMBOLS)


def test_library_exports_every_symbol():
    graft = importlib.import_module("__graft_entry__")
    graft.build()                                   # nvcc cross-compiles here without a GPU
    lib = ctypes.CDLL(ffi.DEFAULT_LIB)
    for name in declared_functions():
        assert hasattr(lib, name), name
    L = ffi.Library()
    assert L.version().startswith(b"g753")
    assert L.group_coord_limbs(ffi.MNT6_G2) == 36
    # EvaluationDomain::new -> None (domain.rs:70-72) is decided without touching a device
    assert L.domain_check(ffi.FIELD_MNT6_FR, 15) == ffi.ERR_DOMAIN
    assert L.domain_check(ffi.FIELD_MNT6_FR, 14) == ffi.OK
    assert L.domain_check(ffi.FIELD_MNT4_FR, 29) == ffi.OK
    assert L.domain_check(ffi.FIELD_MNT4_FR, 30) == ffi.ERR_DOMAIN


def test_no_cpu_fallback_without_a_device():
    L = ffi.Library()
    if L.device_count_safe() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = L.ctx_create(0, ctypes.byref(h))
    assert rc == ffi.ERR_NO_DEVICE and not h.value
    with pytest.raises(ffi.G753Error):
        G.Context(0)
