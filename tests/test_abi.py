"""The C-ABI library loads and exports exactly the symbols include/pylc_b200.h declares
(no compute calls: this runs without a GPU)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pylc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PYLC_API[^;(]*?\b(pylc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pylc_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_declares_something():
    syms = declared_symbols()
    assert len(syms) >= 15 and "pylc_stitch_argmax_colour" in syms


def test_every_declared_symbol_is_exported_and_bound(lib):
    handle = lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pylc_[a-z0-9_]+)", out))
    declared = set(declared_symbols())
    assert declared == exported, (declared ^ exported)
    assert declared == set(lib.SIGNATURES), (declared ^ set(lib.SIGNATURES))
    for name in declared:
        assert getattr(handle, name) is not None


def test_host_only_entry_points(lib):
    import ctypes
    handle = lib.load()
    assert handle.pylc_abi_version() == lib.ABI_VERSION
    nH, nW = ctypes.c_int(), ctypes.c_int()
    assert handle.pylc_tile_grid(4000, 6000, 512, 512, ctypes.byref(nH), ctypes.byref(nW)) == 0
    assert (nH.value, nW.value) == (7, 11)
    assert handle.pylc_tile_grid(3584, 5632, 512, 256, ctypes.byref(nH), ctypes.byref(nW)) == 0
    assert (nH.value, nW.value) == (13, 21)
    assert handle.pylc_tile_grid(100, 100, 512, 512, ctypes.byref(nH), ctypes.byref(nW)) == 0
    assert (nH.value, nW.value) == (0, 0)
    assert b"geometry" in handle.pylc_error_string(-3)
    assert handle.pylc_error_string(0) == b"ok"


def test_library_is_sm100a_with_sass(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_missing_library_fails_loudly(monkeypatch, lib):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libpylc_b200.so")
    with pytest.raises(lib.PylcError):
        lib.load()
