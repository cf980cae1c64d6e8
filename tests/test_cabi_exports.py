"""CPU: the C-ABI library loads and exports every symbol include/hvae_b200.h declares (no compute)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    from hvae import _cabi

    protos = _cabi.declared_functions()
    assert len(protos) >= 20
    L = ctypes.CDLL(_cabi.LIB_PATH)
    for name in protos:
        assert hasattr(L, name), name
    L.hvae_version.restype = ctypes.c_int
    assert L.hvae_version() == 100
    L.hvae_strerror.restype = ctypes.c_char_p
    assert b"shape" in L.hvae_strerror(-1)


def test_sass_is_sm100a():
    from hvae import _cabi

    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:500]
