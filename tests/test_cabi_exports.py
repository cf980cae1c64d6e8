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


def test_gyroplane_backward_grid_fills_whole_waves():
    """Host-only planner of the SIMT gyroplane backward: plane chunks are sized so that the pair kernel's CTAs fill the
    740 resident slots (5 per SM at D <= 16) in whole waves - config 2 (4096 x 10 -> 600) ran 800 CTAs of 24 planes
    (a second wave 8 % full) before; now one wave."""
    from hvae import _cabi

    L = ctypes.CDLL(_cabi.LIB_PATH)
    L.hvae_gyroplane_bwd_plan.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64] + [ctypes.POINTER(ctypes.c_int)] * 3
    ppc, nch, ctas = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.hvae_gyroplane_bwd_plan(4096, 10, 600, ctypes.byref(ppc), ctypes.byref(nch), ctypes.byref(ctas)) == 0
    assert (ppc.value, nch.value, ctas.value) == (27, 23, 736)
    for B, D, P in [(128, 2, 16), (1024, 5, 100), (4096, 10, 600), (65536, 16, 1024), (300, 33, 50), (1, 3, 1)]:
        assert L.hvae_gyroplane_bwd_plan(B, D, P, ctypes.byref(ppc), ctypes.byref(nch), ctypes.byref(ctas)) == 0
        assert ppc.value >= 1 and nch.value >= 1 and ppc.value * nch.value >= P and ppc.value * (nch.value - 1) < P
        slots = 148 * (5 if D <= 16 else 2)
        waves = -(-ctas.value // slots)
        assert ctas.value > (waves - 1) * slots + 0.5 * slots or waves == 1   # the last wave is at least half full
    assert L.hvae_gyroplane_bwd_plan(10, 65, 10, None, None, None) != 0      # D beyond the SIMT path
