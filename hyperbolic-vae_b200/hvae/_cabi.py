"""ctypes binding of libhvae_b200.so (the C ABI in include/hvae_b200.h).  Fails loudly when the
library is missing: there is no fallback implementation."""
import ctypes
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HVAE_LIB_PATH") or os.path.join(HERE, "_lib", "libhvae_b200.so")  # override: experiment builds
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "hvae_b200.h")

_lib = None
_checked_devices = set()

_C = ctypes
_TYPES = {
    "const float*": _C.c_void_p, "const int64_t*": _C.c_void_p, "float*": _C.c_void_p, "int*": _C.c_void_p, "void*": _C.c_void_p, "const void*": _C.c_void_p,
    "int64_t": _C.c_int64, "float": _C.c_float, "int": _C.c_int, "uint32_t": _C.c_uint32,
    "uint64_t": _C.c_uint64, "size_t": _C.c_size_t, "double": _C.c_double,
}
_RET = {"int": _C.c_int, "size_t": _C.c_size_t, "const char*": _C.c_char_p, "int64_t": _C.c_int64}


def declared_functions(header_path=HEADER_PATH):
    """Parse `ret name(args);` prototypes out of the public header -> {name: (ret, [argtypes])}."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*(int|int64_t|size_t|const char\*)\s+(hvae_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S | re.M):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argt = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                ty = a.rsplit(" ", 1)[0] if not a.endswith("*") else a
                ty = ty.replace(" *", "*")
                argt.append(ty)
        protos[name] = (ret, argt)
    return protos


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "hvae: %s is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU / eager fallback." % LIB_PATH
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, argt) in declared_functions().items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = _RET[ret]
            fn.argtypes = [_TYPES[t] for t in argt]
        if L.hvae_version() != 100:
            raise RuntimeError("hvae: header/library version mismatch")
        _lib = L
    return _lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = lib().hvae_strerror(code).decode()
        raise RuntimeError("hvae_b200 %s failed: %s (code %d)" % (what, msg, code))


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("hvae ops run only on CUDA (sm_100a) tensors; there is no CPU fallback")
        if t.dtype != torch.float32:
            raise RuntimeError("hvae ops are float32; got %s" % t.dtype)
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        check(lib().hvae_device_check(), "device_check")
        _checked_devices.add(dev)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


launch_count = 0  # number of C-ABI kernel entry calls (bench.py reports it as gpu_launches)


def call(name, *args):
    global launch_count
    launch_count += 1
    check(getattr(lib(), name)(*args), name)
