"""In-tree build of libhvae_b200.so with nvcc for sm_100a (no torch headers: the library is a plain
C ABI).  Called by __graft_entry__.build(); the .so is git-ignored but travels with gpurun."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libhvae_b200.so")
OBJ_DIR = os.path.join(PKG_ROOT, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src):
    h = hashlib.sha256()
    for p in [src] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))):
        h.update(open(p, "rb").read())
    h.update(open(os.path.join(os.path.dirname(PKG_ROOT), "include", "hvae_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(verbose=False, force=False, jobs=None, experiment=False):
    """experiment=True builds libhvae_b200_exp.so with -DHVAE_EXPERIMENT (debug knobs + probe entry points of the
    tensor-core kernels; never loaded by the package unless HVAE_LIB_PATH points at it)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs, procs = [], []
    global NVCC_FLAGS, LIB_PATH
    flags0, lib0 = NVCC_FLAGS, LIB_PATH
    if experiment:
        NVCC_FLAGS = NVCC_FLAGS + ["-DHVAE_EXPERIMENT"]
        LIB_PATH = os.path.join(LIB_DIR, "libhvae_b200_exp.so")
    try:
        return _build(nvcc, verbose, force, "exp_" if experiment else "")
    finally:
        NVCC_FLAGS, LIB_PATH = flags0, lib0


def _build(nvcc, verbose, force, prefix):
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, prefix + os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp(src)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, stamp_file, stamp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, stamp_file, stamp, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(out.decode())
        else:
            if verbose:
                sys.stderr.write(out.decode())
            open(stamp_file, "w").write(stamp)
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or force or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, experiment="--exp" in sys.argv))
