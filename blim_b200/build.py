"""Build libblim_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The shared object lands next to the sources (blim_b200/csrc/libblim_b200.so) so that it travels to the GPU box with the
repo snapshot.  No torch headers are involved: the boundary is a plain C ABI (include/blim_b200.h) bound with ctypes.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libblim_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")

SOURCES = ["engine.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "blim_b200.h"),
                                                                              os.path.join(ROOT, "include", "blim_vision.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the BLiM B200 engine cannot be built (there is no CPU fallback)")


NVCC_FLAGS += os.environ.get("BLIM_NVCC_EXTRA", "").split()   # e.g. -DBLIM_EXACT_SILU for numerical A/B builds


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    """True when the shared object next to the sources was built from exactly these sources and flags."""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile the engine if the sources changed.  Returns the path of the shared object."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libblim_b200.so")
    if verbose:
        print(log)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
