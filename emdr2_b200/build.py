"""In-tree build of libemdr2_b200.so (hand-written sm_100a CUDA behind the C ABI of include/emdr2_b200.h).

Plain ``nvcc -shared``: no torch headers, no pybind; the library is loaded with ctypes
(emdr2_b200/_lib.py).  nvcc cross-compiles sm_100a without a GPU, so this runs in the CPU container.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libemdr2_b200.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libemdr2_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-shared",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "emdr2_b200.h")
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [inc]
    for p in files:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libemdr2_b200.so")
    return p


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + _sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libemdr2_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    with open(STAMP_PATH, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
