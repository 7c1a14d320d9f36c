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
        h.update(os.path.basename(p).encode())      # names, not absolute paths: the digest must survive a move
        with open(p, "rb") as f:                     # of the tree (the GPU box runs a snapshot elsewhere)
            h.update(f.read())
    return h.hexdigest()


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libemdr2_b200.so")
    return p


OBJ_DIR = os.path.join(PKG_DIR, "build")


def _compile_objects(force, verbose):
    """One object per .cu, compiled in parallel and only when that source, a header or the flags
    changed (build/<name>.o + .stamp) — editing one kernel recompiles one file."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "emdr2_b200.h")
    for p in sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")) + [inc]:
        with open(p, "rb") as f:
            hdr.update(os.path.basename(p).encode() + f.read())
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def one(src):
        name = os.path.splitext(os.path.basename(src))[0]
        obj, stamp = os.path.join(OBJ_DIR, name + ".o"), os.path.join(OBJ_DIR, name + ".stamp")
        h = hdr.copy()
        with open(src, "rb") as f:
            h.update(f.read())
        digest = h.hexdigest()
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
            return obj, ""
        cmd = [nvcc_path()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (os.path.basename(src), res.stdout + res.stderr))
        with open(stamp, "w") as f:
            f.write(digest)
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as ex:
        results = list(ex.map(one, _sources()))
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    return [obj for obj, _ in results]


def _up_to_date(digest):
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == digest


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library (if the sources changed since the last
    build: the stamp holds a digest of csrc/, the header and the flags).  Returns its path.

    Safe under N ranks importing at once on a fresh checkout: one process compiles under an exclusive
    file lock into a temporary file that is renamed over the library; the others wait on the lock and
    find the stamp current."""
    import fcntl
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB_PATH
    with open(os.path.join(PKG_DIR, ".libemdr2_b200.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):       # another rank built it while we waited
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            objects = _compile_objects(force, verbose)
            cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objects
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed linking libemdr2_b200.so")
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + ".tmp", "w") as f:
                f.write(digest)
            os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
