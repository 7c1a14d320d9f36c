"""A/B timing of two builds of the library on the step's GEMM shapes, interleaved in ONE process so clock and
power state are shared: tools/_ab/libemdr2_old.so (a build of another commit) against the in-tree library."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import _lib, ops

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

DEV = "cuda:0"
dtype = torch.bfloat16
new = _lib.load()
old = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ab", os.environ.get("AB_OLD", "libemdr2_old.so")))
for name, (restype, argtypes) in _lib._SIGNATURES.items():
    if hasattr(old, name):
        fn = getattr(old, name)
        fn.restype, fn.argtypes = restype, argtypes
LIBS = {"old": old, "new": new}
g = torch.Generator(device=DEV).manual_seed(0)
SHAPES = [(185600, 768, 3072, False, False), (185600, 2304, 768, False, True), (185600, 2304, 768, False, False), (185600, 768, 768, False, True), (185600, 3072, 768, True, False),
          (185600, 768, 3072, False, True), (204800, 1536, 768, False, False), (66000, 2304, 768, False, False),
          (66000, 3072, 768, True, False), (12800, 30720, 768, False, False)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def run_interleaved(fns, iters):
    """Every variant once per iteration, in rotating order, L2 flushed before each launch; median per variant."""
    evs = {k: [] for k in fns}
    keys = list(fns)
    for i in range(iters):
        for j in range(len(keys)):
            key = keys[(i + j) % len(keys)]
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fns[key]()
            b.record()
            evs[key].append((a, b))
    torch.cuda.synchronize()
    out = {}
    for key, lst in evs.items():
        ts = sorted(a.elapsed_time(b) for a, b in lst[len(lst) // 4:])
        out[key] = ts[len(ts) // 2]
    return out


for (m, n, k, gelu, res) in SHAPES:
    x = torch.randn(m, k, generator=g, device=DEV).to(dtype)
    w = (torch.randn(n, k, generator=g, device=DEV) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g, device=DEV).to(dtype)
    r = torch.randn(m, n, generator=g, device=DEV).to(dtype) if res else None
    y = torch.empty(m, n, dtype=dtype, device=DEV)
    fl = 2.0 * m * n * k
    def variant(lib, mode, wide):
        def fn():
            _lib._LIB = lib
            ops.set_option("gemm_pair", mode)
            if lib is new:
                ops.set_option("gemm_wide", wide)
            ops.linear(x, w, b, gelu=gelu, residual=r, out=y)
        return fn

    fns = {"old": variant(old, 0, 0), "narrow": variant(new, 0, 0), "wide": variant(new, 0, 2), "pair": variant(new, 1, 0)}
    fns["ref"] = lambda: torch.nn.functional.linear(x, w, b)
    med = run_interleaved(fns, 60)
    _lib._LIB = new
    print("m=%d n=%d k=%d gelu=%d res=%d | one-CTA old %.0f narrow %.0f wide %.0f | pair %.0f | cuBLAS bias-only %.0f TF/s"
          % (m, n, k, gelu, res, fl / med["old"] / 1e9, fl / med["narrow"] / 1e9, fl / med["wide"] / 1e9,
             fl / med["pair"] / 1e9, fl / med["ref"] / 1e9), flush=True)
    del x, w, r, y
