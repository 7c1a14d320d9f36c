"""Sustained (power-capped) A/B of two library builds: back-to-back launches of the step's four projections for
~1.5 s per segment, alternating builds, no L2 flush — the regime the live step runs in (sw_power_cap, ~1.5 GHz)."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import _lib, ops

DEV = "cuda:0"
dtype = torch.bfloat16
new = _lib.load()
old = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ab", os.environ.get("AB_OLD", "libemdr2_old.so")))
for name, (restype, argtypes) in _lib._SIGNATURES.items():
    if hasattr(old, name):
        fn = getattr(old, name)
        fn.restype, fn.argtypes = restype, argtypes
g = torch.Generator(device=DEV).manual_seed(0)
m, h = 185600, 768
x = torch.randn(m, h, generator=g, device=DEV).to(dtype)
w_qkv = (torch.randn(3 * h, h, generator=g, device=DEV) * h ** -0.5).to(dtype)
w_o = (torch.randn(h, h, generator=g, device=DEV) * h ** -0.5).to(dtype)
w_1 = (torch.randn(4 * h, h, generator=g, device=DEV) * h ** -0.5).to(dtype)
w_2 = (torch.randn(h, 4 * h, generator=g, device=DEV) * (4 * h) ** -0.5).to(dtype)
b3, b1, b4 = (torch.zeros(n, dtype=dtype, device=DEV) for n in (3 * h, h, 4 * h))
qkv = torch.empty(m, 3 * h, dtype=dtype, device=DEV)
a1 = torch.empty(m, h, dtype=dtype, device=DEV)
u = torch.empty(m, 4 * h, dtype=dtype, device=DEV)
a2 = torch.empty(m, h, dtype=dtype, device=DEV)
flops = 2.0 * m * h * h * (3 + 1 + 4 + 4)


def layer():
    ops.linear(x, w_qkv, b3, out=qkv)
    ops.linear(qkv[:, :h], w_o, b1, residual=x, out=a1)
    ops.linear(a1, w_1, b4, gelu=True, out=u)
    ops.linear(u, w_2, b1, residual=a1, out=a2)


def segment(lib, seconds=1.5):
    _lib._LIB = lib
    for _ in range(3):
        layer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            layer()
        n += 10
        torch.cuda.current_stream().synchronize() if n % 200 == 0 else None
    e1.record()
    torch.cuda.synchronize()
    _lib._LIB = new
    return flops * n / e0.elapsed_time(e1) / 1e9


res = {"old": [], "new": []}
for rnd in range(4):
    for tag, lib in (("old", old), ("new", new)):
        res[tag].append(segment(lib))
for tag in res:
    print(tag, " ".join("%.0f" % v for v in res[tag]), "TF/s  mean of last 3: %.0f" % (sum(res[tag][1:]) / 3))
