"""LayerNorm forward / backward timings at the step's row counts (CUDA events, L2 flushed), GB/s of algorithmic traffic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import autograd as ag, ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def med(fn, iters=30):
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs[5:])
    return ts[len(ts) // 2]


import ctypes
from emdr2_b200 import _lib
new = _lib.load()
old_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ab", os.environ.get("AB_OLD", "libemdr2_old.so"))
old = None
if os.path.exists(old_path):      # optional A/B against another build (tools/make_ab_lib.sh)
    old = ctypes.CDLL(old_path)
    for name, (restype, argtypes) in _lib._SIGNATURES.items():
        if hasattr(old, name):
            fn = getattr(old, name)
            fn.restype, fn.argtypes = restype, argtypes


def with_lib(lib, fn):
    def run():
        _lib._LIB = lib
        fn()
        _lib._LIB = new
    return run


for rows in (185600, 66000, 256):
    x = torch.randn(rows, 768, device=DEV).to(torch.bfloat16)
    g = torch.randn(768, device=DEV).to(torch.bfloat16)
    b = torch.randn(768, device=DEV).to(torch.bfloat16)
    y = torch.empty_like(x)
    ms = med(lambda: ops.layernorm(x, g, b, out=y))
    if old is not None:
        ms_old = med(with_lib(old, lambda: ops.layernorm(x, g, b, out=y)))
        ms = med(lambda: ops.layernorm(x, g, b, out=y))
        print("rows=%d other build %.4f ms %.0f GB/s" % (rows, ms_old, 2 * x.numel() * 2 / ms_old / 1e6))
    ref = torch.nn.functional.layer_norm(x.float(), (768,), g.float(), b.float(), 1e-5)
    err = (ops.layernorm(x, g, b).float() - ref).abs().max().item()
    print("rows=%d forward %.4f ms %.0f GB/s (max|err| vs fp32 torch %.3e)" % (rows, ms, 2 * x.numel() * 2 / ms / 1e6, err))

for rows in (185600, 66000):
    x = torch.randn(rows, 768, device=DEV).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(768, device=DEV).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(768, device=DEV).to(torch.bfloat16).requires_grad_(True)
    dy = torch.randn(rows, 768, device=DEV).to(torch.bfloat16)

    def bwd():
        x.grad = g.grad = b.grad = None
        y = ag.layernorm(x, g, b)
        y.backward(dy)

    ms_fb = med(bwd)
    ms_f = med(lambda: ops.layernorm(x.detach(), g.detach(), b.detach(), return_stats=True))
    ms_b = ms_fb - ms_f
    print("rows=%d backward ~%.4f ms (forward+backward %.4f - forward %.4f) %.0f GB/s of 3 x rows x h x 2 B"
          % (rows, ms_b, ms_fb, ms_f, 3 * x.numel() * 2 / ms_b / 1e6))
