"""A/B of two library builds on the attention kernels, interleaved in one process (see gpu_ab_gemm.py)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emdr2_b200 import _lib, autograd as ag, ops
from emdr2_b200.packed import PackedBatch

DEV = torch.device("cuda:0")
new = _lib.load()
old = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ab", os.environ.get("AB_OLD", "libemdr2_old.so")))
for name, (restype, argtypes) in _lib._SIGNATURES.items():
    if hasattr(old, name):
        fn = getattr(old, name)
        fn.restype, fn.argtypes = restype, argtypes
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
heads, h = 12, 768


def run_interleaved(fns, iters=40):
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
    return {k: sorted(a.elapsed_time(b) for a, b in v[len(v) // 4:])[len(v) * 3 // 8] for k, v in evs.items()}


def with_lib(lib, fn):
    def run():
        _lib._LIB = lib
        fn()
        _lib._LIB = new
    return run


for name, lens in (("reader_400x~450", np.random.RandomState(0).randint(380, 513, size=400)),
                   ("context_400x~150", np.random.RandomState(1).randint(105, 192, size=400))):
    pb = PackedBatch(lens, 512, heads, DEV)
    qkv = (torch.randn(pb.T, 3 * h, device=DEV) * 0.5).to(torch.bfloat16)
    o = torch.empty(pb.T, h, dtype=torch.bfloat16, device=DEV)
    fn = lambda: ops.attention_varlen(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], heads, pb.items, pb.n_items, scale=0.125, out=o)
    med = run_interleaved({"old": with_lib(old, fn), "new": with_lib(new, fn)})
    print("varlen %s: old %.4f ms  new %.4f ms" % (name, med["old"], med["new"]), flush=True)

b, s = 100, 512
g = torch.Generator(device=DEV).manual_seed(0)
qkv = (torch.randn(b * s, 3 * h, generator=g, device=DEV) * 0.5).to(torch.bfloat16).requires_grad_(True)
lens = torch.randint(380, 513, (b,), generator=torch.Generator().manual_seed(1))
pad = (torch.arange(s)[None, :] >= lens[:, None]).to(DEV)
live = ops.live_blocks(pad.to(torch.uint8))
gout = torch.randn(b * s, h, generator=g, device=DEV).to(torch.bfloat16)
for p in (0.1, 0.0):
    def fwd():
        with torch.no_grad():
            ag.self_attention(qkv.detach(), b, heads, s, pad=pad, live=live, scale=0.125, dropout_p=p)

    def fwdbwd():
        qkv.grad = None
        ag.self_attention(qkv, b, heads, s, pad=pad, live=live, scale=0.125, dropout_p=p).backward(gout)

    med = run_interleaved({"old_f": with_lib(old, fwd), "new_f": with_lib(new, fwd), "old_fb": with_lib(old, fwdbwd),
                           "new_fb": with_lib(new, fwdbwd)})
    print("dense 100x512 p=%.1f: forward old %.4f new %.4f ms | forward+backward old %.4f new %.4f ms"
          % (p, med["old_f"], med["new_f"], med["old_fb"], med["new_fb"]), flush=True)
