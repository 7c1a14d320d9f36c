"""Attention-only timings (plain, padded, skip) for quick A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf_blocks import timeit
DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
dtype = torch.bfloat16
CASES = [(400, 12, 256, 256, 1.0), (400, 12, 512, 512, 1.0), (400, 12, 256, 256, 0.7), (400, 12, 512, 512, 0.45),
         (8, 12, 32, 25600, 0.45), (400, 12, 32, 512, 0.45)]
if len(sys.argv) > 1:      # e.g. `gpu_perf_attn.py 1 3`: only those cases (for ncu captures)
    CASES = [CASES[int(i)] for i in sys.argv[1:]]
ITERS = int(os.environ.get("ATTN_ITERS", "20"))
for (b_, heads, sq, sk, frac) in CASES:
    w = heads * 64
    q = torch.randn(b_ * sq, w, generator=g, device=DEV).to(dtype)
    k = torch.randn(b_ * sk, w, generator=g, device=DEV).to(dtype)
    v = torch.randn(b_ * sk, w, generator=g, device=DEV).to(dtype)
    o = torch.empty_like(q)
    if sk == 25600:
        kpad = (torch.arange(sk, device=DEV)[None] % 512 >= int(512 * frac)).expand(b_, sk).contiguous()
        qpad = torch.zeros(b_, sq, dtype=torch.bool, device=DEV)
    else:
        kpad = (torch.arange(sk, device=DEV)[None] >= int(sk * frac)).expand(b_, sk).contiguous()
        qpad = kpad if sq == sk else torch.zeros(b_, sq, dtype=torch.bool, device=DEV)
    ms = timeit(lambda: ops.attention(q, k, v, b_, heads, sq, sk, q_pad=qpad, k_pad=kpad, out=o), iters=ITERS)
    ql, kl = ops.live_blocks(qpad), ops.live_blocks(kpad)
    ms2 = timeit(lambda: ops.attention(q, k, v, b_, heads, sq, sk, q_pad=qpad, k_pad=kpad, out=o, q_live=ql, k_live=kl), iters=ITERS)
    print("b=%d sq=%d sk=%d live=%.2f: exact %.3f ms, skip %.3f ms" % (b_, sq, sk, frac, ms, ms2), flush=True)
