"""Bandwidth survey of the row operators at the step's sizes (CUDA events, L2 flushed): algorithmic bytes / time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import autograd as ag, dropout, ops
DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def med(fn, iters=24):
    evs = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs[4:])
    return ts[len(ts) // 2]


rows, h = 185600, 768
x = torch.randn(rows, h, device=DEV).to(torch.bfloat16)
r = torch.randn(rows, h, device=DEV).to(torch.bfloat16)
y = torch.empty_like(x)
spec = dropout.DropoutState(5).next(0.1, torch.device(DEV), h)
ms = med(lambda: ops.dropout_add(x, r, spec, out=y))
print("dropout_add  [%d x %d] + residual: %.4f ms %.0f GB/s (3 tensors)" % (rows, h, ms, 3 * x.numel() * 2 / ms / 1e6))
ms = med(lambda: ops.dropout_add(x, None, spec, out=y))
print("dropout (backward form, no residual): %.4f ms %.0f GB/s (2 tensors)" % (ms, 2 * x.numel() * 2 / ms / 1e6))
for n in (768, 3072):
    d = torch.randn(rows, n, device=DEV).to(torch.bfloat16)
    sink = torch.zeros(n, dtype=torch.float32, device=DEV)
    ms = med(lambda: ag._bias_grad(d, sink))
    print("colsum [%d x %d]: %.4f ms %.0f GB/s" % (rows, n, ms, d.numel() * 2 / ms / 1e6))
    del d
ids = torch.randint(1000, 30000, (rows,), device=DEV)
word = torch.randn(30522, h, device=DEV).to(torch.bfloat16)
pos = torch.randn(512, h, device=DEV).to(torch.bfloat16)
ids2 = ids.view(400, 464)
ms = med(lambda: ops.embedding(ids2, word, pos, None, None))
print("embedding forward [%d tokens]: %.4f ms %.0f GB/s (read word rows + write)" % (rows, ms, 2 * rows * h * 2 / ms / 1e6))
logits = torch.randn(12800, 30720, device=DEV).to(torch.bfloat16)
labels = torch.randint(0, 30000, (12800,), device=DEV)
ms = med(lambda: ops.token_logprob(logits.view(8, 50, 32, 30720), labels.view(8, 50, 32)))
print("token_logprob [12800 x 30720]: %.4f ms %.0f GB/s" % (ms, logits.numel() * 2 / ms / 1e6))
