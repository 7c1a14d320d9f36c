"""Time (CUDA events) and, under ncu, profile the varlen attention kernel on the reader's shape: 400 packed
sequences of ~450 tokens (NQ-shaped extended contexts), 12 heads, bf16; also the dense 400 x 512 case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emdr2_b200 import ops
from emdr2_b200.packed import PackedBatch

DEV = torch.device("cuda:0")
heads, h = 12, 768
out = {}
for name, lens in (("reader_400x~450", np.random.RandomState(0).randint(380, 513, size=400)),
                   ("dense_400x512", np.full(400, 512)),
                   ("context_400x~150", np.random.RandomState(1).randint(105, 192, size=400))):
    pb = PackedBatch(lens, 512, heads, DEV)
    qkv = (torch.randn(pb.T, 3 * h, device=DEV) * 0.5).to(torch.bfloat16)
    o = torch.empty(pb.T, h, dtype=torch.bfloat16, device=DEV)
    run = lambda: ops.attention_varlen(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], heads, pb.items, pb.n_items,
                                       scale=0.125, out=o)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    iters = 1 if os.environ.get("NCU") else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out[name] = dict(ms=ms, tokens=pb.T, items=pb.n_items, tflops=pb.attention_flops / ms / 1e9)
print(json.dumps(out))
