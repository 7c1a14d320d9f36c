"""One self-attention forward + backward (dropout 0.1) at the reader's training shape — 100 sequences x 512 tokens,
12 heads, bf16 — for `ncu --set full -k regex:attention`; without NCU in the environment it prints CUDA-event times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import autograd as ag, ops
DEV = "cuda:0"
b, heads, s = 100, 12, 512
h = heads * 64
g = torch.Generator(device=DEV).manual_seed(0)
qkv = (torch.randn(b * s, 3 * h, generator=g, device=DEV) * 0.5).to(torch.bfloat16).requires_grad_(True)
lens = torch.randint(380, 513, (b,), generator=torch.Generator().manual_seed(1))
pad = (torch.arange(s)[None, :] >= lens[:, None]).to(DEV)
live = ops.live_blocks(pad.to(torch.uint8))
gout = torch.randn(b * s, h, generator=g, device=DEV).to(torch.bfloat16)
p = float(os.environ.get("P", "0.1"))


def step():
    qkv.grad = None
    out = ag.self_attention(qkv, b, heads, s, pad=pad, live=live, scale=0.125, dropout_p=p)
    out.backward(gout)


for _ in range(2 if os.environ.get("NCU") else 5):
    step()
torch.cuda.synchronize()
if not os.environ.get("NCU"):
    from emdr2_b200 import _lib
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(20):
        step()
    ev[1].record()
    torch.cuda.synchronize()
    fl = 4.0 * heads * 64 * float((lens.double() ** 2).sum())
    ms = ev[0].elapsed_time(ev[1]) / 20
    print("fwd+bwd %.3f ms; algorithmic fwd flops %.3f TF -> %.0f TF/s at 3.5x (fwd + 2.5x bwd)" % (ms, fl / 1e12, 3.5 * fl / ms / 1e9))
