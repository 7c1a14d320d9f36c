"""One launch of each representative block kernel shape (for `ncu --set full` captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import ops

DEV = "cuda:0"
dtype = torch.bfloat16
g = torch.Generator(device=DEV).manual_seed(0)
reps = int(os.environ.get("REPS", "1"))


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g, device=DEV) * scale).to(dtype)


m = 102400
x, x4 = rnd(m, 768), rnd(m, 3072)
w_qkv, w_o, w_1, w_2 = rnd(2304, 768, scale=0.03), rnd(768, 768, scale=0.03), rnd(3072, 768, scale=0.03), rnd(768, 3072, scale=0.02)
b_qkv, b_o, b_1 = rnd(2304), rnd(768), rnd(3072)
for _ in range(reps):
    ops.linear(x, w_qkv, b_qkv)                      # QKV projection
    ops.linear(x, w_o, b_o, residual=x)              # attention output projection + residual
    ops.linear(x, w_1, b_1, gelu=True)               # h -> 4h + GeLU
    ops.linear(x4, w_2, b_o, residual=x)             # 4h -> h + residual
    for (b_, s) in [(400, 256), (200, 512)]:
        qkv = rnd(b_ * s, 2304)
        pad = torch.zeros(b_, s, dtype=torch.bool, device=DEV)
        pad[:, int(s * 0.8):] = True
        ops.attention(qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:], b_, 12, s, s, q_pad=pad, k_pad=pad)
torch.cuda.synchronize()
print("done")
