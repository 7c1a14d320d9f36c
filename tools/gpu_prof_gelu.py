"""One h->4h + GeLU GEMM at the reader's token count, for `ncu --set full --import-source on -k regex:gemm_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import ops
DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
m, n, k = 185600, 3072, 768
x = torch.randn(m, k, generator=g, device=DEV).to(torch.bfloat16)
w = (torch.randn(n, k, generator=g, device=DEV) * k ** -0.5).to(torch.bfloat16)
b = torch.randn(n, generator=g, device=DEV).to(torch.bfloat16)
y = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
ops.set_option("gemm_pair", 0)
for _ in range(3):
    ops.linear(x, w, b, gelu=True, out=y)
torch.cuda.synchronize()
