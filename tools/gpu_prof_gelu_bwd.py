"""One dU = (dA . W2) * GeLU'(u) product at the reader's token count, for `ncu --set full -k regex:gemm_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200 import ops
DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)
m, n, k = 204800, 3072, 768
dy = torch.randn(m, k, generator=g, device=DEV).to(torch.bfloat16)
w2 = (torch.randn(k, n, generator=g, device=DEV) * k ** -0.5).to(torch.bfloat16)
u = torch.randn(m, n, generator=g, device=DEV).to(torch.bfloat16)
out = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
for _ in range(3):
    ops.gemm_ex(dy, w2, b_mn=True, out=out, gelu_bwd_aux=u)
torch.cuda.synchronize()
