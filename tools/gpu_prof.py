"""Minimal c2 workload for ncu captures: 1M x 768 bf16 evidence, 64 queries, top-50, a few searches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200.mips import ShardSearcher

n = int(os.environ.get("N", 1000000)); d = 768; nq = 64; k = 50
iters = int(os.environ.get("ITERS", 6))
g = torch.Generator(device="cuda").manual_seed(1234)
E = (torch.randn(n, d, generator=g, device="cuda") / d ** 0.5).to(torch.bfloat16)
Q = torch.randn(nq, d, generator=g, device="cuda").to(torch.bfloat16)
s = ShardSearcher(d, torch.bfloat16, "cuda:0")
for name in ("probe", "share", "max_ctas"):
    if name.upper() in os.environ:
        s.set_option(name, int(os.environ[name.upper()]))
s.set_shard(E, None, 1)
torch.cuda.synchronize()
for _ in range(iters):
    s.search(Q, k)
torch.cuda.synchronize()
print("done")
