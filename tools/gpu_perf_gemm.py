"""GEMM timings, one-CTA kernel vs CTA-pair kernel vs cuBLAS, on the step's large shapes (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import ops

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_perf_blocks import timeit

DEV = "cuda:0"
dtype = torch.bfloat16
g = torch.Generator(device=DEV).manual_seed(0)
SHAPES = [(185600, 2304, 768, False, False), (185600, 768, 768, False, True), (185600, 3072, 768, True, False),
          (185600, 768, 3072, False, True), (204800, 1536, 768, False, False), (66000, 2304, 768, False, False),
          (66000, 3072, 768, True, False), (12800, 30720, 768, False, False)]
for (m, n, k, gelu, res) in SHAPES:
    x = torch.randn(m, k, generator=g, device=DEV).to(dtype)
    w = (torch.randn(n, k, generator=g, device=DEV) * k ** -0.5).to(dtype)
    b = torch.randn(n, generator=g, device=DEV).to(dtype)
    r = torch.randn(m, n, generator=g, device=DEV).to(dtype) if res else None
    y = torch.empty(m, n, dtype=dtype, device=DEV)
    fl = 2.0 * m * n * k
    res_ms = {}
    for mode in (0, 1):
        ops.set_option("gemm_pair", mode)
        res_ms[mode] = timeit(lambda: ops.linear(x, w, b, gelu=gelu, residual=r, out=y), iters=10)
    ref = timeit(lambda: torch.nn.functional.linear(x, w, b), iters=10)
    print("m=%d n=%d k=%d gelu=%d res=%d: one-CTA %.3f ms %.0f TF/s | pair %.3f ms %.0f TF/s | cuBLAS (bias only) %.3f ms %.0f TF/s"
          % (m, n, k, gelu, res, res_ms[0], fl / res_ms[0] / 1e9, res_ms[1], fl / res_ms[1] / 1e9, ref, fl / ref / 1e9), flush=True)
    del x, w, r, y
