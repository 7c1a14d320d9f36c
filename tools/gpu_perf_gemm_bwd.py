"""Backward GEMM timings (CUDA events, L2 flushed between launches) against cuBLAS on the training step's shapes:
dX = dY . W (W read in place as the [k, n] operand) and dW += dY^T . X (split-K, fp32 atomics into the sink)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import autograd as ag, ops

DEV = "cuda:0"
dtype = torch.bfloat16
g = torch.Generator(device=DEV).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def run_interleaved(fns, iters=30):
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
    out = {}
    for key, lst in evs.items():
        ts = sorted(a.elapsed_time(b) for a, b in lst[len(lst) // 4:])
        out[key] = ts[len(ts) // 2]
    return out


M = 204800
CAP = int(os.environ.get("CAP", "0"))
ops.set_option("gemm_max_ctas", CAP)
print("gemm_max_ctas", CAP)
for (n_out, k_in, gelu_bwd) in [(2304, 768, False), (768, 768, False), (3072, 768, False), (768, 3072, True)]:
    # forward y[M, n_out] = x[M, k_in] . W[n_out, k_in]^T
    dy = torch.randn(M, n_out, generator=g, device=DEV).to(dtype)
    x = torch.randn(M, k_in, generator=g, device=DEV).to(dtype)
    w = (torch.randn(n_out, k_in, generator=g, device=DEV) * k_in ** -0.5).to(dtype)
    u = torch.randn(M, k_in, generator=g, device=DEV).to(dtype) if gelu_bwd else None
    dx = torch.empty(M, k_in, dtype=dtype, device=DEV)
    sink = torch.zeros(n_out, k_in, dtype=torch.float32, device=DEV)
    bsink = torch.zeros(n_out, dtype=torch.float32, device=DEV)
    fl = 2.0 * M * n_out * k_in
    fns = {
        "dx": lambda: ops.gemm_ex(dy, w, b_mn=True, out=dx, gelu_bwd_aux=u),
        "dx_narrow": lambda: (ops.set_option("gemm_wide", 0), ops.gemm_ex(dy, w, b_mn=True, out=dx, gelu_bwd_aux=u),
                              ops.set_option("gemm_wide", 1)),
        "dx_cublas": lambda: torch.matmul(dy, w, out=dx),
        "dw": lambda: ag._weight_grad(dy, x, sink),
        "dw_cublas": lambda: torch.matmul(dy.t(), x),
        "db": lambda: ag._bias_grad(dy, bsink),
    }
    med = run_interleaved(fns)
    print("n_out=%d k_in=%d gelu_bwd=%d | dX %.3f ms %.0f TF/s (narrow %.0f, cuBLAS plain %.0f) | dW %.3f ms %.0f TF/s (cuBLAS bf16-out %.0f) | "
          "bias grad %.3f ms %.0f GB/s" % (n_out, k_in, gelu_bwd, med["dx"], fl / med["dx"] / 1e9, fl / med["dx_narrow"] / 1e9,
                                           fl / med["dx_cublas"] / 1e9,
                                           med["dw"], fl / med["dw"] / 1e9, fl / med["dw_cublas"] / 1e9, med["db"],
                                           M * n_out * 2 / med["db"] / 1e6), flush=True)
    del dy, x, w, u, dx
