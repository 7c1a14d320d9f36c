"""Kernel-level timings of the block operators on one B200 (CUDA events, warm-up, L2-sized inputs)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emdr2_b200 import ops

DEV = "cuda:0"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():  # noqa
    dtype = torch.bfloat16
    out = {"gemm": [], "attention": [], "layernorm": [], "encoder": []}
    g = torch.Generator(device=DEV).manual_seed(0)
    for (m, n, k, gelu, res) in [(102400, 2304, 768, False, False), (102400, 768, 768, False, True),
                                 (102400, 3072, 768, True, False), (102400, 768, 3072, False, True),
                                 (204800, 1536, 768, False, False), (12800, 30720, 768, False, False),
                                 (2048, 2304, 768, False, False), (256, 30720, 768, False, False)]:
        x = torch.randn(m, k, generator=g, device=DEV).to(dtype)
        w = (torch.randn(n, k, generator=g, device=DEV) * k ** -0.5).to(dtype)
        b = torch.randn(n, generator=g, device=DEV).to(dtype)
        r = torch.randn(m, n, generator=g, device=DEV).to(dtype) if res else None
        y = torch.empty(m, n, dtype=dtype, device=DEV)
        ms = timeit(lambda: ops.linear(x, w, b, gelu=gelu, residual=r, out=y))
        ms_ref = timeit(lambda: torch.nn.functional.linear(x, w, b))
        fl = 2.0 * m * n * k
        out["gemm"].append(dict(m=m, n=n, k=k, gelu=gelu, residual=res, ms=ms, tflops=fl / ms / 1e9,
                                cublas_ms=ms_ref, cublas_tflops=fl / ms_ref / 1e9))
        print(out["gemm"][-1], flush=True)
        del x, w, r, y
    for (b_, heads, sq, sk, causal) in [(400, 12, 256, 256, False), (400, 12, 512, 512, False),
                                       (8, 12, 32, 25600, False), (400, 12, 32, 512, False),
                                       (400, 12, 32, 32, True)]:
        wdt = heads * 64
        q = torch.randn(b_ * sq, wdt, generator=g, device=DEV).to(dtype)
        k = torch.randn(b_ * sk, wdt, generator=g, device=DEV).to(dtype)
        v = torch.randn(b_ * sk, wdt, generator=g, device=DEV).to(dtype)
        o = torch.empty_like(q)
        pad = torch.zeros(b_, sk, dtype=torch.uint8, device=DEV)
        ms = timeit(lambda: ops.attention(q, k, v, b_, heads, sq, sk, k_pad=pad, causal=causal, out=o))
        fl = 4.0 * b_ * heads * sq * sk * 64
        q4 = q.view(b_, sq, heads, 64).transpose(1, 2)
        k4 = k.view(b_, sk, heads, 64).transpose(1, 2)
        v4 = v.view(b_, sk, heads, 64).transpose(1, 2)
        ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4, is_causal=causal))
        out["attention"].append(dict(batch=b_, heads=heads, sq=sq, sk=sk, causal=causal, ms=ms,
                                     tflops=fl / ms / 1e9, sdpa_ms=ms_ref, sdpa_tflops=fl / ms_ref / 1e9))
        print(out["attention"][-1], flush=True)
        del q, k, v, o
    x = torch.randn(102400, 768, generator=g, device=DEV).to(dtype)
    gam = torch.ones(768, dtype=dtype, device=DEV)
    bet = torch.zeros(768, dtype=dtype, device=DEV)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, gam, bet, out=y))
    out["layernorm"].append(dict(rows=102400, h=768, ms=ms, gbs=2 * x.numel() * 2 / ms / 1e6))
    print(out["layernorm"][-1], flush=True)
    del x, y

    from emdr2_b200.blocks import BertTower, bert_base_config
    cfg = bert_base_config(dtype)
    model = BertTower(cfg).to(DEV).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.normal_(0, 0.02)
    for (b_, s) in [(400, 256), (128, 256), (400, 512), (8, 256)]:
        ids = torch.randint(1, 30000, (b_, s), generator=g, device=DEV)
        ms = timeit(lambda: model(ids, None, None), iters=5, warm=2)
        tokens = b_ * s
        fl = tokens * (12 * (2 * 768 * (2304 + 768 + 3072 + 3072)) + 12 * 4 * s * 768)
        out["encoder"].append(dict(batch=b_, seq=s, ms=ms, tflops=fl / ms / 1e9, tokens_per_s=tokens / ms * 1e3))
        print(out["encoder"][-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/perf_blocks.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
