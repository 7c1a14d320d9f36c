"""Quick on-GPU sanity run of the MIPS scan against torch (debug aid; the real parity tests live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emdr2_b200.mips import ShardSearcher

def run(n, d, nq, k, dtype, dist, opts=None, timing=False):
    g = torch.Generator(device="cuda").manual_seed(1234)
    if dist == "X":
        E = (torch.randint(-127, 128, (n, d), generator=g, device="cuda").float() / 64).to(dtype)
        Q = (torch.randint(-127, 128, (nq, d), generator=g, device="cuda").float() / 64).to(dtype)
    else:
        E = (torch.randn(n, d, generator=g, device="cuda") / d ** 0.5).to(dtype)
        Q = torch.randn(nq, d, generator=g, device="cuda").to(dtype)
    s = ShardSearcher(d, dtype, "cuda:0")
    s.set_option("stats", 1)
    for kname, v in (opts or {}).items():
        s.set_option(kname, v)
    s.set_shard(E, None, 1)
    sc, ids = s.search(Q, k)
    torch.cuda.synchronize()
    # torch reference in fp32 (chunked), (score desc, id asc) via stable sort on ids
    best_s, best_i = None, None
    for c0 in range(0, n, 1 << 18):
        S = Q.float() @ E[c0:c0 + (1 << 18)].float().T
        idx = torch.arange(c0, c0 + S.shape[1], device="cuda").expand_as(S)
        cs = S if best_s is None else torch.cat([best_s, S], 1)
        ci = idx if best_i is None else torch.cat([best_i, idx], 1)
        o = torch.sort(cs, dim=1, descending=True, stable=True)
        best_s, best_i = o.values[:, :k], torch.gather(ci, 1, o.indices[:, :k])
    ref_i = best_i + 1
    same = (ids == ref_i)
    sdiff = (sc - best_s).abs().max().item()
    print("n=%d d=%d nq=%d k=%d %s %s opts=%s: ids match %.4f (%d/%d) max|ds|=%.3g  ctas=%d stages=%d appends=%d compactions=%d probe_wait_max_ns=%d wait_sum_ns=%d" % (
        n, d, nq, k, str(dtype).split('.')[-1], dist, opts, same.float().mean().item(), same.sum().item(), same.numel(), sdiff,
        s.stat("ctas"), s.stat("stages"), s.stat("appends"), s.stat("compactions"), s.stat("probe_wait_ns"), s.stat("probe_wait_sum_ns")), flush=True)
    if not same.all():
        bad = (~same).nonzero()[:5]
        for q, j in bad.tolist():
            print("   q=%d rank=%d got id %d score %.6f ; ref id %d score %.6f" % (q, j, ids[q, j].item(), sc[q, j].item(), ref_i[q, j].item(), best_s[q, j].item()))
    if timing:
        s.set_option("stats", 0)
        for _ in range(3):
            s.search(Q, k)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        iters = 20
        for _ in range(iters):
            s.search(Q, k)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        gb = n * d * 2 / 1e9
        print("   %.3f ms/search  %.1f GB/s  %.0f q/s" % (ms, gb / ms * 1e3, nq / ms * 1e3), flush=True)
    s.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    run(1000, 128, 32, 5, torch.float16, "X")
    run(128, 64, 64, 5, torch.bfloat16, "X")
    run(5000, 768, 64, 50, torch.bfloat16, "X")
    run(100000, 768, 64, 50, torch.bfloat16, "X", timing=True)
    run(100000, 768, 64, 50, torch.float16, "G")
    run(1000000, 768, 64, 50, torch.bfloat16, "X", timing=True)
    run(1000000, 768, 64, 50, torch.bfloat16, "G", timing=True)
    run(1000000, 768, 64, 50, torch.bfloat16, "G", opts={"probe": 0}, timing=True)
    run(1000000, 768, 64, 50, torch.bfloat16, "G", opts={"share": 0}, timing=True)
