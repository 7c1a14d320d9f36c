#!/usr/bin/env python
"""Benchmark of the retrieve hot path: brute-force MIPS top-k of a query batch over the evidence matrix.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU (FAISS-flat) path
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, NCCL

Workload (BASELINE.json `metric`: "queries/sec ... @21M docs, 1/2/4/8 GPU"): 21 000 000 x 768 fp16
synthetic evidence, batch of 64 queries, top-50 — configs[2].  The corpus is FIXED and row-sharded
over the N ranks with the reference's torch.chunk rule (strong scaling); it fits a single B200
(32.3 GB of 180 GB), so N=1 runs the same corpus.  One "step" = one search of the 64-query batch:
fused GEMM+top-k scan of the local shard, pool merge, and for N>1 one all-gather of [64,50]
(score,id) pairs + k-way merge on every rank.

The one JSON line printed by rank 0 carries:
  value        queries/s, evidence AND queries resident in HBM, CUDA-event timed, max over ranks
  e2e          same, through B200BruteForceIndex.search_mips_index with pinned HOST query buffers
               copied in and fp16 scores / int32 ids copied back out inside the timed region
  roofline     the scan kernel: algorithmic bytes per launch / its mean launch duration (CUDA events
               recorded around every scan launch on the launching stream by the library), against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline the reference's CPU path (FAISS IndexFlatIP algorithm, oracle/flat_ip.py: blocked
               SGEMM + running top-k) timed on this box's host cores over a bounded row sample and
               scaled linearly to the full corpus (rank 0, N=1 only)
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--rows", type=int, default=21_000_000, help="total evidence rows (all ranks)")
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nq", type=int, default=64)
    p.add_argument("--k", type=int, default=50)
    p.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    p.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gpu-reference", action="store_true")
    return p.parse_args()


def workload_name(a):
    return "%d x %d %s evidence, %d queries, top-%d" % (a.rows, a.dim, a.dtype, a.nq, a.k)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(rows_per_gpu, dim):
    """DRAM bytes per scan launch from the committed `ncu --set full` capture of this workload
    (profiles/roofline_traffic.json), or None when no capture of this shard size exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            table = json.load(f)
        return table.get("%dx%d" % (rows_per_gpu, dim), {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU through NVML while a region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period=0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._period = period
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self._period)

    def __enter__(self):
        if self._h is not None:
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._h is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_step_fn(a):
    """Returns (step, sample_description, cores): step() runs one search of the nq-query batch over
    the bounded row sample with the FAISS-flat port and returns its wall seconds."""
    import torch
    from oracle.flat_ip import flat_ip_search, host_threads
    cores = host_threads()
    torch.set_num_threads(cores)
    s_rows = min(a.cpu_sample_rows, a.rows)
    g = torch.Generator().manual_seed(1234)
    rows = torch.empty(s_rows, a.dim)
    for r0 in range(0, s_rows, 1 << 17):               # generated in slices: bounded temporaries
        r1 = min(s_rows, r0 + (1 << 17))
        rows[r0:r1] = (torch.randn(r1 - r0, a.dim, generator=g) / a.dim ** 0.5).half().float()
    queries = torch.randn(a.nq, a.dim, generator=g).half().float()

    def step():
        t0 = time.perf_counter()
        flat_ip_search(rows, None, queries, a.k)
        return time.perf_counter() - t0

    sample = ("%d of %d rows (fp32 copies of fp16 values), %d queries, top-%d; time scaled x%.3f "
              "to the full corpus" % (s_rows, a.rows, a.nq, a.k, a.rows / s_rows))
    return step, sample, cores, a.rows / s_rows


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    step, sample, cores, scale = cpu_reference_step_fn(a)
    for _ in range(max(1, min(a.warmup, 2))):
        step()
    steps = max(1, a.steps)
    budget_s, times = 150.0, []
    for _ in range(steps):                       # bounded: stop early rather than run for hours
        times.append(step())
        if sum(times) > budget_s:
            break
    sec = statistics.mean(times) * scale
    value = a.nq / sec
    line = {
        "impl": "reference", "metric": "queries/sec retrieve (brute-force MIPS top-%d) @%.0fM docs" % (a.k, a.rows / 1e6),
        "value": value, "unit": "queries/s", "n_gpus": a.gpus, "steps": len(times), "warmup": a.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "path": "FaissMIPSIndex CPU (IndexFlatIP) restated by oracle/flat_ip.py; faiss itself is absent from the image"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------- B200 arm
def make_shard(torch, n_local, dim, dtype, seed, device):
    """Gaussian evidence rows generated on the device in 1M-row slices (never on the host)."""
    g = torch.Generator(device=device).manual_seed(seed)
    rows = torch.empty((n_local, dim), dtype=dtype, device=device)
    for r0 in range(0, n_local, 1 << 20):
        r1 = min(n_local, r0 + (1 << 20))
        rows[r0:r1] = (torch.randn(r1 - r0, dim, generator=g, device=device) / dim ** 0.5).to(dtype)
    return rows


def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    from emdr2_b200.index import B200BruteForceIndex, chunk_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (a.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: emdr2_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    tdtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype]

    lo, hi = chunk_range(a.rows, world, rank)
    rows = make_shard(torch, hi - lo, a.dim, tdtype, 1234 + rank, device)
    index = B200BruteForceIndex(a.dim, dtype=tdtype, device=device,
                                group=dist.group.WORLD if world > 1 else None)
    index.add_local_shard(None, rows, num_rows=a.rows, row_lo=lo)       # ids = 1-based row numbers
    searcher = index._searcher

    gq = torch.Generator(device=device).manual_seed(99)                 # same queries on all ranks
    n_batches = 8
    q_dev = [torch.randn(a.nq, a.dim, generator=gq, device=device).to(tdtype) for _ in range(n_batches)]
    q_host = [q.cpu().pin_memory() for q in q_dev]
    out_d = torch.empty((a.nq, a.k), dtype=torch.float16).pin_memory()
    out_i = torch.empty((a.nq, a.k), dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_step(i):
        return index.search(q_dev[i % n_batches], a.k)

    def e2e_step(i):
        q = q_host[i % n_batches].to(device, non_blocking=True)
        d, ix = index.search_mips_index(q, a.k, reconstruct=False)
        out_d.copy_(d, non_blocking=True)
        out_i.copy_(ix, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller consumes the ids on the host

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    warm = max(3, a.warmup)
    for i in range(warm):
        resident_step(i)
        e2e_step(i)
    torch.cuda.synchronize()

    with ClockSampler(local_rank) as clocks:
        searcher.set_option("timing", 1)
        ms_total = timed(resident_step, a.steps)
        scan_launches = searcher.stat("scan_launches")
        scan_ns = searcher.stat("scan_ns")
        searcher.set_option("timing", 0)
        ms_e2e = timed(e2e_step, a.steps)
    ms_step = ms_total / a.steps
    value = a.nq / (ms_step * 1e-3)
    e2e_value = a.nq / (ms_e2e / a.steps * 1e-3)

    n_local = hi - lo
    algo_bytes = n_local * a.dim * 2 + a.nq * a.dim * 2 + a.nq * a.k * 12
    scan_ms = max_over_ranks(scan_ns / max(1, scan_launches) * 1e-6)
    peak, peak_src = measured_peak()
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    launches_per_step = (2 if world == 1 else 3) * (-(-a.nq // 64))

    line = {
        "metric": "queries/sec retrieve (brute-force MIPS top-%d) @%.0fM docs" % (a.k, a.rows / 1e6),
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": a.dtype + " inputs, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
        "config": {
            "workload": workload_name(a),
            "rows_per_gpu": n_local, "sharding": "torch.chunk row ranges, one rank per GPU",
            "exchange": "none" if world == 1 else "one all-gather of [nq,k] (fp32 score,int64 id) + k-way merge per rank",
            "l2": "inputs larger than L2 (%.2f GB evidence per GPU streamed per step vs 126 MB L2, evict-first)" % (n_local * a.dim * 2 / 1e9),
            "stage": "retrieve only (BERT encoders and the T5 reader are not in this timed region)",
        },
        "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.dim * 2,
                "d2h_bytes_per_step": a.nq * a.k * (2 + 4),
                "api": "B200BruteForceIndex.search_mips_index (pinned host queries in, fp16 scores + int32 ids out)"},
        "gpu_launches": launches_per_step * a.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": recorded_traffic(n_local, a.dim),
                     "kernel": "emdr2::mips_scan_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                     "kernel_ms": scan_ms, "launches_timed": scan_launches, "peak_source": peak_src},
        "clocks": clocks.summary(),
    }

    if world == 1 and not a.no_gpu_reference:
        line["gpu_reference"] = gpu_reference_leg(torch, rows, q_dev, a)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        step, sample, cores, scale = cpu_reference_step_fn(a)
        step()
        times = []
        while len(times) < 5 and sum(times) < 20.0:
            times.append(step())
        sec = min(times) * scale
        line["cpu_baseline"] = {"value": a.nq / sec, "unit": "queries/s", "cores": cores,
                                "kind": "port", "sample": sample + "; best of %d" % len(times)}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def gpu_reference_leg(torch, rows, q_dev, a):
    """What the reference's training loop executes on a GPU (DistributedBruteForceIndex,
    megatron/data/emdr2_index.py:281-295, restated with library calls): matmul into a
    C[nq, N] score matrix in the input dtype, then torch.topk.  Reported for context only."""
    try:
        def step(i):
            c = torch.matmul(q_dev[i % len(q_dev)], rows.T)
            return torch.topk(c, a.k, dim=1)
        for i in range(2):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 5
        for i in range(iters):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"value": a.nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
                "what": "torch.matmul (cuBLAS) -> C[nq,N] in HBM -> torch.topk, same GPU, same shard"}
    except Exception as exc:        # e.g. out of memory for the C matrix on a small GPU
        return {"unavailable": str(exc).splitlines()[0][:200]}


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_b200_arm(a)


if __name__ == "__main__":
    sys.exit(main())
