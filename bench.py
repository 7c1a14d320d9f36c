#!/usr/bin/env python
"""Benchmark of the retrieve-and-read hot path (BASELINE.json metric: "queries/sec retrieve+read @21M
docs, 1/2/4/8 GPU; MIPS HBM GB/s vs peak").

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, NCCL

One step = EMDR2Model.forward (reference megatron/model/emdr2_model.py:87-214, eval path) for a
batch of `--batch` questions per GPU (8, the reference recipe): query tower (BERT-base, S=256) ->
all-gather of the queries -> brute-force top-50 MIPS over the 21 000 000 x 768 fp16 evidence matrix
row-sharded over the N ranks (one all-gather of [nq,k] pairs + merge) -> passage lookup + formatting
on the host -> context tower over 50 passages per question (S=256) -> fresh scores -> T5-base encoder
over the 50 (question, passage) pairs (S=512) -> FiD decoder (L=32, cross-attention over 25 600
keys) + LM head.  The default run times the forward path (what BASELINE's retrieve+read metric names);
`--train` times the whole training step (forward incl. the no-grad one-context pass, both losses, backward,
gradient all-reduce, optimizer).
Synthetic NQ-shaped data: random-init weights N(0, 0.02), question length
U[8,24], passages U[100,180] tokens, titles U[2,8], answers U[2,6].

Scaling is "weak": every rank reads 8 questions (global batch 8N) over a FIXED 21 M-row corpus, so
per-GPU reader work is constant and the per-GPU MIPS shard shrinks as 1/N.

JSON line (rank 0): value = questions/s, CUDA-event timed, max over ranks, inputs resident on the
device; e2e = same through pinned HOST input buffers with the predicted token ids and passage
log-probabilities copied back; roofline = the dominant kernel (emdr2::gemm_kernel, tensor-bound:
sum of 2mnk over its launches / sum of their CUDA-event durations vs the measured sustained cuBLAS
bf16 rate); roofline_mips = emdr2::mips_scan_kernel (HBM-bound: algorithmic bytes / launch duration
vs the measured copy bandwidth) — the "MIPS HBM GB/s vs peak" half of the metric; cpu_baseline = the
reference's CPU path (FAISS-flat port + fp32 PyTorch reader restatement) on a bounded sample.
`--retrieve-only` benches the search alone with a 64-question batch (BASELINE configs[1]/[2]).
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

NCCL_CTAS = int(os.environ.get("EMDR2_NCCL_CTAS", "4"))   # SMs left to NCCL while gradient all-reduces overlap the backward pass
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=None, help="default 20 (retrieve+read) / 200 (--retrieve-only)")
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--rows", type=int, default=21_000_000, help="total evidence rows (all ranks)")
    p.add_argument("--dim", type=int, default=768)
    p.add_argument("--nq", type=int, default=64, help="questions per search with --retrieve-only")
    p.add_argument("--k", type=int, default=50)
    p.add_argument("--batch", type=int, default=8, help="questions per GPU per step (retrieve+read)")
    p.add_argument("--seq-ret", type=int, default=256)
    p.add_argument("--seq", type=int, default=512)
    p.add_argument("--dec", type=int, default=32)
    p.add_argument("--layers", type=int, default=12)
    p.add_argument("--retrieve-only", action="store_true")
    p.add_argument("--train", action="store_true",
                   help="time the full EMDR2 training step: forward (incl. the no-grad one-context pass), "
                        "losses, backward, gradient all-reduce and a fused AdamW update on fp32 masters")
    p.add_argument("--model-dtype", default="bf16", choices=["fp16", "bf16"])
    p.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    p.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-gpu-reference", action="store_true")
    p.add_argument("--no-train-step", action="store_true", help="skip the train_step block of the default line")
    p.add_argument("--train-steps", type=int, default=8, help="timed training steps of the train_step block")
    p.add_argument("--no-dropout", action="store_true", help="train with hidden/attention dropout 0 (A/B only)")
    p.add_argument("--refresh-rows", type=int, default=32768,
                   help="passages each GPU re-encodes in the index_refresh block (BASELINE config 5, bounded sample); 0 = skip")
    a = p.parse_args()
    if a.steps is None:
        a.steps = 200 if a.retrieve_only else 20
    return a


def workload_name(a):
    if a.retrieve_only:
        return "%d x %d %s evidence, %d queries, top-%d" % (a.rows, a.dim, a.dtype, a.nq, a.k)
    return ("%d x %d %s evidence; %d questions/GPU; top-%d; BERT-base towers S=%d; T5-base-shaped reader "
            "S=%d L=%d, %d layers" % (a.rows, a.dim, a.dtype, a.batch, a.k, a.seq_ret, a.seq, a.dec, a.layers))


def measured_peak(key="hbm_gbs"):
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)[key]), "measured (MEASURED_PEAKS.json %s)" % key
    except Exception:
        fallback = {"hbm_gbs": FALLBACK_HBM_GBS, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}[key]
        return fallback, "fallback (B200_PROFILING.md)"


def recorded_gemm_traffic():
    """Mean DRAM bytes per launch of emdr2::gemm_kernel over the layer's four projections, from the
    committed `ncu --set full` capture (profiles/roofline_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get("gemm_kernel", {}).get("dram_bytes_per_launch")
    except Exception:
        return None


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


# ------------------------------------------------------------------------------- synthetic corpus
def synthetic_token_store(lo, hi, seed, pool=65536):
    """passages_map / title_map stand-in as a flat token store (emdr2_b200/tokens.py — the layout of the
    reference's memory-mapped indexed datasets): x[doc_id - 1] -> int64 token array.  A pool of
    pre-generated documents indexed modulo its size (21 M real passages would be 25 GB of tokens)."""
    import numpy as np
    from emdr2_b200.tokens import FlatTokenStore

    class ModuloTokenStore(FlatTokenStore):
        def __getitem__(self, i):
            return FlatTokenStore.__getitem__(self, i % len(self))

        def spans(self, index):
            index = np.asarray(index, dtype=np.int64)
            return FlatTokenStore.spans(self, np.where(index >= 0, index % len(self), -1))

    rng = np.random.RandomState(seed)
    lens = rng.randint(lo, hi + 1, size=pool)
    offsets = np.zeros(pool + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return ModuloTokenStore(rng.randint(1000, 30000, size=int(offsets[-1])).astype(np.int64), offsets)


class SyntheticTitleMap(object):
    """Articles of 4 consecutive passages (tools/inverted_title_index.py:22-37 semantics), answering one
    doc id (`get_neighbour_paragraphs`) or a whole batch (`lookup`, like titlemap.NeighbourTable)."""

    def __init__(self, num_docs, per_article=4):
        self.n, self.per = num_docs, per_article

    def get_neighbour_paragraphs(self, doc_id):
        first = ((doc_id - 1) // self.per) * self.per + 1
        docs = list(range(first, min(self.n, first + self.per - 1) + 1))
        i = docs.index(doc_id)
        if i == 0:
            return docs[0:3], 0
        if i == len(docs) - 1:
            return docs[i - 2:i + 1], -1
        return docs[i - 1:i + 2], 1

    def lookup(self, doc_ids):
        import numpy as np
        ids = np.asarray(doc_ids, dtype=np.int64).reshape(-1)
        first = ((ids - 1) // self.per) * self.per + 1
        length = np.minimum(self.n, first + self.per - 1) - first + 1
        i = ids - first
        is_first = i == 0
        is_last = (~is_first) & (i == length - 1)
        alone = is_last & (i < 2)
        start = np.where(is_first, 0, np.where(is_last, np.maximum(i - 2, 0), i - 1))
        start = np.where(alone, i, start)
        count = np.where(is_first, np.minimum(length, 3), np.where(alone, 1, 3))
        cols = np.arange(3)[None, :]
        docs = np.where(cols < count[:, None], first[:, None] + start[:, None] + cols, -1)
        main = np.where(is_first, 0, np.where(is_last, -1, 1)).astype(np.int32)
        return docs, count.astype(np.int32), main


def synthetic_questions(a, rank):
    """NQ-shaped batch as CPU int64 tensors."""
    import numpy as np
    import torch
    rng = np.random.RandomState(1234 + rank)
    b = a.batch
    q_bert = torch.zeros(b, a.seq_ret, dtype=torch.int64)
    q_t5 = torch.zeros(b, 64, dtype=torch.int64)
    q_len = torch.zeros(b, dtype=torch.int64)
    dec = torch.zeros(b, a.dec, dtype=torch.int64)
    for i in range(b):
        n = int(rng.randint(8, 25))
        toks = torch.from_numpy(rng.randint(1000, 30000, size=n))
        q_bert[i, 0], q_bert[i, 1:1 + n], q_bert[i, 1 + n] = 101, toks, 102
        q_t5[i, :n], q_len[i] = toks, n
        m = int(rng.randint(2, 7))
        dec[i, 0], dec[i, 1:1 + m] = 30522, torch.from_numpy(rng.randint(1000, 30000, size=m))
    return dict(uid=-torch.arange(1, b + 1), q_bert=q_bert, q_types=torch.zeros_like(q_bert), q_t5=q_t5,
                q_len=q_len, dec=dec)


# ------------------------------------------------------------------------------------ CPU baseline
def cpu_mips_step_fn(a, nq):
    """step() runs one search of an nq-question batch over the bounded row sample with the
    FAISS-flat port (oracle/flat_ip.py) and returns wall seconds; scale = rows / sample rows."""
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
    queries = torch.randn(nq, a.dim, generator=g).half().float()

    def step():
        t0 = time.perf_counter()
        flat_ip_search(rows, None, queries, a.k)
        return time.perf_counter() - t0

    sample = "search: %d of %d rows (fp32 copies of fp16 values), %d questions, top-%d, time x%.3f" % (
        s_rows, a.rows, nq, a.k, a.rows / s_rows)
    return step, sample, cores, a.rows / s_rows


def cpu_reader_step_fn(a, passages=2):
    """step() runs the fp32 PyTorch restatement of the reader path (oracle/blocks.py) for ONE question
    and `passages` of its top-k passages: query tower, context tower, T5 encoder, FiD decoder + LM
    head; scale = batch * k / passages (the encoders dominate and are linear in passages)."""
    import torch
    from oracle import blocks as ob
    h, heads, ffn, layers, vocab = 768, 12, 3072, a.layers, 30720
    g = torch.Generator().manual_seed(7)

    def weights(decoder):
        w = {}

        def lin(name, o, i):
            w[name + ".weight"] = torch.randn(o, i, generator=g) * 0.02
            w[name + ".bias"] = torch.zeros(o)

        def ln(name):
            w[name + ".weight"], w[name + ".bias"] = torch.ones(h), torch.zeros(h)

        w["language_model.embedding.word_embeddings.weight"] = torch.randn(vocab, h, generator=g) * 0.02
        w["language_model.embedding.position_embeddings.weight"] = torch.randn(512, h, generator=g) * 0.02
        w["language_model.embedding.tokentype_embeddings.weight"] = torch.randn(2, h, generator=g) * 0.02
        for stack in (["encoder", "decoder"] if decoder else ["encoder"]):
            for i in range(layers):
                p = "language_model.%s.layers.%d" % (stack, i)
                ln(p + ".input_layernorm")
                lin(p + ".self_attention.query_key_value", 3 * h, h)
                lin(p + ".self_attention.dense", h, h)
                ln(p + ".post_attention_layernorm")
                if stack == "decoder":
                    lin(p + ".inter_attention.query", h, h)
                    lin(p + ".inter_attention.key_value", 2 * h, h)
                    lin(p + ".inter_attention.dense", h, h)
                    ln(p + ".post_inter_attention_layernorm")
                lin(p + ".mlp.dense_h_to_4h", ffn, h)
                lin(p + ".mlp.dense_4h_to_h", h, ffn)
            ln("language_model.%s.final_layernorm" % stack)
        w["lm_head.bias"] = torch.zeros(vocab)
        return w

    wb, wt = weights(False), weights(True)
    q = synthetic_questions(argparse.Namespace(batch=1, seq_ret=a.seq_ret, dec=a.dec), 0)
    ctx = torch.randint(1000, 30000, (passages, a.seq_ret), generator=g)
    ctx[:, 180:] = 0
    ext = torch.randint(1000, 30000, (passages, a.seq), generator=g)
    ext[:, 200:] = 0

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            ob.bert_pooled(q["q_bert"], q["q_types"], wb, heads, layers)
            ob.bert_pooled(ctx, torch.zeros_like(ctx), wb, heads, layers)
            enc = ob.t5_encode(ext, wt, heads, layers)
            ob.t5_decode(q["dec"], enc.reshape(1, passages * a.seq, h), ext.reshape(1, passages * a.seq), wt,
                         heads, layers)
        return time.perf_counter() - t0

    scale = a.batch * a.k / passages
    sample = "read: 1 question x %d of %d passages in fp32 on the host (oracle/blocks.py), time x%.1f" % (
        passages, a.k, scale)
    return step, sample, scale


def cpu_baseline(a, budget_s=25.0):
    """Reference CPU path on a bounded sample -> dict for the JSON line (rank 0, N=1 only)."""
    nq = a.nq if a.retrieve_only else a.batch * a.gpus
    mips_step, sample, cores, mips_scale = cpu_mips_step_fn(a, nq)
    mips_step()
    t_mips = min(mips_step() for _ in range(3)) * mips_scale
    if a.retrieve_only:
        return dict(value=nq / t_mips, unit="queries/s", cores=cores, kind="port", sample=sample + "; best of 3")
    read_step, read_sample, read_scale = cpu_reader_step_fn(a)
    read_step()
    t_read = read_step() * read_scale
    return dict(value=a.batch / (t_mips + t_read), unit="queries/s", cores=cores, kind="port",
                sample=sample + "; " + read_sample, search_s_per_step=t_mips, read_s_per_step=t_read)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nq = a.nq if a.retrieve_only else a.batch * a.gpus
    mips_step, sample, cores, mips_scale = cpu_mips_step_fn(a, nq)
    read = None if a.retrieve_only else cpu_reader_step_fn(a)
    for _ in range(max(1, min(a.warmup, 2))):
        mips_step()
        if read:
            read[0]()
    budget_s, times, wall0 = 150.0, [], time.perf_counter()
    for _ in range(max(1, a.steps)):             # bounded: stop early rather than run for hours
        t = mips_step() * mips_scale
        if read:
            t += read[0]() * read[2] * a.gpus        # the step reads batch * gpus questions (weak scaling)
        times.append(t)
        if time.perf_counter() - wall0 > budget_s:
            break
    sec = statistics.mean(times)
    queries_per_step = nq if a.retrieve_only else a.batch * a.gpus
    value = queries_per_step / sec
    full_sample = sample + ("" if not read else "; " + read[1] + (", x%d ranks' questions" % a.gpus if a.gpus > 1 else ""))
    line = {
        "impl": "reference", "metric": metric_name(a), "value": value, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": len(times), "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if a.retrieve_only else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a),
                   "path": "FaissMIPSIndex CPU (IndexFlatIP) restated by oracle/flat_ip.py + fp32 PyTorch "
                           "restatement of the reader (oracle/blocks.py); faiss is absent from the image and the "
                           "reference's own modules are CUDA-only"},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": full_sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def metric_name(a):
    if a.retrieve_only:
        return "queries/sec retrieve (brute-force MIPS top-%d) @%.0fM docs" % (a.k, a.rows / 1e6)
    if getattr(a, "train", False):
        return "queries/sec retrieve+read training step (fwd+bwd+optimizer) @%.0fM docs" % (a.rows / 1e6)
    return "queries/sec retrieve+read (forward) @%.0fM docs" % (a.rows / 1e6)


# ------------------------------------------------------------------------------------- B200 arm
def make_shard(torch, n_local, dim, dtype, seed, device):
    """Gaussian evidence rows generated on the device in 1M-row slices (never on the host)."""
    g = torch.Generator(device=device).manual_seed(seed)
    rows = torch.empty((n_local, dim), dtype=dtype, device=device)
    for r0 in range(0, n_local, 1 << 20):
        r1 = min(n_local, r0 + (1 << 20))
        rows[r0:r1] = (torch.randn(r1 - r0, dim, generator=g, device=device) / dim ** 0.5).to(dtype)
    return rows


class Dist(object):
    """Rank bookkeeping + the barrier / max-over-ranks timing helpers of the bench contract."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (a.gpus, self.world))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: emdr2_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner to stdout when NCCL_DEBUG is set; stdout carries the JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            # the collectives of this workload are latency-bound (51 KB all-gathers) or hidden under backward (gradient
            # buckets): a few CTAs are enough, and every SM NCCL holds is one the persistent GEMM grids must avoid
            os.environ.setdefault("NCCL_MAX_CTAS", str(NCCL_CTAS))
            dist.init_process_group("nccl", device_id=self.device)
        self.group = dist.group.WORLD if self.world > 1 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, step_fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        self.barrier()
        return self.max_over_ranks(ms)

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def retrieval_parity(d, index, rows, row_lo, queries, k, chunk=1 << 18):
    """Outside the timed region: the ids `index.search` returns for `queries` (the same collective call
    the step makes) against an INDEPENDENT exact search of the same resident rows — fp64 chunked
    torch.matmul over this rank's shard, exact local top-k, all-gather of the per-rank lists over NCCL
    and a (score desc, id asc) sort: the oracle's ranking definition (oracle/mips_oracle.c) evaluated
    with library calls only, none of this repo's kernels.  Tie-aware like
    tests/helpers.assert_ids_equal_outside_ties: ranks whose neighbouring exact scores are within
    2^-20 relative may swap (fp32 tensor-core accumulation order) and are compared as sets."""
    torch = d.torch
    got_s, got_i = index.search(queries, k)                   # [nq, k] on every rank
    q64 = queries.double()
    nq = q64.shape[0]
    n_local = rows.shape[0]
    kk = k + 1                                                # one extra rank: shows a tie at the cut
    dev = rows.device
    best_s = torch.zeros((nq, 0), dtype=torch.float64, device=dev)
    best_i = torch.zeros((nq, 0), dtype=torch.int64, device=dev)
    for r0 in range(0, n_local, chunk):
        r1 = min(n_local, r0 + chunk)
        s = q64 @ rows[r0:r1].double().T                      # exact products, fp64 accumulation
        ids = torch.arange(row_lo + r0 + 1, row_lo + r1 + 1, device=dev).expand(nq, -1)
        s, ids = torch.cat([best_s, s], 1), torch.cat([best_i, ids], 1)
        top = torch.topk(s, min(kk, s.shape[1]), dim=1)
        best_s, best_i = top.values, torch.gather(ids, 1, top.indices)
    if best_s.shape[1] < kk:                                  # short / empty shard: pad like the kernel
        pad = kk - best_s.shape[1]
        best_s = torch.cat([best_s, torch.full((nq, pad), float("-inf"), dtype=torch.float64, device=dev)], 1)
        best_i = torch.cat([best_i, torch.full((nq, pad), -1, dtype=torch.int64, device=dev)], 1)
    if d.world > 1:
        all_s = [torch.empty_like(best_s) for _ in range(d.world)]
        all_i = [torch.empty_like(best_i) for _ in range(d.world)]
        d.dist.all_gather(all_s, best_s.contiguous())
        d.dist.all_gather(all_i, best_i.contiguous())
        best_s, best_i = torch.cat(all_s, 1), torch.cat(all_i, 1)
    order = torch.argsort(best_i, dim=1, stable=True)         # id asc, then score desc (stable)
    best_s, best_i = torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
    order = torch.argsort(best_s, dim=1, descending=True, stable=True)[:, :kk]
    want_s, want_i = torch.gather(best_s, 1, order), torch.gather(best_i, 1, order)
    w32 = want_s.float()
    gap = (w32[:, :-1] - w32[:, 1:]).abs() <= w32[:, :-1].abs() * 2.0 ** -20      # [nq, k]
    tie = torch.zeros((nq, kk), dtype=torch.bool, device=dev)
    tie[:, :-1] |= gap
    tie[:, 1:] |= gap
    cut_tie = gap[:, -1]                                      # rank k ties with rank k+1
    tie, want_s, want_i = tie[:, :k], want_s[:, :k], want_i[:, :k]
    differ = (got_i != want_i) & ~tie
    sets_equal = (torch.sort(got_i, 1).values == torch.sort(want_i, 1).values).all(dim=1) | cut_tie
    scores_ok = bool(torch.allclose(got_s.double(), want_s, rtol=1e-5, atol=1e-6))
    n_differ = int(differ.sum().item())
    ok = n_differ == 0 and scores_ok and bool(sets_equal.all().item())
    return {"ids": "identical" if ok else "DIFFERENT", "checked_rows": int(d.sum_over_ranks(n_local)),
            "checked_queries": nq, "k": k, "ranks": d.world, "ids_differing_outside_ties": n_differ,
            "ranks_in_numerical_ties": int(tie.sum().item()), "scores_within_1e-5": scores_ok,
            "against": "independent fp64 chunked torch.matmul + exact top-k over the same resident rows, "
                       "(score desc, id asc), gathered over all ranks; tie-aware (2^-20 relative)"}


def mips_roofline(a, d, searcher, n_local, nq):
    scan_launches = searcher.stat("scan_launches")
    scan_ns = searcher.stat("scan_ns")
    algo_bytes = n_local * a.dim * 2 + nq * a.dim * 2 + nq * a.k * 12
    scan_ms = d.max_over_ranks(scan_ns / max(1, scan_launches) * 1e-6)
    peak, peak_src = measured_peak("hbm_gbs")
    achieved = algo_bytes / (scan_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": recorded_traffic(n_local, a.dim), "kernel": "emdr2::mips_scan_kernel",
            "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": scan_ms, "launches_timed": scan_launches,
            "peak_source": peak_src}


def run_retrieve_only(a):
    import torch
    from emdr2_b200.index import B200BruteForceIndex, chunk_range
    d = Dist(a)
    device, world, rank = d.device, d.world, d.rank
    tdtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype]
    lo, hi = chunk_range(a.rows, world, rank)
    rows = make_shard(torch, hi - lo, a.dim, tdtype, 1234 + rank, device)
    index = B200BruteForceIndex(a.dim, dtype=tdtype, device=device, group=d.group)
    index.add_local_shard(None, rows, num_rows=a.rows, row_lo=lo)       # ids = 1-based row numbers
    searcher = index._searcher
    gq = torch.Generator(device=device).manual_seed(99)                 # same queries on all ranks
    n_batches = 8
    q_dev = [torch.randn(a.nq, a.dim, generator=gq, device=device).to(tdtype) for _ in range(n_batches)]
    q_host = [q.cpu().pin_memory() for q in q_dev]
    out_d = torch.empty((a.nq, a.k), dtype=torch.float16).pin_memory()
    out_i = torch.empty((a.nq, a.k), dtype=torch.int32).pin_memory()

    def resident_step(i):
        return index.search(q_dev[i % n_batches], a.k)

    def e2e_step(i):
        q = q_host[i % n_batches].to(device, non_blocking=True)
        dd, ix = index.search_mips_index(q, a.k, reconstruct=False)
        out_d.copy_(dd, non_blocking=True)
        out_i.copy_(ix, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller consumes the ids on the host

    warm = max(3, a.warmup)
    for i in range(warm):
        resident_step(i)
        e2e_step(i)
    torch.cuda.synchronize()
    with ClockSampler(d.local_rank) as clocks:
        searcher.set_option("timing", 1)
        ms_total = d.timed(resident_step, a.steps)
        roof = mips_roofline(a, d, searcher, hi - lo, a.nq)
        searcher.set_option("timing", 0)
        ms_e2e = d.timed(e2e_step, a.steps)
    ms_step = ms_total / a.steps
    n_local = hi - lo
    line = {
        "metric": metric_name(a), "value": a.nq / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world,
        "steps": a.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": a.dtype + " inputs, fp32 accumulate (tcgen05 kind::f16)", "data": "synthetic",
        "config": {
            "workload": workload_name(a), "rows_per_gpu": n_local,
            "sharding": "torch.chunk row ranges, one rank per GPU",
            "exchange": "none" if world == 1 else "one all-gather of [nq,k] (fp32 score,int64 id) + k-way merge per rank",
            "l2": "inputs larger than L2 (%.2f GB evidence per GPU streamed per step vs 126 MB L2, evict-first)" % (n_local * a.dim * 2 / 1e9),
            "stage": "retrieve only (--retrieve-only)"},
        "e2e": {"value": a.nq / (ms_e2e / a.steps * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.dim * 2,
                "d2h_bytes_per_step": a.nq * a.k * (2 + 4),
                "api": "B200BruteForceIndex.search_mips_index (pinned host queries in, fp16 scores + int32 ids out)"},
        "gpu_launches": (2 if world == 1 else 3) * (-(-a.nq // 64)) * a.steps,
        "roofline": roof, "clocks": clocks.summary(),
    }
    line["parity"] = retrieval_parity(d, index, rows, lo, q_dev[0], a.k)
    if world == 1 and not a.no_gpu_reference:
        line["gpu_reference"] = gpu_reference_leg(torch, rows, q_dev, a)
    line["cpu_baseline"] = cpu_baseline(a) if (rank == 0 and world == 1 and not a.no_cpu_baseline) else None
    if rank == 0:
        emit(line)
    d.finish()
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
        return {"value": q_dev[0].shape[0] / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
                "what": "torch.matmul (cuBLAS) -> C[nq,N] in HBM -> torch.topk, same GPU, same shard"}
    except Exception as exc:        # e.g. out of memory for the C matrix on a small GPU
        return {"unavailable": str(exc).splitlines()[0][:200]}


def gpu_reference_read_leg(a, d, model, rows, all_q, dev, iters=3):
    """What the REFERENCE executes on a GPU for one step of this workload, on the same B200 and the same
    inputs: its retrieval (DistributedBruteForceIndex, megatron/data/emdr2_index.py:281-295: fp16 matmul into a
    C[nq, N] score matrix in HBM, then torch.topk) and its towers / reader as eager fp16 PyTorch — cuBLAS + ATen
    kernels, dense [b, s, s] masks, padded rectangles, unfused softmax — through oracle/blocks.py, the restatement
    of megatron/model/{transformer,language_model,t5_model,dualencoder_model}.py that tests/golden pins to the
    reference's own modules.  (The reference itself cannot travel to the GPU box, and its FusedLayerNorm / fused
    softmax need apex / THC.)  A reported baseline: nothing of this repo's kernels runs in it, and nothing of it
    runs in the product path.  The reference's Python retrieval tail (400 iterations with bisect + mmap reads per
    step, emdr2_model.py:457-468) and formatting loops are NOT included, which flatters the reference."""
    torch = d.torch
    try:
        from oracle import blocks as ob
        from emdr2_b200 import formatter
        st = model.settings
        h, heads, layers = a.dim, a.dim // 64, a.layers

        def w16(module):
            return {n: p.detach().to(torch.float16) for n, p in module.named_parameters()}

        wq, wc = w16(model.retriever_model.query_model), w16(model.retriever_model.context_model)
        wt = w16(model.language_model)
        rows16 = rows if rows.dtype == torch.float16 else rows.to(torch.float16)
        with torch.no_grad():
            topk_data, _ = model.evidence_retriever.get_topk(all_q[:a.batch].contiguous() if d.world == 1 else
                                                             all_q[d.rank * a.batch:(d.rank + 1) * a.batch].contiguous(),
                                                             as_packed=True)
            (ctx_ids, ctx_types, ext_ids, _one), _ = formatter.postprocess(
                dev["uid"], dev["q_t5"], dev["q_len"], topk_data, a.k, a.seq_ret, a.seq, st["cls_id"], st["sep_id"],
                st["pad_id"], device=dev["q_bert"].device, return_lengths=True)
        ctx_ids, ctx_types = ctx_ids.reshape(-1, a.seq_ret), ctx_types.reshape(-1, a.seq_ret)
        ext_ids = ext_ids.reshape(-1, a.seq)
        parts = {}

        def step():
            with torch.no_grad():
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
                q = ob.bert_pooled(dev["q_bert"], dev["q_types"], wq, heads, layers)
                ev[1].record()
                c_mat = torch.matmul(all_q.to(torch.float16), rows16.T)              # C[nq, N] fp16 in HBM (:281-292)
                torch.topk(c_mat, a.k, dim=1)                                         # (:295)
                ev[2].record()
                ctx = ob.bert_pooled(ctx_ids, ctx_types, wc, heads, layers).reshape(a.batch, a.k, h)
                sim = torch.bmm(q.unsqueeze(1).float(), ctx.float().transpose(1, 2)) / h ** 0.5
                torch.log_softmax(sim, dim=2)
                enc = ob.t5_encode(ext_ids, wt, heads, layers)
                logits = ob.t5_decode(dev["dec"], enc.reshape(a.batch, a.k * a.seq, h),
                                      ext_ids.reshape(a.batch, a.k * a.seq), wt, heads, layers)
                ev[3].record()
            return ev, logits

        step()
        torch.cuda.synchronize()
        tot = [0.0, 0.0, 0.0]
        for _ in range(iters):
            ev, _ = step()
            torch.cuda.synchronize()
            for j in range(3):
                tot[j] += ev[j].elapsed_time(ev[j + 1]) / iters
        ms = d.max_over_ranks(sum(tot))
        return {"value": a.batch * d.world / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
                "query_tower_ms": tot[0], "retrieve_ms": tot[1], "read_ms": tot[2], "steps": iters, "dtype": "fp16",
                "what": "eager fp16 PyTorch (cuBLAS + ATen) restatement of the reference's GPU step on the same B200, same "
                        "inputs, full shapes (oracle/blocks.py towers + reader, matmul -> C[nq,N] -> topk retrieval); the "
                        "reference's Python retrieval tail and formatting loops excluded"}
    except Exception as exc:
        return {"unavailable": (type(exc).__name__ + ": " + str(exc)).splitlines()[0][:300]}


def run_retrieve_read(a):
    import torch
    from emdr2_b200 import ops
    from emdr2_b200.index import chunk_range
    from emdr2_b200.model import EMDR2Model
    from emdr2_b200.retriever import B200EvidenceRetriever
    d = Dist(a)
    device, world, rank = d.device, d.world, d.rank
    mdtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.model_dtype]
    edtype = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype]

    # ---- evidence shard + retriever (ids = 1-based row numbers, the TSV numbering)
    lo, hi = chunk_range(a.rows, world, rank)
    rows = make_shard(torch, hi - lo, a.dim, edtype, 1234 + rank, device)
    retriever = B200EvidenceRetriever(a.k, a.dim, allow_trivial_doc=True, group=d.group, dtype=edtype,
                                      passages_map=synthetic_token_store(100, 180, 1), title_map=synthetic_token_store(2, 8, 2),
                                      wikititledocmap=SyntheticTitleMap(a.rows))
    retriever.mips_index.add_local_shard(None, rows, num_rows=a.rows, row_lo=lo)
    searcher = retriever.mips_index._searcher

    # ---- model: 2 x BERT-base towers + T5-base-shaped reader, random init N(0, 0.02)
    p_drop = 0.0 if a.no_dropout else 0.1        # the reference's --hidden-dropout / --attention-dropout defaults
    cfg = dict(hidden=a.dim, heads=a.dim // 64, layers=a.layers, ffn=4 * a.dim, vocab=30720, max_pos=512, dtype=mdtype,
               hidden_dropout=p_drop, attention_dropout=p_drop)
    settings = dict(topk_retrievals=a.k, seq_length=a.seq, seq_length_ret=a.seq_ret, retriever_score_scaling=True,
                    update_retriever=a.train, cls_id=101, sep_id=102, pad_id=0)
    torch.manual_seed(1234)
    model = EMDR2Model(cfg, retriever, settings, t5_vocab_size=30720, bert_vocab_size=30592).to(device)
    model.train(a.train)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "layernorm" in name:
                p.fill_(1.0 if name.endswith("weight") else 0.0)
            elif name.endswith("bias"):
                p.zero_()
            else:
                p.normal_(0.0, 0.02)

    host = synthetic_questions(a, rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.to(device) for k, v in host.items()}
    out_ids = torch.empty((a.batch, a.dec), dtype=torch.int64).pin_memory()
    out_lp = torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory()

    def forward(x):
        return model(x["uid"], x["q_bert"], x["q_types"], None, x["q_t5"], x["q_len"], x["dec"])

    sm_count = torch.cuda.get_device_properties(device).multi_processor_count
    labels_host = host["dec"].roll(-1, dims=1)
    labels_host[:, -1] = 0
    host["labels"] = labels_host
    pinned["labels"] = labels_host.pin_memory()
    dev["labels"] = labels_host.to(device)
    out_loss = torch.empty(2, dtype=torch.float32).pin_memory()
    trainer = {}

    def train_step(x):
        """forward (dropout on, incl. the no-grad one-context pass) -> both losses -> backward with the bucketed
        gradient all-reduce running underneath it -> AdamW on flat fp32 masters -> copy-back, one op per bucket."""
        from emdr2_b200 import losses
        if not trainer:
            from emdr2_b200.data_parallel import GradientBuckets, flatten_parameters
            trainer["buckets"] = gb = GradientBuckets(list(model.parameters()), group=d.group, main_grad=True)
            trainer["pflat"] = flatten_parameters(gb)
            trainer["masters"] = [f.float().requires_grad_(True) for f in trainer["pflat"]]
            for m, b in zip(trainer["masters"], gb.buckets):
                m.grad = b.grad              # the fp32 main-grad bucket IS the master gradient: nothing to cast or copy
            # torch.optim (library): the reference uses apex FusedAdam; the optimizer is a caller of the hot path
            trainer["opt"] = torch.optim.AdamW(trainer["masters"], lr=2e-5, weight_decay=0.01, fused=True)
        gb = trainer["buckets"]
        gb.start_step()
        lm_logits, topk_log_probs, one_ctx = forward(x)
        if world > 1 or os.environ.get("EMDR2_BENCH_FORCE_GEMM_CAP"):   # (the env switch: A/B of the cap alone at N = 1)
            # leave NCCL's CTAs their SMs while all-reduces overlap the backward GEMMs
            ops.set_option("gemm_max_ctas", sm_count - NCCL_CTAS)
        mask = (x["labels"] > 0).float()
        lm_loss = losses.reader_cross_entropy(lm_logits, x["labels"], mask)
        r_loss, _, _ = losses.get_loss_and_retriever_utility(one_ctx, topk_log_probs, x["labels"], mask, 30523)
        (lm_loss + r_loss).backward()
        gb.finish()
        ops.set_option("gemm_max_ctas", 0)
        if world > 1:
            for b in gb.buckets:
                b.grad.mul_(1.0 / world)
        trainer["opt"].step()
        with torch.no_grad():
            for f, m in zip(trainer["pflat"], trainer["masters"]):
                f.copy_(m)
        return lm_loss, r_loss

    def train_resident_step(i):
        return train_step(dev)

    def train_e2e_step(i):
        x = {k: v.to(device, non_blocking=True) for k, v in pinned.items()}
        lm_loss, r_loss = train_step(x)
        out_loss.copy_(torch.stack([lm_loss.detach(), r_loss.detach()]), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def fwd_resident_step(i):
        with torch.no_grad():
            return forward(dev)

    def fwd_e2e_step(i):
        x = {k: v.to(device, non_blocking=True) for k, v in pinned.items()}
        with torch.no_grad():
            lm_logits, topk_log_probs, _, _ = forward(x)
        out_ids.copy_(lm_logits.argmax(dim=-1), non_blocking=True)
        out_lp.copy_(topk_log_probs, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    nq_search = a.batch * world
    peak_t, peak_t_src = measured_peak("bf16_tflops_sustained")

    def measure(resident_step, e2e_step, steps, warm):
        """Timed regions of one mode: (1) `steps` resident steps, nothing else running -> value; (2) the same steps
        through pinned host buffers -> e2e; (3) the same steps again with every launch bracketed by CUDA events on
        its stream (emdr2_ops_timing) -> the per-kernel-kind split and the rooflines.  The split is taken in a pass
        of its own because two event records and a mutex per launch cost host time that a launch-bound step (the
        training step at N = 8) would pay in its headline number."""
        import gc
        for i in range(warm):
            resident_step(i)
        e2e_step(0)
        torch.cuda.synchronize()
        # Python's cyclic collector stays out of the timed regions (collected between them): a full collection walks
        # the millions of objects of the synthetic corpus / title maps and showed up as a sporadic +25 % on one region
        # of the launch-heavy training step (the same reason trainers collect manually between steps).
        gc.collect()
        gc.disable()
        try:
            with ClockSampler(d.local_rank, period=0.05) as clocks:
                ms_total = d.timed(resident_step, steps)
                gc.collect()
                ms_e2e = d.timed(e2e_step, steps)
            gc.collect()
            searcher.set_option("timing", 1)
            ops.timing(True)
            ms_inst = d.timed(resident_step, steps)
        finally:
            gc.enable()
        roof_mips = mips_roofline(a, d, searcher, hi - lo, nq_search)
        gemm_s, gemm_n, gemm_fl = ops.timing_read(ops.KIND_GEMM)
        attn_s, attn_n, attn_fl = ops.timing_read(ops.KIND_ATTENTION)
        row_s, row_n, _ = ops.timing_read(ops.KIND_ROWOP)
        ops.timing(False)
        searcher.set_option("timing", 0)
        gemm_tf = gemm_fl / d.max_over_ranks(gemm_s) / 1e12
        return dict(ms_step=ms_total / steps, ms_e2e_step=ms_e2e / steps, ms_instrumented_step=ms_inst / steps,
                    roof_mips=roof_mips, clocks=clocks.summary(),
                    gemm_ms=gemm_s / steps * 1e3, attn_ms=attn_s / steps * 1e3, row_ms=row_s / steps * 1e3,
                    gemm_tf=gemm_tf, attn_tf=attn_fl / max(attn_s, 1e-12) / 1e12, gemm_flops_step=gemm_fl / steps,
                    gemm_launches=gemm_n, launches_step=(gemm_n + attn_n + row_n) // steps + (2 if world == 1 else 3))

    warm = max(3, a.warmup)
    tokens = a.batch * (a.seq_ret + a.k * a.seq_ret + a.k * a.seq)
    fmt_h2d = a.batch * a.k * (a.seq_ret + 2 * a.seq) * 8      # ids of the three layouts (types are zeros made on the device)
    in_bytes = sum(v.numel() * 8 for v in host.values())

    def train_block(steps, warm_steps):
        """The full EMDR2 training step (BASELINE config 4: forward with dropout 0.1, backward, gradient all-reduce,
        optimizer) measured in this same run."""
        from emdr2_b200 import dropout
        dropout.manual_seed(1234 + rank)
        model.train(True)
        model.settings["update_retriever"] = True
        m = measure(train_resident_step, train_e2e_step, steps, warm_steps)
        gb = trainer["buckets"]
        # host time to ENQUEUE one step on an idle GPU (no synchronisation inside): how far the launching thread runs
        # ahead of the device — a step whose enqueue time approaches its device time is exposed to host noise
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        train_resident_step(0)
        host_enqueue_ms = (time.perf_counter() - t_host) * 1e3
        torch.cuda.synchronize()
        if os.environ.get("EMDR2_BENCH_PROFILE_HOST") and rank == 0:      # where the launching thread's time goes
            import cProfile
            import pstats
            prof = cProfile.Profile()
            prof.enable()
            for i in range(2):
                train_resident_step(i)
            prof.disable()
            torch.cuda.synchronize()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(45)
        ar_ms = None
        if world > 1:          # the exchange alone: all buckets back to back, nothing to hide behind
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                for b in gb.buckets:
                    d.dist.all_reduce(b.grad)
            e1.record()
            torch.cuda.synchronize()
            ar_ms = d.max_over_ranks(e0.elapsed_time(e1) / 3)
        return {"value": a.batch * world / (m["ms_step"] * 1e-3), "unit": "queries/s", "ms_per_step": m["ms_step"],
                "steps": steps, "warmup": warm_steps, "host_enqueue_ms_per_step": host_enqueue_ms,
                "e2e": {"value": a.batch * world / (m["ms_e2e_step"] * 1e-3), "unit": "queries/s",
                        "h2d_bytes_per_step": in_bytes + fmt_h2d, "d2h_bytes_per_step": 8 + a.batch * a.k * 4},
                "kernel_time_ms_per_step": {"gemm": m["gemm_ms"], "attention": m["attn_ms"], "rowops": m["row_ms"],
                                            "mips_scan": m["roof_mips"]["kernel_ms"], "gemm_tflops": m["gemm_tf"],
                                            "attention_tflops": m["attn_tf"],
                                            "measured_in": "a separate instrumented pass of the same steps",
                                            "instrumented_ms_per_step": m["ms_instrumented_step"]},
                "gemm_frac_of_sustained_peak": m["gemm_tf"] / peak_t,
                "gradient_allreduce": {"buckets": len(gb.buckets), "bytes": sum(b.grad.numel() * b.grad.element_size() for b in gb.buckets),
                                       "launched_from_backward_hooks": gb.launched,
                                       "exposed_alone_ms": ar_ms,
                                       "how": "one async NCCL all-reduce per 64 MB fp32 bucket, launched when the bucket's "
                                              "last gradient contribution is enqueued (overlaps the rest of backward); the "
                                              "backward kernels accumulate straight into the flat fp32 buckets, which are also "
                                              "the optimizer's master gradients (no zero fills, casts, flatten/unflatten)"},
                "dropout": {"hidden": cfg["hidden_dropout"], "attention": cfg["attention_dropout"],
                            "how": "counter-based masks regenerated in the backward kernels (csrc/dropout.cuh)"},
                "gpu_launches": int(m["launches_step"] * steps), "clocks": m["clocks"],
                "stage": "full training step: forward incl. the no-grad one-context pass (dropout on), reader + retriever "
                         "losses, backward, bucketed gradient all-reduce overlapped with backward, fused AdamW on flat fp32 "
                         "masters (torch.optim, library — the reference uses apex FusedAdam)"}

    def refresh_block(train_ms_alone):
        """BASELINE config 5 on the SAME GPUs: every rank re-encodes rows of its own shard with a frozen copy of the
        context tower on a side stream (async_indexer.ConcurrentShardRefresher) while the training steps keep
        running; then all ranks swap the standby shard in between two steps.  Bounded sample: --refresh-rows
        passages per GPU (S = 256, indexer batch 128, arguments.py:589), the rest of the standby copied."""
        import numpy as np
        from emdr2_b200.async_indexer import ConcurrentShardRefresher
        index = retriever.mips_index
        n_rows = min(a.refresh_rows, hi - lo) // 128 * 128
        if n_rows <= 0:
            return {"unavailable": "shard smaller than one indexer batch"}
        rng = np.random.RandomState(77 + rank)
        host_batches = []
        for b0 in range(0, n_rows, 128):
            ids = np.zeros((128, a.seq_ret), dtype=np.int64)
            for r, ln in enumerate(rng.randint(105, 192, size=128)):
                ids[r, 0] = 101
                ids[r, 1:ln] = rng.randint(1000, 30000, size=ln - 1)
            host_batches.append((torch.arange(lo + b0 + 1, lo + b0 + 129), torch.from_numpy(ids).pin_memory(),
                                 torch.zeros((128, a.seq_ret), dtype=torch.int64).pin_memory()))
        refresher = ConcurrentShardRefresher(index, model.retriever_model.context_model, lambda: iter(host_batches),
                                             group=d.group, partial=True)
        # (1) the refresh alone (no training running): the rate the GPU gives the indexer when it has it all
        d.barrier()
        refresher.start()
        while not refresher.maybe_swap():
            time.sleep(0.002)
        alone_s = d.max_over_ranks(refresher.seconds)
        # (2) the refresh while training steps run
        d.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        refresher.start()
        e0.record()
        steps_during, swapped = 0, False
        while not swapped and steps_during < 400:
            train_resident_step(steps_during)
            steps_during += 1
            swapped = refresher.maybe_swap()
        e1.record()
        torch.cuda.synchronize()
        busy_s = d.max_over_ranks(refresher.seconds)
        ms_during = d.max_over_ranks(e0.elapsed_time(e1)) / max(1, steps_during)
        # (3) after the swap: the step's retrieval against an independent search of the NEW resident rows, and the
        # refreshed rows themselves against the frozen tower (first indexer batch)
        model.train(False)
        with torch.no_grad():
            q_emb = model.retriever_embedder(dev["q_bert"], None, dev["q_types"], "query").to(edtype).contiguous()
            if world > 1:
                all_q2 = torch.empty((world * q_emb.shape[0], q_emb.shape[1]), dtype=q_emb.dtype, device=device)
                d.dist.all_gather_into_tensor(all_q2, q_emb)
            else:
                all_q2 = q_emb
            par = retrieval_parity(d, index, index.evidence_embeds, lo, all_q2, retriever.topk)
            rid, tok, typ = host_batches[0]
            t = tok.numpy()
            lens = ((t != 0) * np.arange(1, t.shape[1] + 1)).max(axis=1)
            again = refresher.tower(tok.to(device), None, typ.to(device), max_len=int(lens.max()), row_lengths=lens)
            rows_match = bool(torch.equal(again.to(edtype), index.evidence_embeds[:128]))
        model.train(True)
        return {"passages_per_s_while_training": n_rows * world / busy_s, "passages_per_s_alone": n_rows * world / alone_s,
                "unit": "passages/s (all %d GPUs)" % world, "rows_reencoded_per_gpu": n_rows,
                "train_ms_per_step_alone": train_ms_alone, "train_ms_per_step_during_refresh": ms_during,
                "step_time_inflation": ms_during / train_ms_alone, "training_steps_during_refresh": steps_during,
                "full_refresh_21M_s_at_this_rate": a.rows / (n_rows * world / busy_s),
                "swap": {"collective": "all ranks agree (all-reduce of ready flags) and swap between the same two steps",
                         "swapped": bool(swapped), "rounds": refresher.rounds, "parity_after_swap": par["ids"],
                         "refreshed_rows_equal_frozen_tower_output": rows_match},
                "how": "each rank re-encodes its OWN row range (frozen context-tower copy, side CUDA stream, worker "
                       "thread) straight into a standby shard in HBM while training steps run on the main stream; "
                       "bounded sample, remaining standby rows copied from the live shard"}

    if a.train:
        steps = a.steps
        tb = train_block(steps, warm)
        head = {"value": tb["value"], "ms_step": tb["ms_per_step"], "e2e": tb["e2e"], "clocks": tb["clocks"],
                "launches": tb["gpu_launches"]}
        m = None
    else:
        model.train(False)
        m = measure(fwd_resident_step, fwd_e2e_step, a.steps, warm)
        head = {"value": a.batch * world / (m["ms_step"] * 1e-3), "ms_step": m["ms_step"], "clocks": m["clocks"],
                "launches": int(m["launches_step"] * a.steps),
                "e2e": {"value": a.batch * world / (m["ms_e2e_step"] * 1e-3), "unit": "queries/s",
                        "h2d_bytes_per_step": in_bytes + fmt_h2d,
                        "d2h_bytes_per_step": out_ids.numel() * 8 + out_lp.numel() * 4 + a.batch * a.k * 4}}
    head["e2e"]["api"] = ("EMDR2Model.forward + losses + backward + optimizer (pinned host batch in; the two losses out)"
                          if a.train else
                          "EMDR2Model.forward (pinned host question tensors in; greedy token ids + passage log-probs out)") + \
        "; includes the host-side passage lookup/formatting"
    line = {
        "metric": metric_name(a), "value": head["value"], "unit": "queries/s", "n_gpus": world,
        "steps": a.steps, "warmup": warm, "ms_per_step": head["ms_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.model_dtype + " reader/retriever, " + a.dtype + " evidence, fp32 accumulate",
        "data": "synthetic",
        "config": {
            "workload": workload_name(a),
            "rows_per_gpu": hi - lo, "global_batch": a.batch * world, "encoder_tokens_per_gpu_per_step": tokens,
            "sharding": "evidence rows by torch.chunk over ranks; questions data-parallel",
            "exchange": "none" if world == 1 else "all-gather of queries [B,768] + all-gather of [nq,k] (score,id) pairs + merge",
            "l2": "inputs larger than L2 (%.2f GB evidence + %.1f GB of activations streamed per step vs 126 MB L2)" % (
                (hi - lo) * a.dim * 2 / 1e9, tokens * a.dim * 2 * 40 / 1e9),
            "stage": tb["stage"] if a.train else
            "retrieve + read FORWARD (EMDR2Model.forward eval path); no backward/optimizer in the timed region "
            "(the full training step measured in the same run is under train_step)"},
        "e2e": head["e2e"], "gpu_launches": head["launches"], "clocks": head["clocks"],
    }
    if m is not None:
        line["roofline"] = {
            "bound": "tensor", "achieved": m["gemm_tf"], "peak": peak_t, "unit": "TFLOP/s", "frac": m["gemm_tf"] / peak_t,
            "traffic": recorded_gemm_traffic(),
            "traffic_note": "mean DRAM bytes/launch of the layer's four projections at 102400 tokens (ncu --set full, "
                            "profiles/roofline_traffic.json); the bound is the tensor pipe, not DRAM",
            "kernel": "emdr2::gemm_kernel (+ emdr2::gemm_pair_kernel for residual epilogues over >=100k rows)",
            "algorithmic_flops_per_step": m["gemm_flops_step"], "kernel_ms_per_step": m["gemm_ms"],
            "launches_timed": m["gemm_launches"], "peak_source": peak_t_src}
        line["roofline_mips"] = m["roof_mips"]
        line["kernel_time_ms_per_step"] = {"gemm": m["gemm_ms"], "attention": m["attn_ms"], "rowops": m["row_ms"],
                                           "mips_scan": m["roof_mips"]["kernel_ms"], "attention_tflops": m["attn_tf"],
                                           "measured_in": "a separate pass of the same steps with CUDA events around every "
                                                          "launch (emdr2_ops_timing), right after the timed region",
                                           "instrumented_ms_per_step": m["ms_instrumented_step"]}
    else:
        line["roofline"] = {"bound": "tensor", "achieved": tb["kernel_time_ms_per_step"]["gemm_tflops"], "peak": peak_t,
                            "unit": "TFLOP/s", "frac": tb["gemm_frac_of_sustained_peak"], "traffic": recorded_gemm_traffic(),
                            "kernel": "emdr2::gemm_kernel (forward, dX and split-K dW products)", "peak_source": peak_t_src}
        line["train_step"] = tb
    # parity of THIS step's retrieval (outside the timed region): the question embeddings the query tower
    # produces for the batch, gathered over the ranks exactly as get_topk does, searched by the index
    model.train(False)
    with torch.no_grad():
        q_emb = model.retriever_embedder(dev["q_bert"], None, dev["q_types"], "query").to(edtype).contiguous()
        if world > 1:
            all_q = torch.empty((world * q_emb.shape[0], q_emb.shape[1]), dtype=q_emb.dtype, device=device)
            d.dist.all_gather_into_tensor(all_q, q_emb)
        else:
            all_q = q_emb
        line["parity"] = retrieval_parity(d, retriever.mips_index, rows, lo, all_q, retriever.topk)
    if not a.train and not a.no_gpu_reference:
        line["gpu_reference"] = gpu_reference_read_leg(a, d, model, rows, all_q, dev)
    if not a.train and not a.no_train_step:
        line["train_step"] = train_block(a.train_steps, 12)   # the first seconds of training on a fresh box run ~10 % slow
    if "train_step" in line and a.refresh_rows > 0:
        try:
            line["index_refresh"] = refresh_block(line["train_step"]["ms_per_step"])
        except Exception as exc:
            line["index_refresh"] = {"unavailable": (type(exc).__name__ + ": " + str(exc)).splitlines()[0][:300]}
    line["cpu_baseline"] = cpu_baseline(a) if (rank == 0 and world == 1 and not a.no_cpu_baseline) else None
    if rank == 0:
        emit(line)
    d.finish()
    return 0


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: keep a private handle on the real stdout and point
    fd 1 at stderr for everything else (NCCL's version banner, library chatter, stray prints)."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_retrieve_only(a) if a.retrieve_only else run_retrieve_read(a)


if __name__ == "__main__":
    sys.exit(main())
