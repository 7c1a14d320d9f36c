"""CPU restatement of the reference's MIPS search (numpy + the C library built from mips_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/mips_oracle.c for the header that states what this follows
(reference megatron/data/emdr2_index.py:164-197 and :241-305) and how it is pinned (PARITY UNPINNED
by the reference's own tests; pinned against tests/golden/mips_ref_*.npz produced by running the
reference's DistributedBruteForceIndex in the build container).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_mips.so")
_LIB = None


def build(force=False):
    """gcc the C restatement (Makefile next to this file)."""
    src = os.path.join(_HERE, "mips_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle_mips.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.oracle_mips_topk.restype = ctypes.c_int
        lib.oracle_mips_topk.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_mips_merge.restype = ctypes.c_int
        lib.oracle_mips_merge.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
        _LIB = lib
    return _LIB


def _as_u16(x):
    """Accept numpy float16, or raw uint16 bit patterns (bf16 has no numpy dtype)."""
    x = np.ascontiguousarray(x)
    if x.dtype == np.float16:
        return x.view(np.uint16), 0
    if x.dtype == np.uint16:
        return x, 1
    raise TypeError("oracle takes float16 arrays or uint16 views of bfloat16, got %s" % x.dtype)


def bf16_bits_to_f32(bits):
    return (np.ascontiguousarray(bits).astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x):
    """Round-to-nearest-even fp32 -> bf16 bit patterns (matches torch .to(torch.bfloat16))."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + 0x7FFF + ((u >> 16) & 1)
    return (rounded >> 16).astype(np.uint16)


def mips_topk(evidence, queries, k, ids=None, id_base=0, round_fp16=False, want_ties=False,
              threads=None):
    """C oracle: exact (double-accumulated, fp32-rounded) scores; ranking (score desc, id asc).

    evidence [n, d], queries [nq, d]: float16, or uint16 bit patterns of bfloat16.
    Returns (scores fp32 [nq, k], ids int64 [nq, k]) (+ tie_mask uint8 [nq, k] if want_ties).
    Mirrors DistributedBruteForceIndex.search_mips_index (emdr2_index.py:268-305); with
    round_fp16=True the scores pass through fp16 like the reference's C matrix (:284).
    """
    e, dt_e = _as_u16(evidence)
    q, dt_q = _as_u16(queries)
    if dt_e != dt_q:
        raise TypeError("evidence and queries must share a dtype")
    n, d = e.shape
    nq = q.shape[0]
    assert q.shape[1] == d
    out_s = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    ties = np.zeros((nq, k), dtype=np.uint8)
    ids_arr = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    lib = _lib()

    def run(q0, q1):  # ctypes releases the GIL: query ranges run on separate host threads
        return lib.oracle_mips_topk(
            e.ctypes.data, None if ids_arr is None else ids_arr.ctypes.data, int(id_base), n, d,
            dt_e, q[q0:q1].ctypes.data, q1 - q0, k, 1 if round_fp16 else 0,
            out_s[q0:q1].ctypes.data, out_i[q0:q1].ctypes.data, ties[q0:q1].ctypes.data)

    nthreads = max(1, min(threads or (os.cpu_count() or 1), nq))
    per = -(-nq // nthreads) if nq else 1
    ranges = [(a, min(nq, a + per)) for a in range(0, nq, per)]
    if len(ranges) <= 1:
        rcs = [run(a, b) for a, b in ranges]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(ranges)) as ex:
            rcs = list(ex.map(lambda ab: run(*ab), ranges))
    if any(rcs):
        raise RuntimeError("oracle_mips_topk failed with %s" % rcs)
    return (out_s, out_i, ties) if want_ties else (out_s, out_i)


def mips_topk_numpy(evidence_f, queries_f, k, ids=None, id_base=0, block=65536):
    """Pure-numpy second opinion (float64 GEMM in row blocks + lexsort); inputs already float."""
    e = np.asarray(evidence_f, dtype=np.float64)
    q = np.asarray(queries_f, dtype=np.float64)
    n = e.shape[0]
    all_ids = (np.arange(n, dtype=np.int64) + id_base) if ids is None else np.asarray(ids, np.int64)
    best_s = np.empty((q.shape[0], 0), dtype=np.float32)
    best_i = np.empty((q.shape[0], 0), dtype=np.int64)
    for r0 in range(0, n, block):
        s = (q @ e[r0:r0 + block].T).astype(np.float32)
        i = np.broadcast_to(all_ids[r0:r0 + block], s.shape)
        cs = np.concatenate([best_s, s], axis=1)
        ci = np.concatenate([best_i, i], axis=1)
        keep_s = np.empty((q.shape[0], min(k, cs.shape[1])), np.float32)
        keep_i = np.empty(keep_s.shape, np.int64)
        for r in range(q.shape[0]):
            order = np.lexsort((ci[r], -cs[r].astype(np.float64)))[:k]
            keep_s[r], keep_i[r] = cs[r][order], ci[r][order]
        best_s, best_i = keep_s, keep_i
    if best_s.shape[1] < k:
        pad = k - best_s.shape[1]
        best_s = np.concatenate([best_s, np.full((q.shape[0], pad), -np.inf, np.float32)], 1)
        best_i = np.concatenate([best_i, np.full((q.shape[0], pad), -1, np.int64)], 1)
    return best_s, best_i


def merge_topk(scores, ids):
    """[parts, nq, k] -> [nq, k] under (score desc, id asc); id < 0 is padding
    (restates the gather + global topk of emdr2_index.py:284-295 on per-shard lists)."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    parts, nq, k = scores.shape
    out_s = np.empty((nq, k), np.float32)
    out_i = np.empty((nq, k), np.int64)
    _lib().oracle_mips_merge(scores.ctypes.data, ids.ctypes.data, parts, nq, k,
                             out_s.ctypes.data, out_i.ctypes.data)
    return out_s, out_i


def chunk_rows(n, world):
    """Row ranges of torch.chunk(block_embeds, world, dim=0) (emdr2_index.py:252-256):
    chunk size ceil(n/world); trailing ranks may get a short or EMPTY range."""
    size = -(-n // world) if n else 0
    out = []
    for r in range(world):
        lo = min(n, r * size)
        hi = min(n, lo + size)
        out.append((lo, hi))
    return out


def brute_force_index_search(embed_items, query_f16, top_k, world=1):
    """Restatement of DistributedBruteForceIndex end to end (emdr2_index.py:241-305):
    dict {doc_id: fp16 row} in insertion order -> chunked shards -> scores -> global top-k ->
    id_map.  Returns what the reference returns: (distances float16 [nq,k], indices int32 [nq,k]),
    with ties ranked (score desc, id asc) — ids must therefore be inserted in ascending order for
    the per-shard / merged rankings to coincide (asserted)."""
    ids = np.array([i for i, _ in embed_items], dtype=np.int64)
    rows = np.array([np.float16(v) for _, v in embed_items], dtype=np.float16)
    parts_s, parts_i = [], []
    for lo, hi in chunk_rows(len(ids), world):
        assert np.all(np.diff(ids[lo:hi]) > 0), "rows must be stored in ascending id order per shard"
        s, i = mips_topk(rows[lo:hi], np.asarray(query_f16, np.float16), top_k, ids=ids[lo:hi],
                         round_fp16=True)
        parts_s.append(s)
        parts_i.append(i)
    s, i = merge_topk(np.stack(parts_s), np.stack(parts_i))
    return s.astype(np.float16), i.astype(np.int32)
