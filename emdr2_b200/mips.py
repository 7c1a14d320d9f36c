"""Torch-tensor front end of the C-ABI MIPS search (one evidence shard on one GPU).

PyTorch is plumbing here: it owns device memory and streams; every FLOP runs in
libemdr2_b200.so (csrc/mips_scan.cu, csrc/mips_merge.cu).
"""
import ctypes

import torch

from . import _lib

_DTYPES = {torch.float16: _lib.EMDR2_DTYPE_FP16, torch.bfloat16: _lib.EMDR2_DTYPE_BF16}


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ShardSearcher(object):
    """Fused GEMM+top-k search over one resident evidence shard ``rows`` [n, d] (fp16/bf16, CUDA).

    ``ids`` is an int64 CUDA tensor [n] of doc ids or None (id = id_base + row).
    Ties rank (score desc, row asc); keep rows in ascending id order for (score desc, id asc).
    """

    def __init__(self, d, dtype, device):
        if dtype not in _DTYPES:
            raise TypeError("evidence dtype must be torch.float16 or torch.bfloat16, got %r" % (dtype,))
        self.lib = _lib.load()
        self.d = int(d)
        self.dtype = dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("emdr2_b200 has no CPU path; device must be a CUDA device")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        handle = ctypes.c_void_p()
        _lib.check(self.lib.emdr2_mips_create(self.d, _DTYPES[dtype], index, ctypes.byref(handle)),
                   "emdr2_mips_create")
        self._h = handle
        self._rows = None
        self._ids = None
        self.n = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.emdr2_mips_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._rows = None
        self._ids = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_shard(self, rows, ids=None, id_base=0):
        if rows.dim() != 2 or rows.shape[1] != self.d:
            raise ValueError("rows must be [n, %d], got %s" % (self.d, tuple(rows.shape)))
        if rows.dtype != self.dtype or rows.device != self.device or not rows.is_contiguous():
            raise ValueError("rows must be a contiguous %s tensor on %s" % (self.dtype, self.device))
        if ids is not None:
            if ids.dtype != torch.int64 or ids.device != self.device or ids.numel() != rows.shape[0] \
                    or not ids.is_contiguous():
                raise ValueError("ids must be a contiguous int64 tensor [n] on %s" % (self.device,))
        n = rows.shape[0]
        _lib.check(self.lib.emdr2_mips_set_shard(
            self._h, ctypes.c_void_p(rows.data_ptr() if n else 0),
            ctypes.c_void_p(ids.data_ptr()) if ids is not None and n else None, n, int(id_base)),
            "emdr2_mips_set_shard")
        self._rows, self._ids, self.n = rows, ids, n   # keep the borrowed buffers alive

    def search(self, queries, k):
        """queries [nq, d] on the shard's device -> (scores fp32 [nq, k], ids int64 [nq, k])."""
        if queries.dim() != 2 or queries.shape[1] != self.d:
            raise ValueError("queries must be [nq, %d], got %s" % (self.d, tuple(queries.shape)))
        if queries.device != self.device:
            raise ValueError("queries must live on %s" % (self.device,))
        q = queries.to(self.dtype).contiguous()
        nq = q.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        _lib.check(self.lib.emdr2_mips_search(
            self._h, ctypes.c_void_p(q.data_ptr() if nq else 0), nq, int(k),
            ctypes.c_void_p(scores.data_ptr() if nq else 0), ctypes.c_void_p(ids.data_ptr() if nq else 0),
            _stream_ptr(self.device)), "emdr2_mips_search")
        return scores, ids

    def search_host(self, queries_cpu, k):
        """Host-buffer round trip (FAISS-style): CPU tensor in, CPU (scores, ids) out; synchronous."""
        q = queries_cpu.to(self.dtype).contiguous()
        if q.device.type != "cpu":
            raise ValueError("search_host takes a CPU tensor")
        nq = q.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32)
        ids = torch.empty((nq, k), dtype=torch.int64)
        _lib.check(self.lib.emdr2_mips_search_host(
            self._h, ctypes.c_void_p(q.data_ptr() if nq else 0), nq, int(k),
            ctypes.c_void_p(scores.data_ptr() if nq else 0), ctypes.c_void_p(ids.data_ptr() if nq else 0),
            _stream_ptr(self.device)), "emdr2_mips_search_host")
        return scores, ids

    def set_option(self, name, value):
        _lib.check(self.lib.emdr2_mips_set_option(self._h, name.encode(), int(value)),
                   "emdr2_mips_set_option")

    def stat(self, name):
        out = ctypes.c_int64()
        _lib.check(self.lib.emdr2_mips_get_stat(self._h, name.encode(), ctypes.byref(out)),
                   "emdr2_mips_get_stat")
        return out.value


def merge_topk(scores, ids):
    """Merge [parts, nq, k] per-shard lists (CUDA fp32 / int64) into [nq, k], (score desc, id asc)."""
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores/ids must both be [parts, nq, k]")
    if scores.dtype != torch.float32 or ids.dtype != torch.int64 or not scores.is_cuda:
        raise ValueError("merge_topk takes CUDA fp32 scores and int64 ids")
    scores, ids = scores.contiguous(), ids.contiguous()
    parts, nq, k = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    lib = _lib.load()
    with torch.cuda.device(scores.device):
        _lib.check(lib.emdr2_mips_merge(
            ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ids.data_ptr()), parts, nq, k,
            ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
            _stream_ptr(scores.device)), "emdr2_mips_merge")
    return out_s, out_i
