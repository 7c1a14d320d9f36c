"""Generate tests/golden/mips_ref_*.npz by running the REFERENCE's own DistributedBruteForceIndex.

Runs only in the build container (needs /root/reference); the outputs are committed so the GPU box
never needs the reference.  The reference class (megatron/data/emdr2_index.py:200-305) hard-codes
'cuda:i' device strings, so it is executed here on CPU tensors with those strings patched out:

  * torch.cuda.device_count()        -> NGPU (the number of row chunks, :204,252)
  * Tensor.to('cuda:i') / .cuda()    -> identity
  * torch.zeros(..., device="cuda")  -> CPU

Everything else — dict -> np.array -> torch.chunk split, fp16 matmul, fp16 score matrix C,
torch.topk, the id_map loop — is the reference's code, unmodified.  Import shims are the four of
SURVEY.md §8c (torch._six, apex, amp_C, np.float).

Cases (seed 1234 = the reference's default --seed, arguments.py:275):
  c1_exact   BASELINE configs[0]: 1000 x 128, 32 queries, top-5, integer-valued rows in [-3, 3] so
             every score is an integer < 2048 — exactly representable in fp16, no rounding anywhere
             (ties are plentiful, which is the point: the tie groups are pinned too).
  c1_gauss   same shape, Gaussian rows/queries rounded to fp16: scores pass through fp16 (:284).
  shard3     777 x 64, 16 queries, top-7, 3 row chunks (last one short) and shuffled doc ids.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    six = types.ModuleType("torch._six")
    six.inf = float("inf")
    sys.modules["torch._six"] = six
    apex = types.ModuleType("apex")
    apex_opt = types.ModuleType("apex.optimizers")
    apex_opt.FusedAdam = object
    apex_mta = types.ModuleType("apex.multi_tensor_apply")
    apex_mta.multi_tensor_applier = None
    apex.optimizers, apex.multi_tensor_apply = apex_opt, apex_mta
    sys.modules.update({"apex": apex, "apex.optimizers": apex_opt,
                        "apex.multi_tensor_apply": apex_mta, "amp_C": types.ModuleType("amp_C")})
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)


class _Store(object):
    """Minimal stand-in for OpenRetreivalDataStore's data members (emdr2_index.py:16-43)."""

    def __init__(self, ids, rows):
        self.embed_data = {int(i): np.float16(r) for i, r in zip(ids, rows)}
        self.embedding_path = "unused.pkl"

    def clear(self):
        self.embed_data = dict()


def run_reference(ids, rows_f16, queries_f16, k, ngpu):
    from megatron.data import emdr2_index as ref
    orig_to, orig_zeros, orig_cuda = torch.Tensor.to, torch.zeros, torch.Tensor.cuda
    orig_count = torch.cuda.device_count

    def to_patch(self, *a, **kw):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        if isinstance(kw.get("device"), str) and kw["device"].startswith("cuda"):
            kw.pop("device")
        return orig_to(self, *a, **kw) if (a or kw) else self

    def zeros_patch(*a, **kw):
        if kw.get("device") == "cuda":
            kw.pop("device")
        return orig_zeros(*a, **kw)

    torch.Tensor.to, torch.zeros = to_patch, zeros_patch
    torch.Tensor.cuda = lambda self, *a, **kw: self
    torch.cuda.device_count = lambda: ngpu
    try:
        index = ref.DistributedBruteForceIndex(embed_size=rows_f16.shape[1],
                                               embed_data=_Store(ids, rows_f16))
        dist, idx = index.search_mips_index(torch.from_numpy(queries_f16), k, reconstruct=False)
    finally:
        torch.Tensor.to, torch.zeros, torch.Tensor.cuda = orig_to, orig_zeros, orig_cuda
        torch.cuda.device_count = orig_count
    assert dist.dtype == torch.float16 and idx.dtype == torch.int32
    return dist.numpy(), idx.numpy()


def make_case(name, n, d, nq, k, ngpu, kind, shuffle_ids):
    rng = np.random.RandomState(1234)
    if kind == "int":
        rows = rng.randint(-3, 4, size=(n, d)).astype(np.float16)
        queries = rng.randint(-3, 4, size=(nq, d)).astype(np.float16)
    else:
        rows = (rng.randn(n, d) / np.sqrt(d)).astype(np.float16)
        queries = rng.randn(nq, d).astype(np.float16)
    ids = np.arange(1, n + 1, dtype=np.int64)
    if shuffle_ids:
        ids = rng.permutation(ids)
    dist, idx = run_reference(ids, rows, queries, k, ngpu)
    out = os.path.join(HERE, "mips_ref_%s.npz" % name)
    np.savez_compressed(out, ids=ids, rows=rows, queries=queries, k=np.int64(k),
                        ngpu=np.int64(ngpu), ref_distances=dist, ref_indices=idx)
    print("%s: n=%d d=%d nq=%d k=%d ngpu=%d -> %s (%d bytes)" % (
        name, n, d, nq, k, ngpu, out, os.path.getsize(out)))


if __name__ == "__main__":
    install_shims()
    torch.manual_seed(1234)
    make_case("c1_exact", 1000, 128, 32, 5, 1, "int", False)
    make_case("c1_gauss", 1000, 128, 32, 5, 1, "gauss", False)
    make_case("shard3", 777, 64, 16, 7, 3, "gauss", True)
