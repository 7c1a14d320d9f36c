"""CPU port of the reference's CPU search path: FaissMIPSIndex -> faiss.IndexIDMap(IndexFlatIP).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/mips_oracle.c): used by tests/ and by bench.py's
``cpu_baseline`` and ``--impl reference`` legs; never by emdr2_b200/.

The arithmetic lives in a third-party dependency that is NOT under /root/reference:
``facebookresearch/faiss``, cloned at an UNPINNED git HEAD by docker/Dockerfile:26-28 and absent
from this image (no wheel, no network).  Call sites in the reference: megatron/data/emdr2_index.py
:123 (IndexFlatIP), :135 (IndexIDMap), :171-177 (add_with_ids of fp32 rows), :188-196 (search of
fp32 queries -> (distances fp32 [nq,k], ids int64 [nq,k])).

Published algorithm restated here (faiss/utils/distances.cpp, ``exhaustive_inner_product_blas``,
used by IndexFlat::search for nq >= 20): for each block of database rows (1024 rows, faiss's
``distance_compute_blas_database_bs``) compute the [nq, block] inner-product panel with one SGEMM
and push it into a per-query top-k collector (faiss: a min-heap; here torch.topk over
[carried k | block], which keeps the same set) — fp32 throughout, results sorted best first with
the id map applied last.  Tie order in faiss is heap order, i.e. unspecified; this port breaks ties
by (score desc, id asc) so it can be checked against oracle.mips.  PARITY UNPINNED by the reference
(it holds no tests or vectors for this path); pinned here against oracle.mips.mips_topk, which is
itself pinned to the reference's DistributedBruteForceIndex outputs (tests/golden).
"""
import os

import numpy as np
import torch

DATABASE_BLOCK = 1024          # faiss distance_compute_blas_database_bs
PANEL_BLOCKS = 64              # blocks folded per topk call (implementation detail, same result)


def flat_ip_search(rows_f32, ids, queries_f32, k, threads=None):
    """rows_f32 [n, d] float32 torch CPU tensor (the 'added' vectors), ids int64 [n] or None,
    queries_f32 [nq, d] float32.  Returns numpy (distances float32 [nq,k], ids int64 [nq,k])."""
    if threads:
        torch.set_num_threads(int(threads))
    n = rows_f32.shape[0]
    nq = queries_f32.shape[0]
    best_s = torch.full((nq, k), float("-inf"))
    best_r = torch.full((nq, k), -1, dtype=torch.int64)
    step = DATABASE_BLOCK * PANEL_BLOCKS
    for r0 in range(0, n, step):
        panel = queries_f32 @ rows_f32[r0:r0 + step].T            # SGEMM
        kk = min(k, panel.shape[1])
        ps, pr = torch.topk(panel, kk, dim=1)
        cs = torch.cat([best_s, ps], dim=1)
        cr = torch.cat([best_r, pr + r0], dim=1)
        # deterministic (score desc, row asc): sort rows first, then a stable sort on the scores
        order = torch.argsort(cr, dim=1, stable=True)
        cs, cr = torch.gather(cs, 1, order), torch.gather(cr, 1, order)
        order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_r = torch.gather(cs, 1, order), torch.gather(cr, 1, order)
    out_i = best_r.numpy().copy()
    live = out_i >= 0
    if ids is not None:
        out_i[live] = np.asarray(ids, dtype=np.int64)[out_i[live]]
    return best_s.numpy(), out_i


def host_threads():
    return os.cpu_count() or 1
