"""In-batch-negative loss of the dual encoder (emdr2_b200/dense_retriever.py) on CPU: a world-size-2 Gloo
run must give every rank the loss, hit count and LOCAL gradients of a single-process computation over the
concatenated batch (reference tasks/openqa/dense_retriever/train_dense_retriever.py:131-190)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from emdr2_b200 import dense_retriever as dr

D, B, NEG = 16, 3, 2


def _data(rank, with_neg):
    g = torch.Generator().manual_seed(50 + rank)
    q = torch.randn(B, D, generator=g)
    c = torch.randn(B + (NEG if with_neg else 0), D, generator=g)
    c[:B] += 2.0 * q                      # positives resemble their questions
    return q, c


def _single_process(world, with_neg, scaling):
    qs, cs = zip(*[_data(r, with_neg) for r in range(world)])
    qs = [q.clone().requires_grad_(True) for q in qs]
    cs = [c.clone().requires_grad_(True) for c in cs]
    scores = torch.cat(qs) @ torch.cat(cs).T
    if scaling:
        scores = scores / math.sqrt(D)
    n_ctx = cs[0].shape[0]
    labels = torch.tensor([r * n_ctx + i for r in range(world) for i in range(B)])
    lp = F.log_softmax(scores, dim=1)
    nll = F.nll_loss(lp, labels)
    (nll * world).backward()
    return nll.item(), int((lp.argmax(1) == labels).sum()), [q.grad for q in qs], [c.grad for c in cs]


def _worker(rank, world, port, out_dir, with_neg, scaling):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    q, c = _data(rank, with_neg)
    q.requires_grad_(True)
    c.requires_grad_(True)
    loss, stats = dr.in_batch_negative_loss(q, c, D, retriever_score_scaling=scaling, train_with_neg=with_neg)
    loss.backward()
    # the reference averages gradients over the data-parallel group afterwards; here: sum of the per-rank
    # graphs = gradient of (world * nll), which each rank holds for its own slice only
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), loss=loss.item(), nll=stats["lm loss"].item(),
             correct=stats["correct_prediction_count"].item(), gq=q.grad.numpy(), gc=c.grad.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("with_neg,scaling", [(False, True), (True, False)])
def test_world2_gloo_equals_the_single_process_batch(tmp_path, with_neg, scaling):
    import torch.multiprocessing as mp
    world = 2
    for attempt in range(2):
        try:
            mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), with_neg, scaling), nprocs=world, join=True)
            break
        except Exception:
            if attempt:
                raise
    want_nll, want_correct, want_gq, want_gc = _single_process(world, with_neg, scaling)
    for rank in range(world):
        r = np.load(str(tmp_path / ("r%d.npz" % rank)))
        assert abs(float(r["nll"]) - want_nll) < 1e-5 and abs(float(r["loss"]) - world * want_nll) < 1e-5
        assert int(r["correct"]) == want_correct
        # each rank's graph covers all rows of the score matrix but only its own slice of Q and C, so the
        # gradient it sees for that slice is the full-batch gradient
        assert np.allclose(r["gq"], want_gq[rank].numpy(), atol=1e-5)
        assert np.allclose(r["gc"], want_gc[rank].numpy(), atol=1e-5)


def test_labels_and_argument_checks():
    assert dr.in_batch_labels(4, 8, 3, True).tolist() == [0, 1, 2, 3, 8, 9, 10, 11, 16, 17, 18, 19]   # :164-166
    assert dr.in_batch_labels(2, 2, 2, False).tolist() == [0, 1, 2, 3]
    q, c = torch.zeros(3, 4), torch.zeros(2, 4)
    with pytest.raises(ValueError):
        dr.in_batch_negative_loss(q, c, 4)
    with pytest.raises(ValueError):
        dr.in_batch_negative_loss(q, c, 4, train_with_neg=True)
    loss, stats = dr.in_batch_negative_loss(torch.eye(3), torch.eye(3) * 5, 3)
    assert stats["correct_prediction_count"].item() == 3 and loss.item() < 0.3
