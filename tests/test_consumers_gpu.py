"""The second consumers of the MIPS index and the BERT towers, driven through the sm_100a kernels
(SURVEY §8 f-2 / f-4): the recall evaluator over B200FaissMIPSIndex at k = 100, the in-batch-negative
retriever step over the DualEncoder towers (forward + backward), and an index refresh from the flat
evidence store.  CPU-side logic of the same modules is covered in test_recall.py, test_dense_retriever.py
and test_host_logic.py; here the engine underneath is the real one."""
import math

import numpy as np
import pytest
import torch

from helpers import TINY, assert_ids_equal_outside_ties, seeded_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_recall_evaluator_top100_on_the_gpu_index():
    """tasks/openqa/dense_retriever/evaluation/evaluate.py:123-168 on B200FaissMIPSIndex: k = 100 goes
    through the exact row-range refinement over the real scan kernel; the rank at which each question's
    answer-bearing passage appears, and hence every top-k accuracy, must equal the oracle's."""
    from emdr2_b200 import recall
    from emdr2_b200.index import B200FaissMIPSIndex
    from oracle import mips as oracle
    rng = np.random.RandomState(2)
    n, d, nq, k = 20000, 64, 12, 100
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    ids = np.arange(1, n + 1, dtype=np.int64)
    gold = rng.choice(n, size=nq, replace=False)
    # half of the questions sit near their passage, half are only weakly related (answer deep in the list)
    pull = np.where(np.arange(nq) % 2 == 0, 2.0, 0.35)[:, None]
    queries = (rows[gold].astype(np.float32) * pull + rng.randint(-64, 65, size=(nq, d)) / 64).astype(np.float16)
    id2text = {int(i): ("filler text number %d" % i, "title") for i in ids}
    answers = []
    for qi, row in enumerate(gold):
        id2text[int(ids[row])] = ("the answer is Zürich-%d indeed" % qi, "title")
        answers.append(["zürich-%d" % qi])
    index = B200FaissMIPSIndex(d, device=DEV)
    index.add_arrays(ids, rows)
    ev = recall.RecallEvaluator(index, id2text, topk_retrievals=k)
    acc, stats, closest = ev.evaluate(torch.from_numpy(queries), answers)
    want_s, want_i, ties = oracle.mips_topk(rows, queries, k, ids=ids, want_ties=True)
    got_i = np.array([c[0] for c in closest])
    assert_ids_equal_outside_ties(got_i, want_i, ties)
    assert np.allclose(np.array([c[1] for c in closest], dtype=np.float32), want_s, rtol=1e-5, atol=1e-6)   # fp32 accumulation order
    for kk in (1, 5, 10, 20, 50, 100):
        want = sum(bool((want_i[qi, :kk] == ids[gold[qi]]).any()) for qi in range(nq)) / nq
        assert acc[kk] == want, (kk, acc[kk], want)
    assert 0 < acc[1] <= acc[100]


@pytest.mark.parametrize("with_neg", [False, True])
def test_in_batch_negative_step_through_the_towers(with_neg):
    """train_dense_retriever.py:89-196 with both BERT towers on the library kernels: loss, hit count and
    parameter gradients against fp32 autograd of the block oracle on the same fp16-rounded weights.
    Tolerance: loss 2e-3 absolute; gradients 4e-2 relative Frobenius (fp16 activations, two towers)."""
    from emdr2_b200 import dense_retriever as dr
    from emdr2_b200.model import DualEncoder
    from oracle import blocks as ob
    dtype = torch.float16
    model = DualEncoder(dict(TINY, dtype=dtype)).to(DEV)
    w32 = {}
    with torch.no_grad():
        for name, p in model.named_parameters():
            w = seeded_weights(name, tuple(p.shape)).to(dtype)
            p.copy_(w)
            w32[name] = w.float().requires_grad_(True)
    rng = np.random.RandomState(8)
    b, s, neg = 4, 32, 3

    def batch(rows):
        ids = rng.randint(1, TINY["vocab"], size=(rows, s)).astype(np.int64)
        for i, n in enumerate(rng.randint(6, s + 1, size=rows)):
            ids[i, n:] = 0
        return torch.from_numpy(ids), torch.from_numpy((rng.rand(rows, s) < 0.5).astype(np.int64) * (ids > 0))

    q, qt = batch(b)
    c, ct = batch(b)
    ng, ngt = batch(neg) if with_neg else (None, None)
    loss, stats = dr.forward_step(model, q.to(DEV), qt.to(DEV), c.to(DEV), ct.to(DEV),
                                  None if ng is None else ng.to(DEV), None if ngt is None else ngt.to(DEV),
                                  hidden_size=TINY["hidden"])
    loss.backward()

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in w32.items() if k.startswith(prefix)}

    call = torch.cat([c, ng]) if with_neg else c
    callt = torch.cat([ct, ngt]) if with_neg else ct
    oq = ob.bert_pooled(q, qt, sub("query_model."), TINY["heads"], TINY["layers"])
    oc = ob.bert_pooled(call, callt, sub("context_model."), TINY["heads"], TINY["layers"])
    scores = oq @ oc.T / math.sqrt(TINY["hidden"])
    lp = torch.log_softmax(scores, dim=1)
    want = torch.nn.functional.nll_loss(lp, torch.arange(b))
    want.backward()
    assert abs(loss.item() - want.item()) < 2e-3
    assert int(stats["correct_prediction_count"].item()) == int((lp.argmax(1) == torch.arange(b)).sum())
    errs = []
    for name, p in model.named_parameters():
        g = w32[name].grad
        if g is None or float(g.norm()) < 1e-5:     # e.g. the context tower's last bias: the softmax over the
            continue                                 # contexts is invariant to it, its gradient is rounding noise
        errs.append((((p.grad.float().cpu() - g).norm() / g.norm()).item(), name, float(g.norm()),
                     float(p.grad.float().norm())))
    errs.sort(reverse=True)
    assert errs[0][0] < 4e-2, errs[:6]


def test_index_refresh_from_the_flat_store_on_the_gpu(tmp_path):
    """emdr2_index.py:232-239 (`update_index`) over the flat evidence store: rows are memory-mapped and go
    straight to HBM; results equal the oracle before and after the refresh."""
    from emdr2_b200.index import B200BruteForceIndex
    from emdr2_b200.store import EvidenceStore
    from oracle import mips as oracle
    path = str(tmp_path / "ev.pkl")
    rng = np.random.RandomState(4)
    n, d = 5000, 128
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(8, d)) / 64).astype(np.float16)
    ids = np.arange(1, n + 1, dtype=np.int64)
    st = EvidenceStore(path, load_from_path=False, rank=0, format="flat")
    st.add_block_data(ids, rows)
    st.save_shard()
    st.merge_shards_and_save()
    index = B200BruteForceIndex(d, EvidenceStore(path), device=DEV)
    s0, i0 = index.search(torch.from_numpy(queries).to(DEV), 20)
    w0 = oracle.mips_topk(rows, queries, 20, ids=ids)
    assert np.array_equal(i0.cpu().numpy(), w0[1]) and np.array_equal(s0.cpu().numpy(), w0[0])
    st2 = EvidenceStore(path, load_from_path=False, rank=0, format="flat")      # the refreshed embeddings
    st2.add_block_data(ids, rows[::-1])
    st2.save_shard()
    st2.merge_shards_and_save()
    index.update_index()
    s1, i1 = index.search(torch.from_numpy(queries).to(DEV), 20)
    w1 = oracle.mips_topk(np.ascontiguousarray(rows[::-1]), queries, 20, ids=ids)
    assert np.array_equal(i1.cpu().numpy(), w1[1]) and np.array_equal(s1.cpu().numpy(), w1[0])
    index.reset_index()
    assert torch.equal(index.search(torch.from_numpy(queries).to(DEV), 20)[1], i1)


def test_concurrent_index_refresh_on_the_gpu_searches_see_old_or_new_never_a_mix():
    """BASELINE config 5 on one GPU (async_indexer.ConcurrentShardRefresher): a worker thread re-encodes the shard
    with a frozen copy of a real (tiny) context tower on a side CUDA stream into a standby buffer while this
    thread keeps searching the live shard through the scan kernel; every answer equals the OLD index's or — after
    the swap — the NEW index's, and the new shard holds exactly the frozen tower's embeddings."""
    import threading
    from emdr2_b200.async_indexer import ConcurrentShardRefresher
    from emdr2_b200.blocks import BertTower
    from emdr2_b200.index import B200BruteForceIndex
    dtype = torch.float16
    tower = BertTower(dict(TINY, dtype=dtype)).to(DEV).eval()
    with torch.no_grad():
        for name, p in tower.named_parameters():
            p.copy_(seeded_weights(name, tuple(p.shape)).to(dtype))
    n, d, k, s = 1536, TINY["hidden"], 10, 32
    rng = np.random.RandomState(6)
    tokens = rng.randint(1, TINY["vocab"], size=(n, s)).astype(np.int64)
    for i, ln in enumerate(rng.randint(5, s + 1, size=n)):
        tokens[i, ln:] = 0
    tok_t = torch.from_numpy(tokens)

    def encode(model):
        with torch.no_grad():
            return torch.cat([model(tok_t[a:a + 128].to(DEV), None, torch.zeros(128, s, dtype=torch.int64, device=DEV))
                              for a in range(0, n, 128)]).to(dtype)

    old_rows = encode(tower)
    index = B200BruteForceIndex(d, device=DEV)
    index.add_local_shard(None, old_rows.clone(), num_rows=n, row_lo=0)
    queries = torch.randn(6, d, generator=torch.Generator().manual_seed(1)).to(dtype).to(DEV)
    want_old = index.search(queries, k)
    with torch.no_grad():                                     # "training" changes the context tower
        for p in tower.parameters():
            p.mul_(0.9)

    def make_batches():
        for a in range(0, n, 128):
            yield (torch.arange(a + 1, a + 129), tok_t[a:a + 128].pin_memory(),
                   torch.zeros(128, s, dtype=torch.int64).pin_memory())

    refresher = ConcurrentShardRefresher(index, tower, make_batches)
    refresher.start()
    with torch.no_grad():                                     # later steps: must not leak into the running refresh
        for p in tower.parameters():
            p.mul_(0.5)
    seen = []
    swapped = False
    for _ in range(20000):
        sc, ids = index.search(queries, k)
        seen.append((sc.clone(), ids.clone()))
        if refresher.maybe_swap():
            swapped = True
            break
    assert swapped and refresher.rows_done == n
    new_rows = encode(refresher.tower)
    assert torch.equal(index.evidence_embeds, new_rows)       # the frozen snapshot (x0.9), not the live weights (x0.45)
    assert not torch.equal(new_rows, old_rows)
    want_new = index.search(queries, k)
    fresh = B200BruteForceIndex(d, device=DEV)
    fresh.add_local_shard(None, new_rows.clone(), num_rows=n, row_lo=0)
    check = fresh.search(queries, k)
    assert torch.equal(want_new[1], check[1]) and torch.equal(want_new[0], check[0])
    for sc, ids in seen:                                      # everything before the swap came from the old shard
        assert torch.equal(ids, want_old[1]) and torch.equal(sc, want_old[0])
    assert len(seen) >= 1
