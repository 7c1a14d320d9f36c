"""Formatter (integer work: must be element-for-element identical) and loss arithmetic against
cases produced by the reference's own functions (tests/golden/make_formatter_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from emdr2_b200 import formatter
from helpers import GOLDEN


def cases():
    with open(os.path.join(GOLDEN, "formatter_ref.json")) as f:
        return json.load(f)


def test_t5_and_bert_formats_are_identical_to_the_reference():
    for c in cases():
        ext = formatter.query_extended_context_t5_format(c["query"], c["title"], c["docs"], c["main"],
                                                         c["max_len"], 3, 0)
        one = formatter.query_single_context_t5_format(c["query"], c["title"], c["docs"][c["main"]],
                                                       c["max_len"], 3, 0)
        ids, types, mask = formatter.context_bert_format(c["title"] + [3] + c["docs"][c["main"]],
                                                         c["max_len"], 2, 3, 0)
        assert ext == c["extended"], c
        assert one == c["single"], c
        assert [list(ids), list(types), mask.tolist()] == c["bert"], c


@pytest.mark.skipif(not os.path.isdir("/root/reference/megatron"), reason="reference not mounted")
def test_formats_against_the_live_reference_on_fresh_random_cases():
    import sys
    sys.path.insert(0, GOLDEN)
    import make_formatter_golden as mk
    mk.make_mips_golden.install_shims()
    from megatron.model.emdr2_model import query_extended_context_t5_format, query_single_context_t5_format
    rng = np.random.RandomState(2024)
    for _ in range(500):
        c = mk.random_case(rng)
        assert formatter.query_extended_context_t5_format(c["query"], c["title"], c["docs"], c["main"],
                                                          c["max_len"], 3, 0) == \
            query_extended_context_t5_format(list(c["query"]), list(c["title"]), [list(d) for d in c["docs"]],
                                             c["main"], c["max_len"], 3, 0)
        assert formatter.query_single_context_t5_format(c["query"], c["title"], c["docs"][c["main"]],
                                                        c["max_len"], 3, 0) == \
            query_single_context_t5_format(list(c["query"]), list(c["title"]), list(c["docs"][c["main"]]),
                                           c["max_len"], 3, 0)


def test_array_fast_path_is_identical_to_the_reference_on_every_golden_case():
    """postprocess_arrays assembles rows from array slices; it must reproduce the reference's three
    formats element for element (including every neighbour-fill branch)."""
    cs = cases()
    for c in cs:
        if len(c["query"]) + len(c["title"]) + 2 > c["max_len"]:
            continue                        # the reference itself overflows the row here
        data = [([1], [([np.array(d, dtype=np.int64) for d in c["docs"]], c["main"], np.array(c["title"]))])]
        q = c["query"] + [0] * 4
        ctx, typ, ext, one = formatter.postprocess_arrays([-1], [q], [len(c["query"])], data, 1, c["max_len"],
                                                          c["max_len"], 2, 3, 0)
        assert ext[0].tolist() == c["extended"], c
        assert one[0].tolist() == c["single"], c
        assert ctx[0, 0].tolist() == c["bert"][0] and typ[0, 0].tolist() == c["bert"][1], c


def test_postprocess_shapes_filtering_and_rows():
    cs = cases()[:12]
    b, k = 3, 3
    topk_data, qt5, qlen, uids = [], [], [], [-1, -2, 7]
    for bi in range(b):
        ids, texts = [], []
        for j in range(k + 1):
            c = cs[bi * 4 + j]
            ids.append(7 if (bi == 2 and j == 1) else 100 + bi * 10 + j)
            texts.append((c["docs"], c["main"], c["title"]))
        topk_data.append((ids, texts))
        q = cs[bi]["query"]
        qt5.append(q + [0] * (16 - len(q)))
        qlen.append(len(q))
    ctx, typ, ext, one = formatter.postprocess_arrays(uids, qt5, qlen, topk_data, k, 32, 64, 2, 3, 0)
    assert ctx.shape == (b, k, 32) and typ.shape == (b, k, 32) and ext.shape == (b * k, 64) == one.shape
    assert (typ == 0).all()
    # query 2 originates from passage id 7 (rank 1): that passage is skipped, the (k+1)-th is used
    kept = [0, 2, 3]
    for slot, j in enumerate(kept):
        c = cs[2 * 4 + j]
        want = formatter.query_extended_context_t5_format(cs[2]["query"], c["title"], c["docs"], c["main"], 64, 3, 0)
        assert ext[2 * k + slot].tolist() == want
    c0 = cs[0]
    want_bert, _, _ = formatter.context_bert_format(c0["title"] + [3] + c0["docs"][c0["main"]], 32, 2, 3, 0)
    assert ctx[0, 0].tolist() == list(want_bert)
    t = formatter.postprocess(torch.tensor(uids), torch.tensor(qt5), torch.tensor(qlen), topk_data, k, 32, 64,
                              2, 3, 0, device="cpu")
    assert all(x.dtype == torch.int64 for x in t) and torch.equal(t[2], torch.from_numpy(ext))
    with pytest.raises(ValueError):
        formatter.postprocess_arrays(uids, qt5, qlen, [(i[:2], x[:2]) for i, x in topk_data], k, 32, 64, 2, 3, 0)


def test_native_formatter_is_identical_to_the_reference_on_every_golden_case():
    """emdr2_format_passages (C, host) against the fixtures the reference's own functions produced."""
    n = 0
    for c in cases():
        if len(c["query"]) + len(c["title"]) + 2 > c["max_len"]:
            continue
        data = [([1], [([np.array(d, dtype=np.int64) for d in c["docs"]], c["main"], np.array(c["title"], dtype=np.int64))])]
        q = c["query"] + [0] * 4
        (ctx, typ, ext, one), lens = formatter.format_passages_native([-1], [q], [len(c["query"])], data, 1,
                                                                      c["max_len"], c["max_len"], 2, 3, 0)
        assert ext[0].tolist() == c["extended"], c
        assert one[0].tolist() == c["single"], c
        assert ctx[0, 0].tolist() == c["bert"][0] and typ[0, 0].tolist() == c["bert"][1], c
        n += 1
    assert n > 100


def _random_batch(rng, b, k, seq_ret, seq, extra, lists):
    uids, qt5, qlen, topk_data = [], [], [], []
    for bi in range(b):
        n = int(rng.randint(0, 12))
        q = rng.randint(1, 50, size=n).tolist()
        qt5.append(q + [0] * (16 - n))
        qlen.append(n)
        ids = rng.choice(np.arange(1, 400), size=k + extra, replace=False).tolist()
        uid = ids[int(rng.randint(0, k + extra))] if (extra and rng.rand() < 0.5) else -(bi + 1)
        uids.append(uid)
        texts = []
        for _ in ids:
            nd = int(rng.randint(1, 4))
            # token 0 (= pad id) appears inside passages now and then: exercises the length rule
            docs = [rng.randint(0, 60, size=int(rng.randint(0, 40))).astype(np.int64) for _ in range(nd)]
            main = [0, 1, -1][int(rng.randint(0, 3))] if nd > 1 else [0, -1][int(rng.randint(0, 2))]
            if main == 1 and nd < 2:
                main = 0
            title = rng.randint(1, 60, size=int(rng.randint(0, 6))).astype(np.int64)
            if lists:
                docs, title = [d.tolist() for d in docs], title.tolist()
            texts.append((docs, main, title))
        topk_data.append((ids, texts))
    return uids, qt5, qlen, topk_data


@pytest.mark.parametrize("lists", [False, True])
def test_native_formatter_matches_the_python_restatement_on_random_batches(lists):
    rng = np.random.RandomState(7)
    for trial in range(60):
        b, k = int(rng.randint(1, 5)), int(rng.randint(1, 6))
        seq_ret, seq = int(rng.randint(20, 60)), int(rng.randint(32, 120))
        uids, qt5, qlen, data = _random_batch(rng, b, k, seq_ret, seq, extra=1, lists=lists)
        want = formatter.postprocess_arrays(uids, qt5, qlen, data, k, seq_ret, seq, 2, 3, 0)
        got, lens = formatter.format_passages_native(uids, qt5, qlen, data, k, seq_ret, seq, 2, 3, 0)
        for w, g in zip(want, got):
            assert w.shape == g.shape and np.array_equal(w, g), trial

        def longest(a):
            a2 = a.reshape(-1, a.shape[-1])
            return int(((a2 != 0) * np.arange(1, a2.shape[1] + 1)).max()) if a2.size else 0
        assert lens == (longest(want[0]), longest(want[2]), longest(want[3]))
        for per_row, w in zip(lens.rows, (want[0], want[2], want[3])):
            w2 = w.reshape(-1, w.shape[-1])
            assert np.array_equal(per_row, ((w2 != 0) * np.arange(1, w2.shape[1] + 1)).max(axis=1))


def test_native_formatter_errors_and_staging_reuse():
    rng = np.random.RandomState(3)
    uids, qt5, qlen, data = _random_batch(rng, 2, 3, 32, 64, extra=0, lists=False)
    with pytest.raises(ValueError, match="kept 2 of 3"):
        formatter.format_passages_native(uids, qt5, qlen, [(i[:2], x[:2]) for i, x in data], 3, 32, 64, 2, 3, 0)
    long_q = [list(range(1, 17))] * 2
    with pytest.raises(ValueError, match="do not fit"):
        formatter.format_passages_native(uids, long_q, [16, 16], data, 3, 32, 18, 2, 3, 0)
    with pytest.raises(ValueError):
        formatter.postprocess_arrays(uids, long_q, [16, 16], data, 3, 32, 18, 2, 3, 0)
    # postprocess() cycles through two staging blocks: results of consecutive calls stay distinct
    first = formatter.postprocess(uids, qt5, qlen, data, 3, 32, 64, 2, 3, 0, device="cpu")
    uids2, qt52, qlen2, data2 = _random_batch(rng, 2, 3, 32, 64, extra=0, lists=False)
    second = formatter.postprocess(uids2, qt52, qlen2, data2, 3, 32, 64, 2, 3, 0, device="cpu")
    third, lens = formatter.postprocess(uids, qt5, qlen, data, 3, 32, 64, 2, 3, 0, device="cpu", return_lengths=True)
    want = formatter.postprocess_arrays(uids, qt5, qlen, data, 3, 32, 64, 2, 3, 0)
    for a, b_, w in zip(first, third, want):
        assert torch.equal(a, b_) and np.array_equal(a.numpy(), w)
    assert not torch.equal(first[2], second[2])
    assert len(lens) == 3


def _flat_corpus(rng, n_articles=120):
    from emdr2_b200.titlemap import TitleDocMap
    pairs, passages, titles, did = [], [], [], 1
    for a in range(n_articles):
        title = rng.randint(1, 60, size=int(rng.randint(0, 6))).astype(np.int64)
        for _ in range(int(rng.randint(1, 6))):
            pairs.append((did, "t%d" % a))
            passages.append(rng.randint(0, 60, size=int(rng.randint(1, 40))).astype(np.int64))
            titles.append(title)
            did += 1
    return TitleDocMap(pairs=pairs), passages, titles


class _FixedSearchRetriever(object):
    """B200EvidenceRetriever with the GPU search replaced by fixed ids (host-tail tests on CPU)."""

    def __new__(cls, topk_ids, passages_map, title_map, titlemap):
        from emdr2_b200.retriever import B200EvidenceRetriever

        class R(B200EvidenceRetriever):
            def __init__(self):                      # no index: only the tail is under test
                self.topk = topk_ids.shape[1]
                self.passages_map, self.title_map, self.wikititledocmap = passages_map, title_map, titlemap
                self.mips_index = type("I", (), {"rank": 0, "world": 1})()

            def search_all(self, query_tensor):
                ids = torch.from_numpy(topk_ids)
                return torch.zeros(ids.shape, dtype=torch.float32), ids

        return R()


@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.uint16])
def test_packed_retrieval_tail_and_flat_formatter_equal_the_nested_path(dtype):
    """FlatTokenStore + NeighbourTable + emdr2_format_passages_flat (tokens read in place, any width)
    against the per-passage Python tail + postprocess_arrays, on a corpus with 1..5-passage articles."""
    from emdr2_b200.titlemap import NeighbourTable
    from emdr2_b200.tokens import FlatTokenStore
    rng = np.random.RandomState(21)
    titlemap, passages, titles = _flat_corpus(rng)
    n = len(passages)
    flat_p, flat_t = FlatTokenStore.from_arrays(passages, dtype=dtype), FlatTokenStore.from_arrays(titles, dtype=dtype)
    assert len(flat_p) == n and np.array_equal(flat_p[7], passages[7]) and np.array_equal(flat_t[-1], titles[-1])
    b, k = 4, 6
    topk_ids = np.stack([rng.choice(np.arange(1, n + 1), size=k + 1, replace=False) for _ in range(b)]).astype(np.int64)
    uids = [-1, int(topk_ids[1, 2]), -3, int(topk_ids[3, 0])]          # two questions drop their own passage
    qt5 = [rng.randint(1, 50, size=8).tolist() + [0] * 8 for _ in range(b)]
    qlen = [8, 5, 0, 3]
    q = torch.zeros(b, 4)
    slow = _FixedSearchRetriever(topk_ids, passages, titles, titlemap)
    assert not slow.supports_packed
    nested, _ = slow.get_topk(q, as_arrays=True)
    fast = _FixedSearchRetriever(topk_ids, flat_p, flat_t, NeighbourTable(titlemap))
    assert fast.supports_packed
    packed, dist = fast.get_topk(q, as_packed=True)
    assert dist.dtype == torch.float16 and len(packed) == b
    # the packed form describes exactly the nested one
    for (ids_a, texts_a), (ids_b, texts_b) in zip(nested, packed.to_nested()):
        assert list(ids_a) == list(ids_b)
        for (docs_a, main_a, title_a), (docs_b, main_b, title_b) in zip(texts_a, texts_b):
            assert main_a == main_b and np.array_equal(title_a, title_b) and len(docs_a) == len(docs_b)
            assert all(np.array_equal(x, y) for x, y in zip(docs_a, docs_b))
    want = formatter.postprocess_arrays(uids, qt5, qlen, nested, k, 40, 96, 2, 3, 0)
    got, lens = formatter.format_passages_native(uids, qt5, qlen, packed, k, 40, 96, 2, 3, 0)
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    via_nested, lens2 = formatter.format_passages_native(uids, qt5, qlen, nested, k, 40, 96, 2, 3, 0)
    assert lens == lens2 and all(np.array_equal(x, y) for x, y in zip(lens.rows, lens2.rows))
    t = formatter.postprocess(uids, qt5, qlen, packed, k, 40, 96, 2, 3, 0, device="cpu")
    assert torch.equal(t[2], torch.from_numpy(want[2])) and torch.equal(t[0], torch.from_numpy(want[0]))


def test_flat_store_checks():
    from emdr2_b200.tokens import FlatTokenStore
    with pytest.raises(TypeError):
        FlatTokenStore(np.zeros(4, dtype=np.float32), [0, 4])
    with pytest.raises(ValueError):
        FlatTokenStore(np.zeros(4, dtype=np.int64), [0, 5])
    st = FlatTokenStore(np.arange(10, dtype=np.int32), [0, 3, 3, 10])
    assert [len(st[i]) for i in range(3)] == [3, 0, 7] and st.token_bytes == 4
    off, ln = st.spans(np.array([[2, -1], [0, 1]]))
    assert off.tolist() == [[3, 0], [0, 3]] and ln.tolist() == [[7, 0], [3, 0]]
    with pytest.raises(IndexError):
        st.spans([3])


def _loss_golden():
    with np.load(os.path.join(GOLDEN, "losses_ref.npz")) as z:
        return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k] for k in z.files}


def test_loss_arithmetic_after_the_gather_matches_reference():
    from emdr2_b200 import losses
    g = _loss_golden()
    labels = g["labels"].masked_fill(~g["loss_mask"].bool(), 0)
    lp = torch.log_softmax(g["logits"], dim=-1)
    gold = torch.gather(lp, -1, labels[:, None, :, None].expand(-1, lp.shape[1], -1, 1)).squeeze(-1)
    lm, ru, nb = losses.loss_and_retriever_utility_from_gold(gold, g["topk_log_probs"], labels, g["loss_mask"],
                                                             int(g["eos_id"]))
    assert np.isclose(float(lm), float(g["lm_loss"]), rtol=1e-6)
    assert np.isclose(float(ru), float(g["retriever_utility"]), rtol=1e-5)
    assert np.isclose(float(nb), float(g["null_block_lm_loss"]), rtol=1e-6)
    kl = losses.kl_div_retriever_from_gold(gold, g["topk_log_probs"], g["loss_mask"])
    assert np.isclose(float(kl), float(g["kl"]), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_losses_on_gpu_match_reference_within_logit_rounding(dtype):
    """Full path through the fused log-prob kernel.  The only difference from the reference's fp32
    numbers is the rounding of the logits to 16 bits on input: |dloss| <= 3e-2 (bf16) / 4e-3 (fp16)."""
    from emdr2_b200 import losses
    g = _loss_golden()
    dev = "cuda:0"
    logits = g["logits"].to(dev).to(dtype)
    args = (g["topk_log_probs"].to(dev), g["labels"].to(dev), g["loss_mask"].to(dev))
    lm, ru, nb = losses.get_loss_and_retriever_utility(logits, *args, int(g["eos_id"]))
    tol = 3e-2 if dtype == torch.bfloat16 else 4e-3
    assert abs(float(lm) - float(g["lm_loss"])) <= tol
    assert abs(float(ru) - float(g["retriever_utility"])) <= tol
    assert abs(float(nb) - float(g["null_block_lm_loss"])) <= tol
    kl = losses.get_kl_div_retriever(logits, *args)
    assert abs(float(kl) - float(g["kl"])) <= tol
    # exactness of the kernel itself: same rounded logits through torch in fp32
    ref = torch.log_softmax(logits.float(), dim=-1)
    labels = g["labels"].to(dev).masked_fill(~g["loss_mask"].to(dev).bool(), 0)
    want = torch.gather(ref, -1, labels[:, None, :, None].expand(-1, ref.shape[1], -1, 1)).squeeze(-1)
    got, _ = losses.gold_log_probs(logits, g["labels"].to(dev), g["loss_mask"].to(dev))
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-5)
    ce = losses.reader_cross_entropy(logits[:, 0], labels, g["loss_mask"].to(dev))
    want_ce = (torch.nn.functional.cross_entropy(logits[:, 0].float().reshape(-1, logits.shape[-1]),
                                                 labels.reshape(-1), reduction="none", ignore_index=0)
               * g["loss_mask"].to(dev).reshape(-1)).sum() / g["loss_mask"].sum()
    assert abs(float(ce) - float(want_ce)) <= 1e-4


@pytest.mark.gpu
def test_token_logprob_large_vocab():
    from emdr2_b200.ops import token_logprob
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    logits = (torch.randn(300, 30720, generator=gen, device="cuda:0") * 3).to(torch.bfloat16)
    labels = torch.randint(0, 30720, (300,), generator=gen, device="cuda:0")
    labels[5] = -1
    lp, lse = token_logprob(logits, labels)
    want_lse = torch.logsumexp(logits.float(), dim=-1)
    want = logits.float().gather(1, labels.clamp(min=0)[:, None]).squeeze(1) - want_lse
    want[5] = 0
    assert torch.allclose(lse, want_lse, rtol=1e-5, atol=1e-4)
    assert torch.allclose(lp, want, rtol=1e-5, atol=1e-4)


@pytest.mark.skipif(not os.path.isdir("/root/reference/megatron"), reason="reference not mounted")
@pytest.mark.parametrize("dtype", [np.uint16, np.int32])
def test_flat_store_wraps_the_reference_indexed_dataset_without_copying(tmp_path, dtype):
    """FlatTokenStore.from_indexed_dataset over a real MMapIndexedDataset written by the reference's own
    builder (megatron/data/indexed_dataset.py): same documents, token buffer shared with the mmap."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_mips_golden
    make_mips_golden.install_shims()
    from megatron.data import indexed_dataset as ref
    from emdr2_b200.tokens import FlatTokenStore
    rng = np.random.RandomState(3)
    docs = [rng.randint(1, 30000, size=int(rng.randint(1, 50))) for _ in range(40)]
    prefix = str(tmp_path / "evidence_text")
    builder = ref.MMapIndexedDatasetBuilder(ref.data_file_path(prefix), dtype=dtype)
    for d in docs:
        builder.add_item(torch.from_numpy(d.astype(np.int64)))
        builder.end_document()
    builder.finalize(ref.index_file_path(prefix))
    ds = ref.MMapIndexedDataset(prefix)
    store = FlatTokenStore.from_indexed_dataset(ds)
    assert len(store) == len(docs) == len(ds) and store.token_bytes == np.dtype(dtype).itemsize
    for i, d in enumerate(docs):
        assert np.array_equal(store[i], d) and np.array_equal(ds[i], store[i])
    assert np.shares_memory(store.tokens, np.frombuffer(ds._bin_buffer, dtype=dtype))
