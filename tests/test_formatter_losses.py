"""Formatter (integer work: must be element-for-element identical) and loss arithmetic against
cases produced by the reference's own functions (tests/golden/make_formatter_golden.py)."""
import json
import os

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
