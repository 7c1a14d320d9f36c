"""End-to-end retrieve-and-read forward (EMDR2Model.forward order, emdr2_model.py:87-214) on the GPU
against the composed CPU oracle (oracle.blocks + oracle.mips, each pinned to the reference) on a
tiny corpus: index build with the context tower, MIPS retrieval, formatting, context tower,
fresh scores, FiD reader, one-context pass, losses."""
import math

import numpy as np
import pytest
import torch

from helpers import TINY, seeded_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CLS, SEP, PAD = 2, 3, 0
TOPK, S_RET, S, L = 4, 32, 64, 8


def corpus(rng, n_docs=300):
    """Articles of 1-4 consecutive passages sharing a title (what the evidence TSV looks like)."""
    passages, titles, pairs, doc_id, article = [], [], [], 1, 0
    while doc_id <= n_docs:
        title = rng.randint(5, TINY["vocab"], size=int(rng.randint(1, 4))).tolist()
        article += 1
        for _ in range(int(rng.randint(1, 5))):
            if doc_id > n_docs:
                break
            passages.append(np.array(rng.randint(5, TINY["vocab"], size=int(rng.randint(5, 20))), dtype=np.int64))
            titles.append(np.array(title, dtype=np.int64))
            pairs.append((doc_id, "article-%d" % article))
            doc_id += 1
    return passages, titles, pairs


def build_model(dtype, retriever, update_retriever):
    from emdr2_b200.model import EMDR2Model
    cfg = dict(TINY, dtype=dtype, max_pos=S)
    settings = dict(topk_retrievals=TOPK, seq_length=S, seq_length_ret=S_RET, retriever_score_scaling=True,
                    update_retriever=update_retriever, cls_id=CLS, sep_id=SEP, pad_id=PAD)
    model = EMDR2Model(cfg, retriever, settings).to(DEV)
    w32 = {}
    with torch.no_grad():
        for name, p in model.named_parameters():
            w = seeded_weights(name, tuple(p.shape)).to(dtype)
            p.copy_(w)
            w32[name] = w.float()
    return model, w32


def sub(w32, prefix):
    return {k[len(prefix):]: v for k, v in w32.items() if k.startswith(prefix)}


@pytest.mark.parametrize("dtype", [torch.float16])
def test_retrieve_and_read_forward_tiny(tmp_path, dtype):
    from emdr2_b200 import formatter, losses
    from emdr2_b200.indexer import IndexBuilder
    from emdr2_b200.retriever import B200EvidenceRetriever
    from emdr2_b200.store import EvidenceStore
    from emdr2_b200.titlemap import TitleDocMap
    from oracle import blocks as ob, mips as om
    rng = np.random.RandomState(11)
    passages, titles, pairs = corpus(rng)
    n = len(passages)
    titlemap = TitleDocMap(pairs=pairs)

    # ---- index build (a11): context tower over the whole evidence set -> pickle store
    model, w32 = build_model(dtype, None, True)
    def batches():
        for lo in range(0, n, 64):
            ids, types = [], []
            for j in range(lo, min(n, lo + 64)):
                i, t, _ = formatter.context_bert_format(titles[j].tolist() + [SEP] + passages[j].tolist(),
                                                        S_RET, CLS, SEP, PAD)
                ids.append(i)
                types.append(t)
            yield torch.arange(lo + 1, min(n, lo + 64) + 1), torch.tensor(ids), torch.tensor(types)
    path = str(tmp_path / "evidence.pkl")
    IndexBuilder(model.retriever_model, batches(), embedding_path=path).build_and_save_index(expected_total=n)
    store = EvidenceStore(path)
    assert len(store.embed_data) == n and next(iter(store.embed_data.values())).dtype == np.float16
    ids_arr, rows_arr = store.to_arrays()
    # the stored embeddings equal the oracle's context tower within 16-bit tolerance
    b0 = next(batches())
    want_emb = ob.bert_pooled(b0[1], b0[2], sub(w32, "retriever_model.context_model."), TINY["heads"], TINY["layers"])
    assert np.allclose(rows_arr[:64].astype(np.float32), want_emb.numpy(), rtol=2e-2, atol=2e-2)

    retriever = B200EvidenceRetriever(TOPK, TINY["hidden"], embedding_path=path, allow_trivial_doc=False,
                                      passages_map=passages, title_map=titles, wikititledocmap=titlemap)
    model.evidence_retriever = retriever
    assert retriever.topk == TOPK + 1

    # ---- a batch of questions; question 1 "originates" from passage 17 (its uid equals that id)
    bsz = 3
    q_bert = torch.zeros(bsz, S_RET, dtype=torch.int64)
    q_t5 = torch.zeros(bsz, 16, dtype=torch.int64)
    q_len = []
    for i in range(bsz):
        ln = int(rng.randint(4, 10))
        toks = rng.randint(5, TINY["vocab"], size=ln)
        q_bert[i, 0], q_bert[i, 1:1 + ln], q_bert[i, 1 + ln] = CLS, torch.from_numpy(toks), SEP
        q_t5[i, :ln] = torch.from_numpy(toks)
        q_len.append(ln)
    q_types = torch.zeros_like(q_bert)
    dec = torch.zeros(bsz, L, dtype=torch.int64)
    labels = torch.zeros(bsz, L, dtype=torch.int64)
    for i in range(bsz):
        ln = int(rng.randint(2, L))
        dec[i, :ln] = torch.from_numpy(rng.randint(5, TINY["vocab"], size=ln))
        labels[i, :ln] = torch.from_numpy(rng.randint(5, TINY["vocab"], size=ln))
    loss_mask = (labels > 0).float()

    with torch.no_grad():
        q_emb = model.retriever_embedder(q_bert.to(DEV), None, q_types.to(DEV), "query")
    got_topk, _ = retriever.get_topk(q_emb)
    uid = torch.tensor([-1, int(got_topk[1][0][1]), -3])            # query 1 came from its rank-1 passage

    # retrieval parity: oracle MIPS over the SAME stored fp16 rows and the SAME query embeddings
    want_s, want_i, ties = om.mips_topk(rows_arr, q_emb.detach().cpu().numpy(), TOPK + 1, ids=ids_arr, want_ties=True)
    got_ids = np.array([t[0] for t in got_topk])
    assert np.array_equal(got_ids[ties == 0], want_i[ties == 0])

    torch.set_grad_enabled(False)
    # the rectangular kernels first (token-packed execution off): exact equalities between the model's own paths
    all_towers = (model.language_model.language_model, model.retriever_model.context_model.language_model,
                  model.retriever_model.query_model.language_model)
    for lm in all_towers:
        lm.packed_varlen = False
    model.train()
    lm_logits, topk_log_probs, one_ctx = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV),
                                               torch.tensor(q_len).to(DEV), dec.to(DEV))
    model.eval()
    ev = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
               dec.to(DEV))
    assert torch.equal(ev[0], lm_logits)
    assert ev[2].shape[0] == bsz and ev[2].shape[1] % TOPK == 0 and ev[2].shape[1] <= TOPK * S   # trimmed FiD axis
    assert ev[3].shape == ev[2].shape[:2]
    again = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                  dec.to(DEV), all_query_context_hidden_states=ev[2], all_query_context_ids_unflat=ev[3],
                  topk_log_probs=ev[1])
    assert torch.equal(again[0], lm_logits)                            # cached-encoder re-entry (:96,213)

    # ---- oracle composition with the retrieved ids
    ctx_ids, ctx_types, ext, one = formatter.postprocess_arrays(uid.tolist(), q_t5.tolist(), q_len, got_topk, TOPK,
                                                                S_RET, S, CLS, SEP, PAD)
    wq, wc = sub(w32, "retriever_model.query_model."), sub(w32, "retriever_model.context_model.")
    wt5 = sub(w32, "language_model.")
    oq = ob.bert_pooled(q_bert, q_types, wq, TINY["heads"], TINY["layers"])
    oc = ob.bert_pooled(torch.from_numpy(ctx_ids).view(-1, S_RET), torch.from_numpy(ctx_types).view(-1, S_RET), wc,
                        TINY["heads"], TINY["layers"]).view(bsz, TOPK, -1)
    want_lp = torch.log_softmax(torch.bmm(oq[:, None], oc.transpose(1, 2)) / math.sqrt(TINY["hidden"]), dim=2)[:, 0]
    assert torch.allclose(topk_log_probs.cpu(), want_lp, rtol=2e-2, atol=2e-2)
    ext_t, one_t = torch.from_numpy(ext), torch.from_numpy(one)
    enc = ob.t5_encode(ext_t, wt5, TINY["heads"], TINY["layers"])
    want_logits = ob.t5_decode(dec, enc.reshape(bsz, TOPK * S, -1), ext_t.reshape(bsz, TOPK * S), wt5,
                               TINY["heads"], TINY["layers"])
    live = dec > 0      # padding decoder positions are masked by the loss and not compared (skip_padding)
    assert torch.allclose(lm_logits.float().cpu()[live], want_logits[live], rtol=2e-2, atol=2e-2)
    want_one, _ = ob.t5_forward(one_t, torch.repeat_interleave(dec, TOPK, dim=0), wt5, TINY["heads"], TINY["layers"])
    assert one_ctx.shape == (bsz, TOPK, L, TINY["vocab"])
    live_rep = torch.repeat_interleave(live, TOPK, dim=0)
    assert torch.allclose(one_ctx.float().cpu().view(-1, L, TINY["vocab"])[live_rep], want_one[live_rep],
                          rtol=2e-2, atol=2e-2)

    # trimming padding columns never changes a non-padding position: same logits without it
    model.settings["trim_padding"] = False
    full = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                 dec.to(DEV))
    model.settings["trim_padding"] = True
    assert full[2].shape == (bsz, TOPK * S, TINY["hidden"])
    assert torch.equal(full[0][live], lm_logits[live]) and torch.equal(full[1], topk_log_probs)

    # length-bucketed towers (blocks.py: encode; forced on for this small batch) do not either
    towers = (model.language_model.language_model, model.retriever_model.context_model.language_model)
    for lm in towers:
        lm.bucket_min_rows, lm.length_buckets = 4, 3
    bucketed = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                     dec.to(DEV))
    for lm in towers:
        del lm.bucket_min_rows, lm.length_buckets
    assert bucketed[2].shape == ev[2].shape
    assert torch.equal(bucketed[0][live], lm_logits[live]) and torch.equal(bucketed[1], topk_log_probs)

    # ---- token-packed execution (the default of the no-grad eval path, emdr2_b200/packed.py): same logits to the
    # rounding of the 16-bit probabilities / the range merge, encoder states come back packed, cached re-entry works
    for lm in all_towers:
        del lm.packed_varlen
    pk = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV), dec.to(DEV))
    assert hasattr(pk[2], "cross_plan") and pk[2].states.shape[1] == TINY["hidden"]
    tol = 3e-2 if dtype == torch.bfloat16 else 4e-3
    assert (pk[0].float()[live] - lm_logits.float()[live]).abs().max().item() < tol
    assert torch.allclose(pk[1], topk_log_probs, atol=tol, rtol=0)
    pk_again = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                     dec.to(DEV), all_query_context_hidden_states=pk[2], all_query_context_ids_unflat=pk[3],
                     topk_log_probs=pk[1])
    assert torch.equal(pk_again[0], pk[0])
    model.settings["packed_states"] = False                  # padded FiD states on request (the reference's layout)
    padded = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV), dec.to(DEV))
    del model.settings["packed_states"]
    assert padded[2].dim() == 3 and padded[2].shape[0] == bsz
    assert (padded[0].float()[live] - lm_logits.float()[live]).abs().max().item() < tol

    # ---- evaluation decoding (search_strategy.py) through the real model: greedy tokens equal a greedy
    # decode of the oracle reader wherever the oracle's top-2 logit margin exceeds the 16-bit tolerance
    from emdr2_b200 import search_strategy as ss
    bos, eos = 3, TINY["vocab"] - 1
    inputs = (uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV))
    greedy = ss.reader_generate(model, inputs, max_decode_len=4, bos_id=bos, eos_id=eos, beam_size=1,
                                topk_evidence=TOPK)
    beam = ss.reader_generate(model, inputs, max_decode_len=4, bos_id=bos, eos_id=eos, beam_size=3,
                              topk_evidence=TOPK)
    assert len(greedy) == bsz == len(beam) and all(1 <= len(g) <= 4 for g in greedy)
    # the decode loops above ran on the decoder cache (blocks.DecoderCache: each token through the stack
    # once, encoder states projected once per layer); the reference's way — the whole decoder for every
    # token — must produce the same tokens, and the same last-position logits
    model.supports_decoder_cache = False
    assert ss.reader_generate(model, inputs, max_decode_len=4, bos_id=bos, eos_id=eos, beam_size=1,
                              topk_evidence=TOPK) == greedy
    assert ss.reader_generate(model, inputs, max_decode_len=4, bos_id=bos, eos_id=eos, beam_size=3,
                              topk_evidence=TOPK) == beam
    del model.supports_decoder_cache
    from emdr2_b200.blocks import DecoderCache
    with torch.no_grad():
        first = model(*inputs, torch.full((bsz, 1), bos, dtype=torch.int64, device=DEV))
        prefix = torch.tensor([[bos] + (g + [eos] * 4)[:3] for g in greedy], device=DEV)
        full = model(*inputs, prefix, all_query_context_hidden_states=first[2], all_query_context_ids_unflat=first[3],
                     topk_log_probs=first[1])[0]
        cache = DecoderCache(8)
        for t in range(1, 5):
            step = model(*inputs, prefix[:, :t], all_query_context_hidden_states=first[2],
                         all_query_context_ids_unflat=first[3], topk_log_probs=first[1], decoder_cache=cache)[0]
            assert step.shape[1] == 1 and torch.allclose(step[:, -1].float(), full[:, t - 1].float(), atol=2e-2, rtol=0)
        assert cache.t == 4 and len(cache.cross_kv) == TINY["layers"]
    enc_flat, ids_flat = enc.reshape(bsz, TOPK * S, -1), ext_t.reshape(bsz, TOPK * S)
    y = torch.full((bsz, 1), bos, dtype=torch.int64)
    for step in range(4):
        ologits = ob.t5_decode(y, enc_flat, ids_flat, wt5, TINY["heads"], TINY["layers"])[:, -1, :]
        top2 = ologits.topk(2, dim=1)
        for i in range(bsz):
            done = eos in greedy[i][:step] or step >= len(greedy[i])
            if not done and float(top2.values[i, 0] - top2.values[i, 1]) > 5e-2:
                assert greedy[i][step] == int(top2.indices[i, 0]), (i, step)
        nxt = torch.tensor([[g[step] if step < len(g) else eos] for g in greedy])
        y = torch.cat([y, nxt], dim=1)

    # ---- losses (a10) on the GPU logits vs fp32 torch on the oracle logits
    lm_loss = losses.reader_cross_entropy(lm_logits, labels.to(DEV), loss_mask.to(DEV))
    want_lm = (torch.nn.functional.cross_entropy(want_logits.view(-1, TINY["vocab"]), labels.view(-1),
                                                 reduction="none", ignore_index=0) * loss_mask.view(-1)).sum() \
        / loss_mask.sum()
    assert abs(float(lm_loss) - float(want_lm)) < 2e-2
    r_loss, util, null_loss = losses.get_loss_and_retriever_utility(one_ctx, topk_log_probs, labels.to(DEV),
                                                                    loss_mask.to(DEV), eos_id=TINY["vocab"] - 2)
    lp = torch.log_softmax(want_one.view(bsz, TOPK, L, -1), dim=-1)
    gold = torch.gather(lp, -1, labels[:, None, :, None].expand(-1, TOPK, -1, 1)).squeeze(-1)
    want_r, want_u, want_n = losses.loss_and_retriever_utility_from_gold(gold, want_lp, labels, loss_mask,
                                                                         TINY["vocab"] - 2)
    assert abs(float(r_loss) - float(want_r)) < 3e-2 and abs(float(null_loss) - float(want_n)) < 3e-2
    assert abs(float(util) - float(want_u)) < 3e-2
    torch.set_grad_enabled(True)

    # ---- one full training step through autograd: every parameter that the loss depends on gets a
    # finite gradient and the reader loss decreases after a plain SGD step
    model.train()
    out = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                dec.to(DEV))
    l0 = losses.reader_cross_entropy(out[0], labels.to(DEV), loss_mask.to(DEV))
    r0, _, _ = losses.get_loss_and_retriever_utility(out[2], out[1], labels.to(DEV), loss_mask.to(DEV),
                                                     eos_id=TINY["vocab"] - 2)
    (l0 + r0).backward()
    with_grad = [n for n, p in model.named_parameters() if p.grad is not None]
    assert any("retriever_model.query_model" in n for n in with_grad)
    assert any("retriever_model.context_model" in n for n in with_grad)
    assert any("language_model.language_model.decoder" in n for n in with_grad)
    assert all(torch.isfinite(p.grad.float()).all() for p in model.parameters() if p.grad is not None)
    with torch.no_grad():
        for p in model.parameters():
            if p.grad is not None:
                p.add_(p.grad.float().clamp(-1, 1).to(p.dtype), alpha=-0.02)
        out2 = model(uid.to(DEV), q_bert.to(DEV), q_types.to(DEV), None, q_t5.to(DEV), torch.tensor(q_len).to(DEV),
                     dec.to(DEV))
        l1 = losses.reader_cross_entropy(out2[0], labels.to(DEV), loss_mask.to(DEV))
    assert float(l1) < float(l0)
