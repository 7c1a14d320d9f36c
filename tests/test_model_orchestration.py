"""EMDR2Model.forward orchestration (emdr2_b200/model.py, reference megatron/model/emdr2_model.py:87-214) on
CPU: the towers and the reader are replaced by recording stand-ins, the search by fixed ids, so what is
under test is the plumbing — which ids reach which tower, with which lengths, how the fresh retriever
scores and the FiD axis are formed, what the train / eval modes return, and the cached re-entry."""
import math

import numpy as np
import torch

from emdr2_b200 import formatter
from emdr2_b200.model import EMDR2Model
from emdr2_b200.titlemap import NeighbourTable
from emdr2_b200.tokens import FlatTokenStore
from test_formatter_losses import _FixedSearchRetriever, _flat_corpus

H, V, K, S_RET, S, L = 64, 50, 3, 24, 64, 5
CFG = dict(hidden=64, heads=1, layers=1, ffn=64, vocab=V, max_pos=64, dtype=torch.float32)


class RecordingTower(torch.nn.Module):
    """BertTower stand-in: embedding = mean of the ids (so it depends on every token) in H columns."""

    def __init__(self, log, name):
        super().__init__()
        self.log, self.name = log, name

    def forward(self, tokens, mask, types, max_len=None, row_lengths=None):
        self.log.append((self.name, tokens.clone(), types.clone(), max_len, None if row_lengths is None else np.array(row_lengths)))
        live = (tokens > 0).float()
        mean = (tokens.float() * live).sum(1, keepdim=True) / live.sum(1, keepdim=True).clamp(min=1)
        return mean * torch.linspace(0.1, 1.0, H)[None, :]


class RecordingDual(torch.nn.Module):
    def __init__(self, log):
        super().__init__()
        self.query_model, self.context_model = RecordingTower(log, "query"), RecordingTower(log, "context")

    @staticmethod
    def embed_text(model, tokens, mask, types, max_len=None, row_lengths=None):
        return model(tokens, mask, types, max_len=max_len, row_lengths=row_lengths)


class RecordingReader(torch.nn.Module):
    """T5Reader stand-in with the call forms EMDR2Model uses (blocks.py:T5Reader.forward)."""

    def __init__(self, log):
        super().__init__()
        self.log = log

    def forward(self, encoder_input_ids, decoder_input_ids, enc_hidden_states=None, output_enc_hidden=False,
                enc_ids_for_mask=None, enc_max_len=None, enc_row_lengths=None):
        if enc_hidden_states is None:
            width = encoder_input_ids.shape[1] if enc_max_len is None else min(
                encoder_input_ids.shape[1], max(64, -(-enc_max_len // 64) * 64))
            self.log.append(("encode", encoder_input_ids.clone(), enc_max_len,
                             None if enc_row_lengths is None else np.array(enc_row_lengths)))
            enc = encoder_input_ids[:, :width, None].float().expand(-1, -1, H).contiguous()
            if output_enc_hidden:
                return enc
            mask_ids = encoder_input_ids[:, :width]
        else:
            enc, mask_ids = enc_hidden_states, enc_ids_for_mask
            self.log.append(("decode", enc.shape, mask_ids.clone(), decoder_input_ids.clone()))
        b, t = decoder_input_ids.shape
        pooled = (enc[..., 0] * (mask_ids > 0).float()).sum(1) / (mask_ids > 0).float().sum(1).clamp(min=1)
        logits = pooled[:, None, None] + decoder_input_ids[:, :, None].float() + torch.arange(V).float()[None, None, :] / V
        return logits, enc


def _setup(packed):
    rng = np.random.RandomState(4)
    titlemap, passages, titles = _flat_corpus(rng, n_articles=40)
    n = len(passages)
    bsz = 2
    topk_ids = np.stack([rng.choice(np.arange(1, n + 1), size=K, replace=False) for _ in range(bsz)]).astype(np.int64)
    if packed:
        retriever = _FixedSearchRetriever(topk_ids, FlatTokenStore.from_arrays(passages), FlatTokenStore.from_arrays(titles),
                                          NeighbourTable(titlemap))
    else:
        retriever = _FixedSearchRetriever(topk_ids, passages, titles, titlemap)
    settings = dict(topk_retrievals=K, seq_length=S, seq_length_ret=S_RET, retriever_score_scaling=True,
                    update_retriever=True, cls_id=2, sep_id=3, pad_id=0)
    model = EMDR2Model(CFG, retriever, settings, t5_vocab_size=V, bert_vocab_size=V)
    log = []
    model.retriever_model = RecordingDual(log)
    model.language_model = RecordingReader(log)
    q_bert = torch.zeros(bsz, S_RET, dtype=torch.int64)
    q_t5 = torch.zeros(bsz, 12, dtype=torch.int64)
    q_len = []
    for i in range(bsz):
        ln = int(rng.randint(3, 9))
        toks = torch.from_numpy(rng.randint(5, V, size=ln))
        q_bert[i, 0], q_bert[i, 1:1 + ln], q_bert[i, 1 + ln] = 2, toks, 3
        q_t5[i, :ln] = toks
        q_len.append(ln)
    dec = torch.from_numpy(rng.randint(1, V, size=(bsz, L)))
    inputs = (torch.tensor([-1, -2]), q_bert, torch.zeros_like(q_bert), None, q_t5, torch.tensor(q_len), dec)
    nested, _ = _FixedSearchRetriever(topk_ids, passages, titles, titlemap).get_topk(torch.zeros(bsz, 4), as_arrays=True)
    want = formatter.postprocess_arrays([-1, -2], q_t5.tolist(), q_len, nested, K, S_RET, S, 2, 3, 0)
    return model, log, inputs, want


def _row_lengths(a):
    a2 = a.reshape(-1, a.shape[-1])
    return ((a2 != 0) * np.arange(1, a2.shape[1] + 1)).max(axis=1)


def test_eval_forward_plumbing_and_cached_reentry():
    for packed in (False, True):
        model, log, inputs, (ctx_ids, ctx_types, ext, one) = _setup(packed)
        model.eval()
        with torch.no_grad():
            lm_logits, topk_log_probs, hidden, ids_unflat = model(*inputs)
        names = [e[0] for e in log]
        assert names == ["query", "context", "encode", "decode"]
        # the context tower sees the [CLS] title [SEP] passage rows, with the host-known lengths
        _, c_tokens, c_types, c_max, c_rows = log[1]
        assert torch.equal(c_tokens, torch.from_numpy(ctx_ids).view(-1, S_RET)) and int(c_types.sum()) == 0
        assert c_max == int(_row_lengths(ctx_ids).max()) and np.array_equal(c_rows, _row_lengths(ctx_ids))
        # the reader encodes the extended rows; the FiD axis is K * trimmed width
        _, e_tokens, e_max, e_rows = log[2]
        assert torch.equal(e_tokens, torch.from_numpy(ext)) and np.array_equal(e_rows, _row_lengths(ext))
        s_enc = hidden.shape[1] // K
        assert hidden.shape == (2, K * s_enc, H) and s_enc == 64 and ids_unflat.shape == (2, K * s_enc)
        assert torch.equal(ids_unflat, torch.from_numpy(ext)[:, :s_enc].reshape(2, K * s_enc))
        # fresh retriever scores: log_softmax(q . c / sqrt(hidden)) over the K passages (:134-145)
        q = log[0][1].float()
        qe = RecordingTower([], "x")(log[0][1], None, log[0][2])
        ce = RecordingTower([], "x")(c_tokens, None, c_types).view(2, K, H)
        want_lp = torch.log_softmax(torch.bmm(qe[:, None], ce.transpose(1, 2)) / math.sqrt(CFG["hidden"]), dim=2)[:, 0]
        assert torch.allclose(topk_log_probs, want_lp, atol=1e-6) and q.shape[0] == 2
        assert lm_logits.shape == (2, L, V)
        # cached re-entry (search_strategy's later steps): no tower, no encoder, same logits
        del log[:]
        with torch.no_grad():
            again = model(*inputs, all_query_context_hidden_states=hidden, all_query_context_ids_unflat=ids_unflat,
                          topk_log_probs=topk_log_probs)
        assert [e[0] for e in log] == ["decode"] and torch.equal(again[0], lm_logits) and again[1] is topk_log_probs


def test_training_forward_runs_the_one_context_pass_and_honours_the_switches():
    model, log, inputs, (ctx_ids, ctx_types, ext, one) = _setup(True)
    model.train()
    lm_logits, topk_log_probs, one_ctx = model(*inputs)
    assert [e[0] for e in log] == ["query", "context", "encode", "decode", "encode"]
    _, o_tokens, o_max, o_rows = log[4]
    assert torch.equal(o_tokens, torch.from_numpy(one)) and np.array_equal(o_rows, _row_lengths(one))
    assert one_ctx.shape == (2, K, L, V) and not one_ctx.requires_grad
    # without --update-retriever the one-context pass is skipped (:185)
    model.settings["update_retriever"] = False
    del log[:]
    assert model(*inputs)[2] is None and [e[0] for e in log][-1] == "decode"
    # trim_padding off: full widths, no lengths handed down; length_buckets off: max only
    model.settings.update(update_retriever=True, trim_padding=False)
    del log[:]
    model(*inputs)
    assert log[1][3] is None and log[1][4] is None and log[2][2] is None and log[2][3] is None
    model.settings.update(trim_padding=True, length_buckets=False)
    del log[:]
    model(*inputs)
    assert log[1][3] is not None and log[1][4] is None and log[2][3] is None
