"""Answer generation for evaluation: greedy / sampling and beam search over the reader.

Mirrors reference megatron/model/search_strategy.py:124-240 (`BeamSearch`, `SampleOrGreedySearch`) as
used by `reader_em_score` (tasks/openqa/e2eqa/train_e2eqa.py:217-262): same constructor arguments,
same `generate_output(model, query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5,
query_ids_t5_len)` and the same outputs (a list of token-id lists, EOS and everything after it cut).
`model` is an `EMDR2Model`-like callable: the first call retrieves and encodes, every later call
re-enters with the cached encoder states (`all_query_context_hidden_states`,
`all_query_context_ids_unflat`, `topk_log_probs`; emdr2_model.py:96,213-214).

Differences behind that surface: tensors live on the device of the inputs (the reference hard-codes
`.cuda()`), beam bookkeeping uses `index_select` on whole tensors instead of splitting them into
per-row lists (search_strategy.py:91-103), and scores are computed in fp32 from the 16-bit logits.
"""
import torch
import torch.nn.functional as F


def length_penalty(length, alpha):
    """PolynomialNormalization.lp (search_strategy.py:28-29)."""
    return pow(5 + length, alpha) / pow(5 + 1, alpha)


def update_beam_state(outs, total_score, topk, topk_score, eos_id, alpha, tokens_enc, z_block, types):
    """One beam-search step (search_strategy.py:44-105).

    outs [B*k, t] hypotheses so far; total_score [B*k] or None on the first step; topk / topk_score
    [rows, k] the k best next tokens of each scored row (rows = B on the first step, B*k afterwards);
    tokens_enc / z_block / types: per-row model state that follows its hypothesis.  Finished
    hypotheses (EOS seen) keep their score and continue only as one EOS-extended candidate;
    running ones are re-normalised with the polynomial length penalty."""
    full = outs.shape[0]
    prev_full, k = topk.shape
    batch = full // k
    prev_k = prev_full // batch
    assert prev_k in (1, k)
    dev = outs.device

    if total_score is None:
        total_score = topk_score
    else:
        is_end = (outs == eos_id).any(dim=1).view(-1, 1).expand_as(topk_score)
        bias = torch.zeros_like(topk_score)
        bias[:, 1:] = -10000.0            # an ended hypothesis survives once, not k times
        new_len = outs.shape[1]
        normalized = (total_score[:, None] * length_penalty(new_len - 1, alpha) + topk_score) \
            / length_penalty(new_len, alpha)
        total_score = torch.where(is_end, total_score[:, None] + bias, normalized)
        assert float(total_score.max()) < 0.0
        topk = torch.where(is_end, torch.full_like(topk, eos_id), topk)

    total_score = total_score.reshape(batch, prev_k * k)
    best_score, best = torch.topk(total_score, k)                        # [B, k] over prev_k*k candidates
    base = torch.arange(batch, device=dev)[:, None]
    next_token = topk.reshape(-1)[(best + base * prev_k * k).reshape(-1)]
    source = (best // k + base * prev_k).reshape(-1)                     # row each survivor extends
    outs = torch.cat([outs.index_select(0, source), next_token[:, None]], dim=1)
    return (outs, best_score.reshape(-1), tokens_enc.index_select(0, source),
            z_block.index_select(0, source), types.index_select(0, source))


def _first_max(scores):
    """Index of the first maximum per row (the reference scans with a strict `<`, :114)."""
    best = scores.max(dim=1, keepdim=True).values
    k = scores.shape[1]
    idx = torch.arange(k, device=scores.device)[None].expand_as(scores)
    return torch.where(scores == best, idx, torch.full_like(idx, k)).min(dim=1).values


def finish_beam(outs, total_score, batchsize, eos_id):
    """Best hypothesis per question, cut at EOS (search_strategy.py:108-121; first maximum wins).
    Returns (id_list, score_list)."""
    k = outs.shape[0] // batchsize
    rows = outs.tolist()
    scores = total_score.reshape(batchsize, k)
    winners = _first_max(scores).tolist()
    id_list, score_list = [], []
    for i, j in enumerate(winners):
        out = rows[i * k + j]
        if eos_id in out:
            out = out[:out.index(eos_id)]
        id_list.append(out)
        score_list.append(float(scores[i, j]))
    return id_list, score_list


class _ReaderState(object):
    """What `EMDR2Model.forward` returns beside the logits and takes back on the next call: the encoder
    states of the retrieved passages, their ids (for the cross-attention mask) and the retriever's
    log-probabilities.  None before the first call, which retrieves and encodes.

    With a model that advertises `supports_decoder_cache` (emdr2_b200.model.EMDR2Model) the state also owns
    a `blocks.DecoderCache`: every token then passes through the decoder once and the K*S encoder positions
    are projected to keys/values once per layer, instead of once per generated token; the encoder states
    stay one copy per QUESTION (beam hypotheses are rows of the cache, not copies of the states)."""

    def __init__(self, max_len=None):
        self.hidden = self.ids_unflat = self.topk_log_probs = None
        self.max_len = max_len
        self.cache = None

    def step(self, model, question, decoder_ids):
        """Next-token logits [rows, vocab] (fp32) for the hypotheses in decoder_ids [rows, t]."""
        extra = {}
        if self.max_len is not None and getattr(model, "supports_decoder_cache", False):
            if self.cache is None:
                from .blocks import DecoderCache
                self.cache = DecoderCache(self.max_len)
            extra["decoder_cache"] = self.cache
        logits, self.topk_log_probs, self.hidden, self.ids_unflat = model(
            *question, decoder_ids, all_query_context_hidden_states=self.hidden,
            all_query_context_ids_unflat=self.ids_unflat, topk_log_probs=self.topk_log_probs, **extra)
        return logits[:, -1, :].float()


class BeamSearch(object):
    def __init__(self, max_decode_len, bos_id, eos_id, beam_size=5, alpha=0.6, topk_evidence=-1):
        self.max_decode_length = max_decode_len
        self.bos_id = bos_id
        self.eos_id = eos_id
        self.k = beam_size
        self.alpha = alpha
        assert topk_evidence >= 1, "this code is customized for retrieval tasks"

    def generate_output(self, model, query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5,
                        query_ids_t5_len):
        question = (query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5, query_ids_t5_len)
        batch, dev = query_ids_bert.shape[0], query_ids_bert.device
        state = _ReaderState(self.max_decode_length + 1)
        y_block = torch.full((batch, 1), self.bos_id, dtype=torch.int64, device=dev)
        outs = torch.full((batch * self.k, 1), self.bos_id, dtype=torch.int64, device=dev)
        total_score = None
        for _ in range(self.max_decode_length):
            topk_score, topk = torch.topk(F.log_softmax(state.step(model, question, y_block), dim=1), self.k)
            assert float(topk_score.max()) <= 0.0
            if state.cache is not None:
                # cached decoding: the per-question state is shared by a question's hypotheses; only the
                # cache rows follow the survivors (`source` comes back through a row-index stand-in)
                rows = torch.arange(topk.shape[0], device=dev)
                outs, total_score, source, _, _ = update_beam_state(
                    outs, total_score, topk, topk_score, self.eos_id, self.alpha, rows, rows, rows)
                state.cache.reorder(source)
            else:
                outs, total_score, state.ids_unflat, state.hidden, state.topk_log_probs = update_beam_state(
                    outs, total_score, topk, topk_score, self.eos_id, self.alpha, state.ids_unflat, state.hidden,
                    state.topk_log_probs)
            y_block = outs
            if bool((outs == self.eos_id).any(dim=1).all()):
                break                    # every hypothesis has produced EOS
        id_list, _ = finish_beam(outs[:, 1:], total_score, batch, self.eos_id)
        return id_list


class SampleOrGreedySearch(object):
    def __init__(self, max_decode_len, bos_id, eos_id, sample=False, topk_evidence=-1):
        self.max_decode_length = max_decode_len
        self.bos_id = bos_id
        self.eos_id = eos_id
        self.sample = sample
        assert topk_evidence >= 1, "this code is customized for retrieval tasks"

    def generate_output(self, model, query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5,
                        query_ids_t5_len):
        question = (query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5, query_ids_t5_len)
        batch, dev = query_ids_bert.shape[0], query_ids_bert.device
        state = _ReaderState(self.max_decode_length + 1)
        y_block = torch.full((batch, 1), self.bos_id, dtype=torch.int64, device=dev)
        eos_seen = torch.zeros(batch, dtype=torch.bool, device=dev)
        for _ in range(self.max_decode_length):
            last = state.step(model, question, y_block)
            if self.sample:
                ys = torch.multinomial(F.softmax(last, dim=1), num_samples=1).reshape(-1)
            else:
                ys = torch.argmax(last, dim=1)       # argmax of log_softmax == argmax of the logits
            y_block = torch.cat([y_block, ys[:, None]], dim=1)
            eos_seen |= ys == self.eos_id
            if bool(eos_seen.all()):                 # one small device->host read per step, as in the reference
                break
        outs = []
        for y in y_block[:, 1:].tolist():            # drop BOS, cut at the first EOS (search_strategy.py:230-238)
            if self.eos_id in y:
                y = y[:y.index(self.eos_id)]
            if len(y) == 0:
                y = [1]
            outs.append(y)
        return outs


def reader_generate(model, batch_inputs, max_decode_len, bos_id, eos_id, beam_size=1, topk_evidence=1):
    """The strategy choice of reader_em_score (train_e2eqa.py:233-248): greedy for beam_size 1, beam
    search above; `batch_inputs` = (query_uid, query_ids_bert, query_types, query_mask_bert,
    query_ids_t5, query_ids_t5_len)."""
    if beam_size == 1:
        obj = SampleOrGreedySearch(max_decode_len, bos_id, eos_id, sample=False, topk_evidence=topk_evidence)
    elif beam_size > 1:
        obj = BeamSearch(max_decode_len, bos_id, eos_id, beam_size=beam_size, topk_evidence=topk_evidence)
    else:
        raise AssertionError("--beam-size < 1 is not supported for ORQA reader.")
    with torch.no_grad():
        return obj.generate_output(model, *batch_inputs)


__all__ = ["BeamSearch", "SampleOrGreedySearch", "update_beam_state", "finish_beam", "reader_generate",
           "length_penalty"]
