"""EMDR2 retrieve-and-read model (forward orchestration) on the B200 kernels.

Mirrors reference megatron/model/emdr2_model.py:31-214 (`EMDR2Model`) and
megatron/model/dualencoder_model.py:27-82 (`DualEncoderModel`): same attribute names
(`language_model`, `retriever_model.query_model/context_model`, checkpoint keys
'encoder/t5_model' and 'retriever/biencoder_model'), same forward signature and return tuples.
The steps are the reference's:

  1 query tower            -> [B, h]                              emdr2_model.py:98-104
  2 retrieve               -> top-k doc ids + passage tokens      :107-108   (emdr2_b200/retriever.py)
  3 format                 -> BERT / T5 input ids                 :110-115   (emdr2_b200/formatter.py)
  4 context tower          -> [B, K, h]                           :118-131
  5 fresh scores           -> log_softmax(q·c / sqrt(h)) [B, K]   :134-145
  6 T5 encoder, K passages -> [B, K*S, h] (FiD concatenation)     :148-164
  7 T5 decoder + LM head   -> logits [B, L, V]                    :166-183
  8 one-context pass       -> logits [B, K, L, V] (training, --update-retriever)   :185-210

Differences behind that surface: masks are never materialised (the kernels derive them from the
ids, which is how the reference builds them: make_attention_mask_3d(ids, ids) < 0.5), gradients
flow through emdr2_b200/autograd.py (library kernels for every backward op), and configuration is
passed explicitly instead of through the global get_args().
"""
import math

import torch
import torch.nn as nn

from . import formatter
from .blocks import BertTower, T5Reader


class DualEncoder(nn.Module):
    """DualEncoderModel: separate query and context BERT towers, CLS embedding without pooler."""

    def __init__(self, cfg, bert_vocab_size=None, only_query_model=False, only_context_model=False):
        super().__init__()
        assert not (only_query_model and only_context_model)
        self.use_query_model = not only_context_model
        self.use_context_model = not only_query_model
        if self.use_query_model:
            self.query_model = BertTower(cfg, num_tokentypes=2, vocab_size=bert_vocab_size)
        if self.use_context_model:
            self.context_model = BertTower(cfg, num_tokentypes=2, vocab_size=bert_vocab_size)

    @staticmethod
    def embed_text(model, tokens, attention_mask, token_types, max_len=None, row_lengths=None):
        """dualencoder_model.py:76-82 (+ the host-known lengths that let the tower trim / bucket)."""
        return model(tokens, attention_mask, token_types, max_len=max_len, row_lengths=row_lengths)

    def forward(self, query_tokens, query_attention_mask, query_types, context_tokens,
                context_attention_mask, context_types):
        q = self.embed_text(self.query_model, query_tokens, query_attention_mask, query_types) \
            if self.use_query_model else None
        c = self.embed_text(self.context_model, context_tokens, context_attention_mask, context_types) \
            if self.use_context_model else None
        return q, c

    def state_dict_for_save_checkpoint(self, destination=None, prefix='', keep_vars=False):
        """{'query_model': {'language_model': ..}, 'context_model': ..} (dualencoder_model.py:84-98)."""
        out = {}
        if self.use_query_model:
            out["query_model"] = self.query_model.state_dict_for_save_checkpoint(destination, prefix, keep_vars)
        if self.use_context_model:
            out["context_model"] = self.context_model.state_dict_for_save_checkpoint(destination, prefix, keep_vars)
        return out

    def load_state_dict(self, state_dict, strict=True):
        if any(isinstance(v, dict) for v in state_dict.values()):      # the reference's nested layout (:100-109)
            if self.use_query_model:
                self.query_model.load_state_dict(state_dict["query_model"], strict)
            if self.use_context_model:
                self.context_model.load_state_dict(state_dict["context_model"], strict)
            return
        return super().load_state_dict(state_dict, strict)


class EMDR2Model(nn.Module):
    """cfg: dict(hidden, heads, layers, ffn, vocab, max_pos, dtype).  `settings` carries what the
    reference reads from get_args()/tokenizers on every call (emdr2_model.py:92,250-303):
    topk_retrievals, seq_length, seq_length_ret, retriever_score_scaling, update_retriever,
    no_query_embedder_training / no_context_embedder_training (detach that tower's embeddings, :103-104,
    :130-131), disable_retriever_dropout (:69-77), cls_id, sep_id, pad_id; trim_padding (default True) runs the towers only on the columns that hold
    a real token in at least one sequence of the batch (lengths are known on the host from the
    formatter), and length_buckets (default True) additionally lets the towers sort the B*K rows by
    length and run them as a few token-packed buckets (blocks.py: `encode`); both leave every
    non-padding position unchanged."""

    #: the decode loops (search_strategy.py) may pass `decoder_cache=` to forward
    supports_decoder_cache = True

    def __init__(self, cfg, evidence_retriever, settings, t5_vocab_size=None, bert_vocab_size=None):
        super().__init__()
        self.cfg = cfg
        self.settings = dict(settings)
        self.topk = int(settings["topk_retrievals"])
        self.language_model = T5Reader(cfg, num_tokentypes=2, vocab_size=t5_vocab_size)
        self._language_model_key = 'encoder/t5_model'
        self.retriever_model = DualEncoder(cfg, bert_vocab_size=bert_vocab_size)
        self._retriever_model_key = 'retriever/biencoder_model'
        self.evidence_retriever = evidence_retriever

    def retriever_embedder(self, tokens, mask, types, embedder_type, disable_dropout=False, max_len=None,
                           row_lengths=None):
        m = self.retriever_model
        if embedder_type not in ("query", "context"):
            raise ValueError("Invalid embedder type.")
        tower = m.query_model if embedder_type == "query" else m.context_model
        if disable_dropout:          # --disable-retriever-dropout: the reference puts the tower in eval mode (:69-77)
            tower.eval()
        return m.embed_text(tower, tokens, mask, types, max_len=max_len, row_lengths=row_lengths)

    def forward(self, query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5,
                query_ids_t5_len, dec_ids, all_query_context_hidden_states=None,
                all_query_context_ids_unflat=None, topk_log_probs=None, decoder_cache=None):
        st = self.settings
        topk = self.topk
        bsize = query_ids_bert.shape[0]
        hidden = self.cfg["hidden"]
        seq_length = int(st["seq_length"])
        query_one_context_ids = None
        len_one = rows_one = None

        if all_query_context_hidden_states is None:
            no_ret_dropout = bool(st.get("disable_retriever_dropout", False))
            query_logits = self.retriever_embedder(query_ids_bert, query_mask_bert, query_types, "query",
                                                   disable_dropout=no_ret_dropout)
            if st.get("no_query_embedder_training", False):        # emdr2_model.py:103-104
                query_logits = query_logits.detach()
            with torch.no_grad():
                if getattr(self.evidence_retriever, "supports_packed", False):
                    topk_evidence_data, _stale = self.evidence_retriever.get_topk(query_logits.detach(),
                                                                                  as_packed=True)
                elif getattr(self.evidence_retriever, "supports_arrays", False):
                    topk_evidence_data, _stale = self.evidence_retriever.get_topk(query_logits.detach(),
                                                                                  as_arrays=True)
                else:
                    topk_evidence_data, _stale = self.evidence_retriever.get_topk(query_logits.clone().detach())
                (all_context_ids, all_context_types, all_query_extended_context_ids, query_one_context_ids), \
                    lengths = \
                    formatter.postprocess(query_uid, query_ids_t5, query_ids_t5_len, topk_evidence_data,
                                          topk, int(st["seq_length_ret"]), seq_length, st["cls_id"],
                                          st["sep_id"], st["pad_id"], device=query_ids_bert.device,
                                          return_lengths=True)
                len_ctx, len_ext, len_one = lengths
                rows_ctx, rows_ext, rows_one = lengths.rows
                if not st.get("trim_padding", True):
                    len_ctx = len_ext = len_one = rows_ctx = rows_ext = rows_one = None
                elif not st.get("length_buckets", True):
                    rows_ctx = rows_ext = rows_one = None
            s_ret = all_context_ids.shape[-1]
            all_context_logits = self.retriever_embedder(all_context_ids.reshape(-1, s_ret), None,
                                                         all_context_types.reshape(-1, s_ret), "context",
                                                         disable_dropout=no_ret_dropout,
                                                         max_len=len_ctx, row_lengths=rows_ctx)
            if st.get("no_context_embedder_training", False):      # emdr2_model.py:130-131
                all_context_logits = all_context_logits.detach()
            all_context_logits = all_context_logits.reshape(bsize, topk, -1).float()
            topk_sim_scores = torch.bmm(query_logits.unsqueeze(1).float(), all_context_logits.transpose(1, 2))
            if st.get("retriever_score_scaling", True):
                topk_sim_scores = topk_sim_scores / math.sqrt(hidden)
            topk_log_probs = torch.log_softmax(topk_sim_scores, dim=2).squeeze(1)

            # Padding columns shared by all B*K rows are not computed (trim_padding): the FiD key axis is
            # then K * s' instead of the reference's K * seq_length, with the ids trimmed alike.
            enc = self.language_model(all_query_extended_context_ids, dec_ids, output_enc_hidden=True,
                                      enc_max_len=len_ext, enc_row_lengths=rows_ext)
            if hasattr(enc, "cross_plan"):
                # token-packed encoder states (no-grad forward path with host-known lengths): question b's keys are
                # its top-k sequences, contiguous and unpadded; there is no id matrix to mask with
                if not st.get("packed_states", True):
                    enc = enc.to_padded(self.language_model.language_model.trimmed_width(seq_length, len_ext))
            if hasattr(enc, "cross_plan"):
                enc.group = topk
                all_query_context_hidden_states = enc
                all_query_context_ids_unflat = all_query_extended_context_ids.view(bsize, topk, -1)
            else:
                s_enc = enc.shape[1]
                all_query_context_hidden_states = enc.reshape(bsize, topk * s_enc, hidden)
                all_query_context_ids_unflat = all_query_extended_context_ids[:, :s_enc].reshape(bsize, topk * s_enc)

        # decoder_cache (blocks.DecoderCache, evaluation decoding only): logits come back for the positions
        # not decoded yet; the decode loops read [:, -1, :] either way
        extra = {} if decoder_cache is None else {"decoder_cache": decoder_cache}
        if hasattr(all_query_context_hidden_states, "cross_plan"):
            extra["fid_group"] = topk
        lm_logits, _ = self.language_model(all_query_context_ids_unflat[:, :1], dec_ids,
                                           enc_hidden_states=all_query_context_hidden_states,
                                           enc_ids_for_mask=all_query_context_ids_unflat, **extra)
        if self.training:
            lm_logits_one_context = None
            if st.get("update_retriever", False) and query_one_context_ids is not None:
                with torch.no_grad():       # "SG": no gradient through the one-context pass (:186)
                    dec_ids_repeated = torch.repeat_interleave(dec_ids, topk, dim=0)
                    flat, _ = self.language_model(query_one_context_ids, dec_ids_repeated, enc_max_len=len_one,
                                                  enc_row_lengths=rows_one)
                    lm_logits_one_context = flat.reshape(bsize, topk, flat.shape[1], flat.shape[2])
            return lm_logits, topk_log_probs, lm_logits_one_context
        return lm_logits, topk_log_probs, all_query_context_hidden_states, all_query_context_ids_unflat

    def state_dict_for_save_checkpoint(self, destination=None, prefix='', keep_vars=False):
        """The reference's nested layout (emdr2_model.py:217-226): what its T5Model / DualEncoderModel
        load_state_dict index into, so a reference indexer or trainer reads checkpoints saved here."""
        return {self._language_model_key: self.language_model.state_dict_for_save_checkpoint(destination, prefix, keep_vars),
                self._retriever_model_key: self.retriever_model.state_dict_for_save_checkpoint(destination, prefix, keep_vars)}

    def load_state_dict(self, state_dict, strict=True):
        if self._language_model_key in state_dict:
            from .blocks import load_reference_state_dict
            load_reference_state_dict(self.language_model, state_dict[self._language_model_key], strict)
            load_reference_state_dict(self.retriever_model, state_dict[self._retriever_model_key], strict)
            return
        return super().load_state_dict(state_dict, strict)
