"""BERT tower and T5 reader modules on the sm_100a block kernels (forward and backward).

Host-side mirror of the reference's model surface (DevSinghSachan/emdr2 @ edb8cf67):

  ParallelMLP / ParallelAttention / ParallelTransformerLayer / ParallelTransformer
                                  megatron/model/transformer.py:58,111,420,566
  Embedding / TransformerLanguageModel
                                  megatron/model/language_model.py:98,246
  BertTower   == PretrainedBertModel   megatron/model/dualencoder_model.py:146-194
  T5Reader    == T5Model               megatron/model/t5_model.py:84-154

Same attribute names, parameter shapes and forward signatures, so `named_parameters()` /
`state_dict()` keys equal the reference's and its checkpoints load key-for-key
(`load_reference_state_dict` also accepts the nested dict its state_dict_for_save_checkpoint
writes).  What differs is the execution: every matmul is csrc/gemm.cu (bias, GeLU and the
residual add fused into the store), attention is csrc/attention.cu (scores/probabilities never
reach HBM; the [b,s,s] bool masks the reference materialises are replaced by the padding vectors
they are built from), LayerNorm/embedding are csrc/rowops.cu.  Activations are [tokens, h] with
tokens = batch*seq (the reference's [s,b,h] transposes, transformer.py:662,692, are not needed).

QKV layout: the reference packs the fused projection as [np, hn, 3] along the output dimension
(transformer.py:232-240).  Parameters keep that layout (checkpoint compatibility); the kernels want
[3, np, hn] so that a head's q/k/v are 64 contiguous columns, and `_packed()` keeps a permuted copy
keyed on the parameter version.

Dropout is off (p = 0 / eval): hidden_dropout and attention_dropout are identity here.
Training: every op dispatches through emdr2_b200/autograd.py, whose backward passes are kernels of
the same library (csrc/gemm.cu MN-major/split-K products, attention_bwd.cu, rowops_bwd.cu); under
torch.no_grad() the plain forward ops run.  No CPU path: forward on a CPU tensor raises.
"""
import math

import torch
import torch.nn as nn

from . import autograd as ag
from . import ops

PAD_ID = 0   # tokenizer.pad; masks are `ids >= 1` (megatron/data/mask_creation_utils.py:17-26)


class Linear(nn.Module):
    def __init__(self, in_features, out_features, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(out_features, dtype=dtype))

    def forward(self, x, residual=None):
        return ag.linear(x, self.weight, self.bias, residual=residual)


class LayerNorm(nn.Module):
    def __init__(self, hidden, eps, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(hidden, dtype=dtype))
        self.eps = eps

    def forward(self, x):
        return ag.layernorm(x, self.weight, self.bias, self.eps)


def _unpack_rows(t, heads, hn, splits):
    """[np*hn*splits (np,hn,splits order), ...] -> [splits*np*hn (splits,np,hn order), ...]."""
    rest = t.shape[1:]
    return t.view(heads, hn, splits, *rest).permute(2, 0, 1, *range(3, 3 + len(rest))) \
        .reshape(splits * heads * hn, *rest).contiguous()


class _PackedProjection(nn.Module):
    """A fused projection stored in the reference's [np, hn, splits] row order."""

    def __init__(self, hidden, heads, splits, dtype):
        super().__init__()
        self.heads, self.hn, self.splits = heads, hidden // heads, splits
        self.weight = nn.Parameter(torch.empty(splits * hidden, hidden, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(splits * hidden, dtype=dtype))
        self._cache = None

    def _packed(self):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), self.bias.data_ptr())
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                self._cache = (key, _unpack_rows(self.weight.detach(), self.heads, self.hn, self.splits),
                               _unpack_rows(self.bias.detach(), self.heads, self.hn, self.splits))
        return self._cache[1], self._cache[2]

    def forward(self, x):
        if torch.is_grad_enabled() and (self.weight.requires_grad or x.requires_grad):
            # training: the permutation is part of the graph so gradients land in the parameter's
            # (reference) row order
            w = _unpack_rows(self.weight, self.heads, self.hn, self.splits)
            b = _unpack_rows(self.bias, self.heads, self.hn, self.splits)
        else:
            w, b = self._packed()
        return ag.linear(x, w, b)


class ParallelAttention(nn.Module):
    def __init__(self, hidden, heads, dtype, attention_type="self"):
        super().__init__()
        if hidden // heads != 64 or hidden % heads:
            raise ValueError("the sm_100a attention kernel supports head dimension 64 only "
                             "(hidden=%d heads=%d)" % (hidden, heads))
        self.hidden, self.heads, self.attention_type = hidden, heads, attention_type
        if attention_type == "self":
            self.query_key_value = _PackedProjection(hidden, heads, 3, dtype)
        else:
            self.query = Linear(hidden, hidden, dtype)
            self.key_value = _PackedProjection(hidden, heads, 2, dtype)
        self.dense = Linear(hidden, hidden, dtype)
        self.scale = 1.0 / math.sqrt(hidden // heads)

    def forward(self, x, batch, sq, q_pad, residual, causal=False, encoder_output=None, sk=None,
                k_pad=None, q_live=None, k_live=None):
        if self.attention_type == "self":
            qkv = self.query_key_value(x)
            ctx = ag.self_attention(qkv, batch, self.heads, sq, pad=q_pad, live=q_live, causal=causal,
                                    scale=self.scale)
        else:
            q = self.query(x)
            kv = self.key_value(encoder_output)
            ctx = ag.cross_attention(q, kv, batch, self.heads, sq, sk, q_pad=q_pad, k_pad=k_pad,
                                     q_live=q_live, k_live=k_live, scale=self.scale)
        return self.dense(ctx, residual=residual)          # bias + residual fused (dropout p=0)


class ParallelMLP(nn.Module):
    def __init__(self, hidden, ffn, dtype):
        super().__init__()
        self.dense_h_to_4h = Linear(hidden, ffn, dtype)
        self.dense_4h_to_h = Linear(ffn, hidden, dtype)

    def forward(self, x, residual):
        return ag.mlp(x, self.dense_h_to_4h.weight, self.dense_h_to_4h.bias, self.dense_4h_to_h.weight,
                      self.dense_4h_to_h.bias, residual=residual)


class ParallelTransformerLayer(nn.Module):
    def __init__(self, hidden, heads, ffn, eps, dtype, layer_type="encoder"):
        super().__init__()
        self.layer_type = layer_type
        self.input_layernorm = LayerNorm(hidden, eps, dtype)
        self.self_attention = ParallelAttention(hidden, heads, dtype)
        self.post_attention_layernorm = LayerNorm(hidden, eps, dtype)
        if layer_type == "decoder":
            self.inter_attention = ParallelAttention(hidden, heads, dtype, attention_type="cross")
            self.post_inter_attention_layernorm = LayerNorm(hidden, eps, dtype)
        self.mlp = ParallelMLP(hidden, ffn, dtype)

    def forward(self, x, batch, seq, pad, causal=False, encoder_output=None, enc_seq=None, enc_pad=None,
                q_live=None, enc_live=None):
        x = self.self_attention(self.input_layernorm(x), batch, seq, pad, residual=x, causal=causal,
                                q_live=q_live)
        ln = self.post_attention_layernorm(x)
        if self.layer_type == "decoder":
            x = self.inter_attention(ln, batch, seq, pad, residual=x, encoder_output=encoder_output,
                                     sk=enc_seq, k_pad=enc_pad, q_live=q_live, k_live=enc_live)
            ln = self.post_inter_attention_layernorm(x)
        return self.mlp(ln, residual=x)


class ParallelTransformer(nn.Module):
    def __init__(self, hidden, heads, ffn, num_layers, eps, dtype, layer_type="encoder"):
        super().__init__()
        self.layers = nn.ModuleList([ParallelTransformerLayer(hidden, heads, ffn, eps, dtype, layer_type)
                                     for _ in range(num_layers)])
        self.final_layernorm = LayerNorm(hidden, eps, dtype)

    def forward(self, x, batch, seq, pad, **kw):
        for layer in self.layers:
            x = layer(x, batch, seq, pad, **kw)
        return self.final_layernorm(x)


class _Table(nn.Module):
    def __init__(self, rows, hidden, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(rows, hidden, dtype=dtype))


class Embedding(nn.Module):
    def __init__(self, hidden, vocab, max_positions, num_tokentypes, dtype):
        super().__init__()
        self.word_embeddings = _Table(vocab, hidden, dtype)
        self.position_embeddings = _Table(max_positions, hidden, dtype)
        self.tokentype_embeddings = _Table(num_tokentypes, hidden, dtype) if num_tokentypes > 0 else None

    def forward(self, input_ids, tokentype_ids=None):
        typ = self.tokentype_embeddings.weight if tokentype_ids is not None else None
        return ag.embedding(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                            tokentype_ids, typ)


class TransformerLanguageModel(nn.Module):
    def __init__(self, cfg, num_tokentypes=0, add_decoder=False, vocab_size=None):
        super().__init__()
        d = cfg["dtype"]
        self.hidden = cfg["hidden"]
        self.embedding = Embedding(cfg["hidden"], vocab_size or cfg["vocab"], cfg["max_pos"],
                                   num_tokentypes, d)
        self.encoder = ParallelTransformer(cfg["hidden"], cfg["heads"], cfg["ffn"], cfg["layers"],
                                           cfg.get("eps", 1e-5), d)
        self.add_decoder = add_decoder
        if add_decoder:
            self.decoder = ParallelTransformer(cfg["hidden"], cfg["heads"], cfg["ffn"], cfg["layers"],
                                               cfg.get("eps", 1e-5), d, layer_type="decoder")

    #: True: attention skips padding (zeros at padding positions, identical results at every
    #: non-padding position); False: the reference's values at padding positions too.
    skip_padding = True

    @staticmethod
    def trimmed_width(s, max_len):
        """Columns to keep when the longest sequence of the batch has max_len tokens (multiple of 64)."""
        if max_len is None:
            return s
        return min(s, max(64, -(-int(max_len) // 64) * 64))

    def encode(self, ids, tokentype_ids=None, max_len=None):
        """Encoder states [b, s', h]; with max_len (the longest non-padding length, known to the
        caller on the host) only the first s' = roundup(max_len, 64) columns are computed — every
        dropped column is padding in every sequence, so the kept positions are unchanged."""
        s_keep = self.trimmed_width(ids.shape[1], max_len)
        if s_keep != ids.shape[1]:
            ids = ids[:, :s_keep].contiguous()
            if tokentype_ids is not None:
                tokentype_ids = tokentype_ids[:, :s_keep].contiguous()
        b, s = ids.shape
        pad = ids < 1
        x = self.embedding(ids, tokentype_ids)
        q_live = ops.live_blocks(pad) if self.skip_padding else None
        return self.encoder(x, b, s, pad, q_live=q_live).view(b, s, self.hidden)

    def decode(self, dec_ids, enc_states, enc_pad):
        """enc_states [b, sk, h] (sk may be K*S: FiD concatenation, emdr2_model.py:159-164)."""
        b, sq = dec_ids.shape
        sk = enc_states.shape[1]
        x = self.embedding(dec_ids)
        enc2d = enc_states.reshape(b * sk, self.hidden)
        dec_pad = dec_ids < 1
        q_live = ops.live_blocks(dec_pad) if self.skip_padding else None
        enc_live = ops.live_blocks(enc_pad) if self.skip_padding else None
        y = self.decoder(x, b, sq, dec_pad, causal=True, encoder_output=enc2d, enc_seq=sk,
                         enc_pad=enc_pad, q_live=q_live, enc_live=enc_live)
        return y.view(b, sq, self.hidden)


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("emdr2_b200 has no CPU path: inputs must be CUDA tensors")


class BertTower(nn.Module):
    """PretrainedBertModel: forward(input_ids, attention_mask, tokentype_ids) -> [b, h] CLS state.
    `attention_mask` (the dense [b,s,s] bool the reference takes) is accepted and ignored: it is
    always make_attention_mask_3d(ids, ids) < 0.5, which the kernel derives from the ids."""

    def __init__(self, cfg, num_tokentypes=2, vocab_size=None):
        super().__init__()
        self.language_model = TransformerLanguageModel(cfg, num_tokentypes, False, vocab_size)

    def forward(self, input_ids, attention_mask=None, tokentype_ids=None, max_len=None):
        _require_cuda(input_ids)
        return self.language_model.encode(input_ids, tokentype_ids, max_len=max_len)[:, 0, :]

    def hidden_states(self, input_ids, tokentype_ids=None):
        return self.language_model.encode(input_ids, tokentype_ids)


class _LMHead(nn.Module):
    def __init__(self, vocab, dtype):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(vocab, dtype=dtype))


class T5Reader(nn.Module):
    """T5Model.forward(encoder_input_ids, decoder_input_ids, encoder_attn_mask, decoder_attn_mask,
    encoder_decoder_attn_mask, tokentype_ids=None, lm_labels=None, enc_hidden_states=None,
    output_enc_hidden=False).  The three dense masks are accepted and ignored (they are the
    padding / causal masks of the ids); with `enc_hidden_states` the encoder is bypassed and
    `encoder_input_ids` must be the ids those states were computed from, flattened to [b, sk]
    (pass `enc_ids_for_mask` if the reference call site truncates them, emdr2_model.py:176)."""

    def __init__(self, cfg, num_tokentypes=2, vocab_size=None):
        super().__init__()
        self.language_model = TransformerLanguageModel(cfg, num_tokentypes, True, vocab_size)
        self.lm_head = _LMHead(vocab_size or cfg["vocab"], cfg["dtype"])

    def forward(self, encoder_input_ids, decoder_input_ids, encoder_attn_mask=None,
                decoder_attn_mask=None, encoder_decoder_attn_mask=None, tokentype_ids=None,
                lm_labels=None, enc_hidden_states=None, output_enc_hidden=False,
                enc_ids_for_mask=None, enc_max_len=None):
        _require_cuda(encoder_input_ids)
        lm = self.language_model
        if enc_hidden_states is None:
            # enc_max_len (opt-in): encoder states come back as [b, s', h] with s' = roundup(max_len, 64)
            enc = lm.encode(encoder_input_ids, tokentype_ids, max_len=enc_max_len)
            mask_ids = encoder_input_ids[:, :enc.shape[1]]
        else:
            enc = enc_hidden_states.to(lm.embedding.word_embeddings.weight.dtype)
            mask_ids = enc_ids_for_mask if enc_ids_for_mask is not None else encoder_input_ids
        if output_enc_hidden:
            return enc
        dec = lm.decode(decoder_input_ids, enc, mask_ids < 1)
        b, sq, h = dec.shape
        word = lm.embedding.word_embeddings.weight
        logits = ag.linear(dec.reshape(b * sq, h), word, self.lm_head.bias).view(b, sq, word.shape[0])
        if lm_labels is None:
            return logits, enc
        loss = -ag.token_logprob(logits, lm_labels)          # vocab_parallel_cross_entropy (t5_model.py:139-146)
        return loss, enc


# ---------------------------------------------------------------------- checkpoint compatibility
def _flatten(nested, prefix=""):
    flat = {}
    for k, v in nested.items():
        key = prefix + k if not prefix else prefix + "." + k
        if isinstance(v, dict):
            flat.update(_flatten(v, key))
        else:
            flat[key] = v
    return flat


def load_reference_state_dict(module, state_dict, strict=True):
    """Load either a flat reference state_dict (named_parameters keys) or the nested dict written
    by the reference's state_dict_for_save_checkpoint (language_model.py:392-410:
    {'language_model': {'embedding': {'word_embeddings': {'weight': ..}, ..}, 'encoder': {..}}})."""
    flat = _flatten(state_dict) if any(isinstance(v, dict) for v in state_dict.values()) else dict(state_dict)
    own = dict(module.named_parameters())
    missing = [k for k in own if k not in flat]
    unexpected = [k for k in flat if k not in own]
    if strict and (missing or unexpected):
        raise KeyError("state dict mismatch: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
    with torch.no_grad():
        for k, p in own.items():
            if k in flat:
                p.copy_(torch.as_tensor(flat[k]).to(device=p.device, dtype=p.dtype))
    return missing, unexpected


def bert_base_config(dtype=torch.bfloat16):
    """BERT-base / T5-base-shaped stack of the reference recipe (examples/openqa/emdr2_nq.sh:73-84)."""
    return dict(hidden=768, heads=12, layers=12, ffn=3072, vocab=30592, max_pos=512, dtype=dtype)
