"""CPU restatement (plain PyTorch, fp32) of the reference's BERT tower and T5 reader forward.

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's baseline
legs; never from emdr2_b200/.

What it follows (reference @ edb8cf67), function by function:
  embedding            megatron/model/language_model.py:169-181   word + position (+ tokentype), dropout off
  attention            megatron/model/transformer.py:212-394      QKV packed [np, hn, 3] (:232-240) / KV
                                                                  packed [np, hn, 2] (:254-262); scores
                                                                  scaled by 1/sqrt(hn) (:309); masked_fill
                                                                  (mask, -10000) then softmax
                                                                  (fused_softmax.py:116-125; bert_model.py:31,
                                                                  t5_model.py:28); P·V; dense
  mlp                  transformer.py:58-108                       h->4h, exact-erf GeLU (F.gelu, :80), 4h->h
  layer                transformer.py:474-563                      pre-LN; x + attn(LN(x)); [+ cross(LN(.))];
                                                                  + mlp(LN(.)); dropout off
  transformer          transformer.py:648-699                      layers, then final LayerNorm
  bert_pooled          megatron/model/dualencoder_model.py:166-181 hidden state of token 0 (no pooler)
  t5_forward           megatron/model/t5_model.py:112-146 and language_model.py:312-358 (encoder bypass
                       via enc_hidden_states :324-330); logits = dec · W_wordᵀ + lm_head.bias (:76-81)
  masks                megatron/data/mask_creation_utils.py:17-42 (True = masked after `< 0.5`)

Weights are a flat {reference parameter name: tensor} dict (the names of the reference modules'
named_parameters()).  Pinned against outputs of the reference modules themselves run on CPU:
tests/golden/make_blocks_golden.py -> tests/golden/blocks_ref_{bert,t5}.npz, checked by
tests/test_blocks_oracle.py (fp32, 1e-5).  The reference ships no tests of its own for this path.
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-5


def pad_mask_3d(source_ids, target_ids):
    """make_attention_mask_3d(source, target) < 0.5: True where the pair is masked."""
    return ~((target_ids[:, None, :] >= 1) & (source_ids[:, :, None] >= 1))


def history_mask_3d(ids):
    n = ids.shape[1]
    ar = torch.arange(n, device=ids.device)
    return (ar[None, :] <= ar[:, None])[None].expand(ids.shape[0], n, n)


def decoder_self_mask(dec_ids):
    keep = (~pad_mask_3d(dec_ids, dec_ids)) & history_mask_3d(dec_ids)
    return ~keep


def _ln(x, w, prefix):
    return F.layer_norm(x, (x.shape[-1],), w[prefix + ".weight"], w[prefix + ".bias"], EPS)


def _softmax_scores(q, k, v, mask, heads):
    """q [b, sq, h], k/v [b, sk, h] already split per head order (head-major columns)."""
    b, sq, h = q.shape
    sk = k.shape[1]
    hn = h // heads
    q = q.view(b, sq, heads, hn).permute(0, 2, 1, 3)
    k = k.view(b, sk, heads, hn).permute(0, 2, 1, 3)
    v = v.view(b, sk, heads, hn).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hn)
    if mask is not None:
        s = s.masked_fill(mask[:, None], -10000.0)
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v).permute(0, 2, 1, 3).reshape(b, sq, h)


def self_attention(x, mask, w, prefix, heads):
    b, s, h = x.shape
    hn = h // heads
    mixed = F.linear(x, w[prefix + ".query_key_value.weight"], w[prefix + ".query_key_value.bias"])
    mixed = mixed.view(b, s, heads, hn, 3)                 # transformer.py:232-240
    q, k, v = (mixed[..., i].reshape(b, s, h) for i in range(3))
    ctx = _softmax_scores(q, k, v, mask, heads)
    return F.linear(ctx, w[prefix + ".dense.weight"]), w[prefix + ".dense.bias"]


def cross_attention(x, enc, mask, w, prefix, heads):
    b, sq, h = x.shape
    sk = enc.shape[1]
    hn = h // heads
    kv = F.linear(enc, w[prefix + ".key_value.weight"], w[prefix + ".key_value.bias"])
    kv = kv.view(b, sk, heads, hn, 2)                      # transformer.py:254-262
    k, v = (kv[..., i].reshape(b, sk, h) for i in range(2))
    q = F.linear(x, w[prefix + ".query.weight"], w[prefix + ".query.bias"])
    ctx = _softmax_scores(q, k, v, mask, heads)
    return F.linear(ctx, w[prefix + ".dense.weight"]), w[prefix + ".dense.bias"]


def mlp(x, w, prefix):
    inter = F.gelu(F.linear(x, w[prefix + ".dense_h_to_4h.weight"], w[prefix + ".dense_h_to_4h.bias"]))
    return F.linear(inter, w[prefix + ".dense_4h_to_h.weight"]), w[prefix + ".dense_4h_to_h.bias"]


def layer(x, mask, w, prefix, heads, enc=None, cross_mask=None):
    ln = _ln(x, w, prefix + ".input_layernorm")
    out, bias = self_attention(ln, mask, w, prefix + ".self_attention", heads)
    x = x + (out + bias)
    ln = _ln(x, w, prefix + ".post_attention_layernorm")
    if enc is not None:
        out, bias = cross_attention(ln, enc, cross_mask, w, prefix + ".inter_attention", heads)
        x = x + (out + bias)
        ln = _ln(x, w, prefix + ".post_inter_attention_layernorm")
    out, bias = mlp(ln, w, prefix + ".mlp")
    return x + (out + bias)


def transformer(x, mask, w, prefix, heads, layers, enc=None, cross_mask=None):
    for i in range(layers):
        x = layer(x, mask, w, "%s.layers.%d" % (prefix, i), heads, enc, cross_mask)
    return _ln(x, w, prefix + ".final_layernorm")


def embedding(ids, w, prefix, types=None):
    s = ids.shape[1]
    e = w[prefix + ".word_embeddings.weight"][ids] + w[prefix + ".position_embeddings.weight"][:s][None]
    if types is not None:
        e = e + w[prefix + ".tokentype_embeddings.weight"][types]
    return e


def bert_hidden(ids, types, w, heads, layers, prefix="language_model"):
    mask = pad_mask_3d(ids, ids)
    x = embedding(ids, w, prefix + ".embedding", types)
    return transformer(x, mask, w, prefix + ".encoder", heads, layers)


def bert_pooled(ids, types, w, heads, layers):
    return bert_hidden(ids, types, w, heads, layers)[:, 0, :]


def t5_encode(enc_ids, w, heads, layers, prefix="language_model"):
    x = embedding(enc_ids, w, prefix + ".embedding")
    return transformer(x, pad_mask_3d(enc_ids, enc_ids), w, prefix + ".encoder", heads, layers)


def t5_decode(dec_ids, enc_states, enc_ids_for_mask, w, heads, layers, prefix="language_model"):
    """Decoder + tied LM head over given encoder states ([b, sk, h]; sk may be K*S for FiD)."""
    x = embedding(dec_ids, w, prefix + ".embedding")
    y = transformer(x, decoder_self_mask(dec_ids), w, prefix + ".decoder", heads, layers,
                    enc=enc_states, cross_mask=pad_mask_3d(dec_ids, enc_ids_for_mask))
    return F.linear(y, w[prefix + ".embedding.word_embeddings.weight"], w["lm_head.bias"])


def t5_forward(enc_ids, dec_ids, w, heads, layers):
    enc = t5_encode(enc_ids, w, heads, layers)
    return t5_decode(dec_ids, enc, enc_ids, w, heads, layers), enc
