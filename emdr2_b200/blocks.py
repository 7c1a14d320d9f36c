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

Dropout: cfg["hidden_dropout"] / cfg["attention_dropout"] (default 0.1 each, the reference's
--hidden-dropout / --attention-dropout, arguments.py:218-221) act while a module is in training mode, exactly
where the reference applies them — on the attention probabilities (transformer.py:345-346), on every
bias-add-residual (:397-419, :511-515) and on the embedding sum (language_model.py:181) — with counter-based
masks that the backward kernels regenerate (emdr2_b200/dropout.py, csrc/dropout.cuh).  eval() turns them off.
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

    def fork(self, x):
        """(LN(x), x) — the sublayer input and the residual branch; under autograd the two gradients of x are
        summed inside the LayerNorm backward kernel instead of a separate elementwise pass."""
        return ag.layernorm_fork(x, self.weight, self.bias, self.eps)


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
        # the kernel-order copy is cached per parameter version; in training the weight gradient is produced
        # directly in the parameter's (reference) row order (autograd._PackedLinearFn)
        w, b = self._packed()
        return ag.packed_linear(x, self.weight, self.bias, w, b, self.heads, self.hn, self.splits)


class ParallelAttention(nn.Module):
    def __init__(self, hidden, heads, dtype, attention_type="self", attention_dropout=0.0, hidden_dropout=0.0):
        super().__init__()
        self.attention_dropout, self.hidden_dropout = float(attention_dropout), float(hidden_dropout)
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
                k_pad=None, q_live=None, k_live=None, groups=None, packed=None, cross_plan=None, cross_kv=None):
        p_attn = self.attention_dropout if self.training else 0.0
        p_hidden = self.hidden_dropout if self.training else 0.0
        if packed is not None:          # token-packed sequences, one varlen launch (no-grad forward path)
            qkv = self.query_key_value(x)
            h = self.hidden
            ctx = ops.attention_varlen(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], self.heads, packed.items,
                                       packed.n_items, scale=self.scale, flops=packed.attention_flops)
            return self.dense(ctx, residual=residual)
        if cross_plan is not None:      # FiD cross-attention over token-packed encoder states
            q = self.query(x)
            kv = cross_kv if cross_kv is not None else self.key_value(encoder_output)
            ctx = ag.cross_attention_packed(q, kv, self.heads, cross_plan, scale=self.scale)
            return self.dense(ctx, residual=residual)
        if self.attention_type == "self":
            qkv = self.query_key_value(x)
            if groups is not None:      # token-packed rows of several [batch_g, seq_g] rectangles
                ctx = ag.self_attention_grouped(qkv, self.heads, groups, causal=causal, scale=self.scale,
                                                dropout_p=p_attn)
            else:
                ctx = ag.self_attention(qkv, batch, self.heads, sq, pad=q_pad, live=q_live, causal=causal,
                                        scale=self.scale, dropout_p=p_attn)
        else:
            q = self.query(x)
            kv = self.key_value(encoder_output)
            ctx = ag.cross_attention(q, kv, batch, self.heads, sq, sk, q_pad=q_pad, k_pad=k_pad,
                                     q_live=q_live, k_live=k_live, scale=self.scale, dropout_p=p_attn)
        if p_hidden:                                       # residual + dropout(ctx W^T + b)   (transformer.py:397-419)
            return ag.dropout_add(self.dense(ctx), residual, p_hidden)
        return self.dense(ctx, residual=residual)          # bias + residual fused into the GEMM epilogue


class ParallelMLP(nn.Module):
    def __init__(self, hidden, ffn, dtype, hidden_dropout=0.0):
        super().__init__()
        self.dense_h_to_4h = Linear(hidden, ffn, dtype)
        self.dense_4h_to_h = Linear(ffn, hidden, dtype)
        self.hidden_dropout = float(hidden_dropout)

    def forward(self, x, residual):
        w = (self.dense_h_to_4h.weight, self.dense_h_to_4h.bias, self.dense_4h_to_h.weight, self.dense_4h_to_h.bias)
        if self.training and self.hidden_dropout:          # residual + dropout(mlp(x))   (transformer.py:511-515)
            return ag.dropout_add(ag.mlp(x, *w, residual=None), residual, self.hidden_dropout)
        return ag.mlp(x, *w, residual=residual)


class ParallelTransformerLayer(nn.Module):
    def __init__(self, hidden, heads, ffn, eps, dtype, layer_type="encoder", attention_dropout=0.0,
                 hidden_dropout=0.0):
        super().__init__()
        self.layer_type = layer_type
        drop = dict(attention_dropout=attention_dropout, hidden_dropout=hidden_dropout)
        self.input_layernorm = LayerNorm(hidden, eps, dtype)
        self.self_attention = ParallelAttention(hidden, heads, dtype, **drop)
        self.post_attention_layernorm = LayerNorm(hidden, eps, dtype)
        if layer_type == "decoder":
            self.inter_attention = ParallelAttention(hidden, heads, dtype, attention_type="cross", **drop)
            self.post_inter_attention_layernorm = LayerNorm(hidden, eps, dtype)
        self.mlp = ParallelMLP(hidden, ffn, dtype, hidden_dropout=hidden_dropout)

    def forward(self, x, batch, seq, pad, causal=False, encoder_output=None, enc_seq=None, enc_pad=None,
                q_live=None, enc_live=None, groups=None, packed=None, cross_plan=None):
        ln, x = self.input_layernorm.fork(x)
        x = self.self_attention(ln, batch, seq, pad, residual=x, causal=causal,
                                q_live=q_live, groups=groups, packed=packed)
        ln, x = self.post_attention_layernorm.fork(x)
        if self.layer_type == "decoder":
            x = self.inter_attention(ln, batch, seq, pad, residual=x, encoder_output=encoder_output,
                                     sk=enc_seq, k_pad=enc_pad, q_live=q_live, k_live=enc_live,
                                     cross_plan=cross_plan)
            ln, x = self.post_inter_attention_layernorm.fork(x)
        return self.mlp(ln, residual=x)


class ParallelTransformer(nn.Module):
    def __init__(self, hidden, heads, ffn, num_layers, eps, dtype, layer_type="encoder", attention_dropout=0.0,
                 hidden_dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([ParallelTransformerLayer(hidden, heads, ffn, eps, dtype, layer_type,
                                                              attention_dropout, hidden_dropout)
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
    def __init__(self, hidden, vocab, max_positions, num_tokentypes, dtype, hidden_dropout=0.0):
        super().__init__()
        self.word_embeddings = _Table(vocab, hidden, dtype)
        self.position_embeddings = _Table(max_positions, hidden, dtype)
        self.tokentype_embeddings = _Table(num_tokentypes, hidden, dtype) if num_tokentypes > 0 else None
        self.hidden_dropout = float(hidden_dropout)

    def forward(self, input_ids, tokentype_ids=None, pos_ids=None):
        typ = self.tokentype_embeddings.weight if tokentype_ids is not None else None
        if pos_ids is not None:         # token-packed sequences (no-grad forward path): explicit positions
            return ops.embedding(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                                 tokentype_ids, typ, seq=input_ids.numel(), pos_ids=pos_ids)
        x = ag.embedding(input_ids, self.word_embeddings.weight, self.position_embeddings.weight,
                         tokentype_ids, typ)
        if self.training and self.hidden_dropout:          # embedding_dropout (language_model.py:181)
            x = ag.dropout_add(x, None, self.hidden_dropout)
        return x


class TransformerLanguageModel(nn.Module):
    def __init__(self, cfg, num_tokentypes=0, add_decoder=False, vocab_size=None):
        super().__init__()
        d = cfg["dtype"]
        self.hidden = cfg["hidden"]
        # the reference's defaults (arguments.py:218-221); only active in training mode
        drop = dict(attention_dropout=cfg.get("attention_dropout", 0.1), hidden_dropout=cfg.get("hidden_dropout", 0.1))
        self.embedding = Embedding(cfg["hidden"], vocab_size or cfg["vocab"], cfg["max_pos"],
                                   num_tokentypes, d, hidden_dropout=drop["hidden_dropout"])
        self.encoder = ParallelTransformer(cfg["hidden"], cfg["heads"], cfg["ffn"], cfg["layers"],
                                           cfg.get("eps", 1e-5), d, **drop)
        self.add_decoder = add_decoder
        if add_decoder:
            self.decoder = ParallelTransformer(cfg["hidden"], cfg["heads"], cfg["ffn"], cfg["layers"],
                                               cfg.get("eps", 1e-5), d, layer_type="decoder", **drop)

    #: True: attention skips padding (zeros at padding positions, identical results at every
    #: non-padding position); False: the reference's values at padding positions too.
    skip_padding = True

    @staticmethod
    def trimmed_width(s, max_len):
        """Columns to keep when the longest sequence of the batch has max_len tokens (multiple of 64)."""
        if max_len is None:
            return s
        return min(s, max(64, -(-int(max_len) // 64) * 64))

    #: Length-bucketed execution of large batches (see `encode`): number of buckets, the smallest
    #: batch it applies to, and the granularity (tokens) a bucket's width is rounded up to.
    length_buckets = 4
    bucket_min_rows = 64
    bucket_granularity = 8

    def _bucket_plan(self, b, s, row_lengths):
        """None, or [(row indices (host int64 array), width)] covering every row once, sorted by
        length.  Pure host arithmetic on lengths the caller already has on the host."""
        import numpy as np
        if row_lengths is None or self.length_buckets < 2 or b < max(self.bucket_min_rows, 2 * self.length_buckets):
            return None
        lengths = np.asarray(row_lengths).reshape(-1)
        if lengths.shape[0] != b:
            raise ValueError("row_lengths must hold one length per sequence (%d != %d)" % (lengths.shape[0], b))
        order = np.argsort(lengths, kind="stable")
        gran = self.bucket_granularity
        plan = []
        for idx in np.array_split(order, self.length_buckets):
            if idx.size == 0:
                continue
            w = int(min(s, max(gran, -(-int(lengths[idx].max()) // gran) * gran)))
            if plan and plan[-1][1] == w:
                plan[-1] = (np.concatenate([plan[-1][0], idx]), w)
            else:
                plan.append((idx, w))
        return plan if len(plan) > 1 else None

    def encode(self, ids, tokentype_ids=None, max_len=None, row_lengths=None, cls_only=False):
        """Encoder states [b, s', h] (or only the position-0 states [b, h] with cls_only).

        max_len (the longest non-padding length, known to the caller on the host): only the first
        s' = roundup(max_len, 64) columns are computed — every dropped column is padding in every
        sequence, so the kept positions are unchanged.

        row_lengths (host array, one non-padding length per sequence): large batches additionally
        run length-bucketed — sequences are sorted by length into `length_buckets` groups, each cut
        to its own longest member, and the groups' tokens are packed into ONE [T, h] activation
        matrix: every GEMM / LayerNorm runs once over T rows (T ~ 0.85-0.9 of b*s' on NQ-shaped
        batches; no wave-quantisation loss from splitting), only attention runs per group on its
        slice.  Cut columns are padding in every member of the group, so non-padding positions are
        unchanged; the cut columns of the returned states are zero (no consumer reads padding
        positions: the decoder masks them through the ids)."""
        if self._use_packed(ids, row_lengths):
            return self._encode_packed(ids, tokentype_ids, row_lengths, cls_only)
        s_keep = self.trimmed_width(ids.shape[1], max_len)
        if s_keep != ids.shape[1]:
            ids = ids[:, :s_keep].contiguous()
            if tokentype_ids is not None:
                tokentype_ids = tokentype_ids[:, :s_keep].contiguous()
        b, s = ids.shape
        plan = self._bucket_plan(b, s, row_lengths)
        if plan is not None:
            return self._encode_bucketed(ids, tokentype_ids, plan, cls_only)
        pad = (ids < 1).to(torch.uint8)
        x = self.embedding(ids, tokentype_ids)
        q_live = ops.live_blocks(pad) if self.skip_padding else None
        y = self.encoder(x, b, s, pad, q_live=q_live).view(b, s, self.hidden)
        return y[:, 0, :] if cls_only else y

    #: Token-packed (variable-length) execution of the no-grad forward path (emdr2_b200/packed.py): applies when
    #: the caller hands over per-row lengths and no gradient is being recorded.
    packed_varlen = True
    packed_min_rows = 2

    def _use_packed(self, ids, row_lengths):
        return (self.packed_varlen and row_lengths is not None and ids.shape[0] >= self.packed_min_rows
                and not torch.is_grad_enabled() and not self.training)

    def _encode_packed(self, ids, tokentype_ids, row_lengths, cls_only):
        """Encoder over the b sequences laid back to back: [T, h] activations, T = sum of the lengths.  Returns the
        position-0 states [b, h] (cls_only) or a packed.PackedStates."""
        from .packed import PackedBatch, PackedStates
        b, s = ids.shape
        heads = self.encoder.layers[0].self_attention.heads
        pb = PackedBatch(row_lengths, s, heads, ids.device)
        flat_ids = ids.reshape(-1).index_select(0, pb.gather)
        flat_types = tokentype_ids.reshape(-1).index_select(0, pb.gather) if tokentype_ids is not None else None
        x = self.embedding(flat_ids, flat_types, pos_ids=pb.pos_ids)
        y = self.encoder(x, None, None, None, packed=pb)                       # [T, h]
        if cls_only:
            return y.index_select(0, pb.first_rows)
        return PackedStates(y, pb.lens, pb.cu)

    def decode_packed(self, dec_ids, states, group):
        """Decoder over token-packed encoder states (packed.PackedStates): question q attends the `group`
        consecutive sequences [q*group, (q+1)*group) — the FiD concatenation (emdr2_model.py:159-164) without the
        padding.  Decoder self-attention keeps the rectangular kernel (b x L is tiny)."""
        b, sq = dec_ids.shape
        heads = self.decoder.layers[0].self_attention.heads
        plan = states.cross_plan(group, sq, heads)
        if plan.n_sets != b:
            raise ValueError("%d key sets for %d decoder rows" % (plan.n_sets, b))
        x = self.embedding(dec_ids)
        dec_pad = dec_ids < 1
        q_live = ops.live_blocks(dec_pad) if self.skip_padding else None
        y = self.decoder(x, b, sq, dec_pad, causal=True, encoder_output=states.states, q_live=q_live,
                         cross_plan=plan)
        return y.view(b, sq, self.hidden)

    def _encode_bucketed(self, ids, tokentype_ids, plan, cls_only):
        import numpy as np
        b, s = ids.shape
        dev = ids.device
        order = np.concatenate([idx for idx, _ in plan])
        perm = torch.from_numpy(order).pin_memory().to(dev, non_blocking=True)
        ids_sorted = ids.index_select(0, perm)
        typ_sorted = tokentype_ids.index_select(0, perm) if tokentype_ids is not None else None
        xs, groups, spans = [], [], []
        off = r0 = 0
        for idx, w in plan:
            n = int(idx.size)
            ids_g = ids_sorted[r0:r0 + n, :w].contiguous()
            typ_g = typ_sorted[r0:r0 + n, :w].contiguous() if typ_sorted is not None else None
            pad_g = (ids_g < 1).to(torch.uint8)          # converted once, not once per layer
            xs.append(self.embedding(ids_g, typ_g))
            groups.append((off, n, w, pad_g, ops.live_blocks(pad_g) if self.skip_padding else None))
            spans.append((off, r0, n, w))
            off += n * w
            r0 += n
        y = self.encoder(torch.cat(xs, dim=0), None, None, None, groups=groups)      # [T, h]
        h = self.hidden
        if cls_only:
            first = torch.cat([y[o:o + n * w].view(n, w, h)[:, 0, :] for o, _, n, w in spans], dim=0)
            out = torch.empty_like(first)
            out[perm] = first                      # back to the caller's row order
            return out
        out = y.new_zeros((b, s, h))
        for o, r, n, w in spans:
            out[perm[r:r + n], :w] = y[o:o + n * w].view(n, w, h)
        return out

    def decode_incremental(self, dec_ids, enc_states, enc_pad, cache):
        """Decoder states [rows, n_new, h] of the positions not yet in `cache` (dec_ids [rows, t_total];
        rows = b * r, see DecoderCache).  Position-for-position the same arithmetic as `decode` — GEMM and
        LayerNorm rows are independent, a query's softmax runs over the same keys in the same order — but
        each token passes through the stack once and the encoder states are projected once per layer."""
        rows, t_total = dec_ids.shape
        packed_states = enc_states if hasattr(enc_states, "cross_plan") else None
        if packed_states is not None:
            b, sk = len(packed_states.lens) // cache.group, None
        else:
            b, sk = enc_states.shape[0], enc_states.shape[1]
        if rows % b:
            raise ValueError("decoder rows (%d) must be a multiple of the question batch (%d)" % (rows, b))
        r = rows // b
        t0 = cache.t
        n_new = t_total - t0
        if n_new < 1 or t_total > cache.max_len:
            raise ValueError("nothing new to decode, or past the cache's %d positions" % cache.max_len)
        if n_new > 1 and t0 > 0:
            raise ValueError("several new tokens are only supported on the first call (prefix)")
        h, heads, dev = self.hidden, self.decoder.layers[0].self_attention.heads, dec_ids.device
        dtype = self.embedding.word_embeddings.weight.dtype
        if cache.cross_kv is None:
            enc2d = packed_states.states if packed_states is not None else enc_states.reshape(b * sk, h)
            cache.cross_kv = [layer.inter_attention.key_value(enc2d) for layer in self.decoder.layers]
            if packed_states is None:
                cache.enc_pad = enc_pad.to(torch.uint8).contiguous()
                cache.enc_live = ops.live_blocks(enc_pad) if self.skip_padding else None
        if cache.self_kv is None or cache.self_kv[0].shape[0] != rows:
            if cache.self_kv is not None:
                raise ValueError("hypothesis rows changed without DecoderCache.reorder")
            cache.self_kv = [torch.zeros((rows, cache.max_len, 2 * h), dtype=dtype, device=dev)
                             for _ in self.decoder.layers]
        new_ids = dec_ids[:, t0:].contiguous()
        emb = self.embedding
        x = ops.embedding(new_ids, emb.word_embeddings.weight, emb.position_embeddings.weight[t0:])
        key_pad = (torch.arange(cache.max_len, device=dev) >= t_total).to(torch.uint8)[None].expand(rows, -1).contiguous()
        for li, layer in enumerate(self.decoder.layers):
            att = layer.self_attention
            qkv = att.query_key_value(layer.input_layernorm(x))                    # [rows*n_new, 3h]
            kv = cache.self_kv[li]
            kv[:, t0:t_total] = qkv[:, h:].view(rows, n_new, 2 * h)
            kv2d = kv.view(rows * cache.max_len, 2 * h)
            ctx = ops.attention(qkv[:, :h], kv2d[:, :h], kv2d[:, h:], rows, heads, n_new, cache.max_len,
                                k_pad=key_pad, causal=n_new > 1, scale=att.scale)
            x = att.dense(ctx, residual=x)
            ln = layer.post_attention_layernorm(x)
            cross = layer.inter_attention
            q = cross.query(ln)                                                     # [b * (r*n_new), h]
            if packed_states is not None:
                ctx = ag.cross_attention_packed(q, cache.cross_kv[li], heads,
                                                packed_states.cross_plan(cache.group, r * n_new, heads),
                                                scale=cross.scale)
            else:
                ctx = ag.cross_attention(q, cache.cross_kv[li], b, heads, r * n_new, sk, k_pad=cache.enc_pad,
                                         k_live=cache.enc_live, scale=cross.scale)
            x = cross.dense(ctx, residual=x)
            x = layer.mlp(layer.post_inter_attention_layernorm(x), residual=x)
        cache.t = t_total
        return self.decoder.final_layernorm(x).view(rows, n_new, h)

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


class DecoderCache(object):
    """Incremental decoding state of one question batch (evaluation decode loop, reference
    megatron/model/search_strategy.py:189-240 re-runs the whole decoder — including the cross-attention
    K/V projection of all K*S encoder positions in every layer — for every generated token).

    cross_kv[l]  [b*sk, 2h]       layer l's key/value projection of the encoder states: computed ONCE
    self_kv[l]   [rows, T, 2h]    keys/values of the tokens decoded so far, per hypothesis row
    t            tokens consumed; rows = b * r hypotheses (r = 1 greedy, beam size in beam search; the r
                 hypotheses of a question are consecutive rows and share its encoder states)."""

    def __init__(self, max_len, group=1):
        self.max_len = int(max_len)
        self.group = int(group)          # sequences per key set when the encoder states are token-packed (FiD top-k)
        self.cross_kv = None
        self.self_kv = None
        self.enc_pad = self.enc_live = None
        self.t = 0

    def reorder(self, source):
        """Hypothesis rows after a beam step: new row i continues old row source[i] (rows may grow from b
        to b*k on the first step)."""
        if self.self_kv is not None:
            self.self_kv = [kv.index_select(0, source) for kv in self.self_kv]


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

    def forward(self, input_ids, attention_mask=None, tokentype_ids=None, max_len=None, row_lengths=None):
        _require_cuda(input_ids)
        return self.language_model.encode(input_ids, tokentype_ids, max_len=max_len, row_lengths=row_lengths,
                                          cls_only=True)

    def hidden_states(self, input_ids, tokentype_ids=None):
        return self.language_model.encode(input_ids, tokentype_ids)

    def state_dict_for_save_checkpoint(self, destination=None, prefix='', keep_vars=False):
        """{'language_model': {...}} exactly as PretrainedBertModel writes it (dualencoder_model.py:183-189),
        so the reference's load_state_dict (:191-194) reads checkpoints saved here."""
        return {"language_model": language_model_checkpoint(self.language_model, keep_vars)}

    def load_state_dict(self, state_dict, strict=True):
        if "language_model" in state_dict and isinstance(state_dict["language_model"], dict):
            load_reference_state_dict(self, state_dict, strict)
            return
        return super().load_state_dict(state_dict, strict)


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
                enc_ids_for_mask=None, enc_max_len=None, enc_row_lengths=None, decoder_cache=None, fid_group=None):
        _require_cuda(encoder_input_ids)
        lm = self.language_model
        if enc_hidden_states is None:
            # enc_max_len (opt-in): encoder states come back as [b, s', h] with s' = roundup(max_len, 64); with
            # per-row lengths and no gradient they come back token-packed (packed.PackedStates)
            enc = lm.encode(encoder_input_ids, tokentype_ids, max_len=enc_max_len, row_lengths=enc_row_lengths)
            mask_ids = None if hasattr(enc, "cross_plan") else encoder_input_ids[:, :enc.shape[1]]
        elif hasattr(enc_hidden_states, "cross_plan"):
            enc, mask_ids = enc_hidden_states, None
        else:
            enc = enc_hidden_states.to(lm.embedding.word_embeddings.weight.dtype)
            mask_ids = enc_ids_for_mask if enc_ids_for_mask is not None else encoder_input_ids
        if output_enc_hidden:
            return enc
        if mask_ids is None:                   # token-packed encoder states: `fid_group` sequences per question
            group = len(enc.lens) // decoder_input_ids.shape[0] if fid_group is None else fid_group
            if decoder_cache is not None:
                decoder_cache.group = group
                with torch.no_grad():
                    dec = lm.decode_incremental(decoder_input_ids, enc, None, decoder_cache)
            else:
                dec = lm.decode_packed(decoder_input_ids, enc, group)
        elif decoder_cache is not None:        # evaluation decode loop: only the not-yet-decoded positions
            with torch.no_grad():
                dec = lm.decode_incremental(decoder_input_ids, enc, mask_ids < 1, decoder_cache)
        else:
            dec = lm.decode(decoder_input_ids, enc, mask_ids < 1)
        b, sq, h = dec.shape
        word = lm.embedding.word_embeddings.weight
        logits = ag.linear(dec.reshape(b * sq, h), word, self.lm_head.bias).view(b, sq, word.shape[0])
        if lm_labels is None:
            return logits, enc
        loss = -ag.token_logprob(logits, lm_labels)          # vocab_parallel_cross_entropy (t5_model.py:139-146)
        return loss, enc

    def state_dict_for_save_checkpoint(self, destination=None, prefix='', keep_vars=False):
        """{'language_model': {...}, 'lm_head': {'bias': ..}} as T5Model writes it (t5_model.py:156-168)."""
        return {"language_model": language_model_checkpoint(self.language_model, keep_vars),
                "lm_head": _module_state(self.lm_head, keep_vars)}

    def load_state_dict(self, state_dict, strict=True):
        if "language_model" in state_dict and isinstance(state_dict["language_model"], dict):
            load_reference_state_dict(self, state_dict, strict)
            return
        return super().load_state_dict(state_dict, strict)


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


def _module_state(module, keep_vars=False):
    return {k: (v if keep_vars else v.detach()) for k, v in module.state_dict(keep_vars=True).items()}


def language_model_checkpoint(lm, keep_vars=False):
    """The nested dict TransformerLanguageModel.state_dict_for_save_checkpoint writes
    (language_model.py:367-387, :183-198): 'embedding' -> one flat state dict per table,
    'encoder' / 'decoder' -> the flat state dict of the stack ('layers.0.input_layernorm.weight', ...)."""
    emb = {"word_embeddings": _module_state(lm.embedding.word_embeddings, keep_vars),
           "position_embeddings": _module_state(lm.embedding.position_embeddings, keep_vars)}
    if lm.embedding.tokentype_embeddings is not None:
        emb["tokentype_embeddings"] = _module_state(lm.embedding.tokentype_embeddings, keep_vars)
    out = {"embedding": emb, "encoder": _module_state(lm.encoder, keep_vars)}
    if lm.add_decoder:
        out["decoder"] = _module_state(lm.decoder, keep_vars)
    return out


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
