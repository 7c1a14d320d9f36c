"""Differentiable front end of the block operators: torch.autograd.Function wrappers whose forward
AND backward run in libemdr2_b200.so (csrc/gemm.cu, attention.cu, attention_bwd.cu, rowops*.cu).

This is the autograd of the reference's transformer layer (megatron/model/transformer.py:58-563,
mpu/layers.py:170-363, language_model.py:98-181) with dropout off.  PyTorch only records the graph
and adds gradients where branches meet (the residual stream); every product, softmax, LayerNorm and
embedding gradient is a kernel of this library:

  linear        dX = dY.W (W read in place as an MN-major operand), dW = dY^T.X (both operands
                MN-major, split-K over the tokens into an fp32 buffer), db = column sums
  mlp           h->4h->h with the GeLU forward/backward fused into the GEMM epilogues; the saved
                pre-activation comes out of the first GEMM's epilogue for free
  attention     csrc/attention_bwd.cu (dQ kernel + dK/dV kernel), gradients written straight into
                the fused [tokens, 3h] / [tokens, 2h] projection-gradient buffers
  layernorm, embedding, token_logprob   csrc/rowops_bwd.cu

Each public function below dispatches on torch.is_grad_enabled(): without grad it is the plain
forward op (emdr2_b200/ops.py), so inference pays nothing for the training path.
"""
import ctypes

import torch

from . import _lib, ops

_DT = ops._DTYPES
_SPLIT_TOKENS = 4096        # tokens per split-K slice of a weight-gradient product


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


_SM_COUNT = {}


def _gemm_ctas(device):
    """CTAs a persistent GEMM grid gets on `device`: the SM count, or the trainer's `gemm_max_ctas` cap."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    sms = _SM_COUNT.get(key)
    if sms is None:
        sms = _SM_COUNT[key] = torch.cuda.get_device_properties(key).multi_processor_count
    cap = ops.get_option("gemm_max_ctas")
    return cap if 0 < cap < sms else sms


def _splits(tokens, tiles, ctas=148):
    """Split-K factor for dW.  The persistent grid hands work items (tile, split) round-robin to `ctas` CTAs, so the
    step takes ceil(items / ctas) item-times: pick the factor whose item count fills whole waves (54 tiles x 6
    splits = 324 items ran as 3 waves at 73 % occupancy; x 8 = 432 items fill 2.92), at least ~2 waves unless one
    already fills the machine, slices of >= 4096 tokens.  Every extra split costs a pass of fp32 atomics over the
    output tile (measured ~2-4 % of the product per split at 204 800 tokens: 72 tiles x 15 splits ran 1.15 ms against
    0.77 ms for x 2), hence the penalty per split."""
    max_s = max(1, min(32, (tokens + _SPLIT_TOKENS - 1) // _SPLIT_TOKENS))
    best, best_score = 1, -1.0
    for s in range(1, max_s + 1):
        items = tiles * s
        waves = -(-items // ctas)
        score = items / float(waves * ctas) - 0.015 * s
        if score > best_score:
            best, best_score = s, score
    return best


# ---- gradient sinks ("main grads").  A trainer may attach to a parameter an fp32 buffer `main_grad` (a view of a
# flat all-reduce bucket, emdr2_b200/data_parallel.py).  The backward kernels then ACCUMULATE straight into it —
# the split-K weight-gradient GEMM, the bias column sums, the LayerNorm and embedding backward kernels all add
# into fp32 in place anyway — and autograd is handed None for that parameter: no zero fill, no fp32 -> 16-bit
# cast, no AccumulateGrad add, no later cast back for the optimizer (~4 small launches per parameter per step).
# `_expect` / `_done` count a parameter's uses between forward and backward so that the trainer learns when the
# LAST contribution has landed (`param._on_main_grad(param)`): that is what launches a bucket's all-reduce.
def _sink(param):
    return getattr(param, "main_grad", None)


def _expect(*params):
    for p in params:
        if p is not None and getattr(p, "main_grad", None) is not None:
            p._pending_main_grads = getattr(p, "_pending_main_grads", 0) + 1


def _done(*params):
    for p in params:
        if p is None or getattr(p, "main_grad", None) is None:
            continue
        p._pending_main_grads = getattr(p, "_pending_main_grads", 1) - 1
        if p._pending_main_grads <= 0:
            cb = getattr(p, "_on_main_grad", None)
            if cb is not None:
                cb(p)


def _weight_grad(dy, x, sink=None):
    """dW[n, k] = dy[m, n]^T . x[m, k] in fp32 (split-K): added into `sink` (returns None) or returned in the
    parameter dtype."""
    m, n = dy.shape
    k = x.shape[1]
    tiles = ((n + 127) // 128) * ((k + 255) // 256)
    if sink is not None:
        ops.gemm_ex(dy, x, a_mn=True, b_mn=True, accumulate_into=sink, splits=_splits(m, tiles, _gemm_ctas(dy.device)))
        return None
    acc = torch.zeros((n, k), dtype=torch.float32, device=dy.device)
    ops.gemm_ex(dy, x, a_mn=True, b_mn=True, accumulate_into=acc, splits=_splits(m, tiles, _gemm_ctas(dy.device)))
    return acc.to(dy.dtype)


def _bias_grad(dy, sink=None):
    out = sink if sink is not None else torch.zeros(dy.shape[1], dtype=torch.float32, device=dy.device)
    lib = _lib.load()
    with ops._OnDevice(dy.device):
        _lib.check(lib.emdr2_colsum(_DT[dy.dtype], ops._ptr(dy), dy.stride(0), ops._ptr(out), dy.shape[0],
                                    dy.shape[1], ops._stream(dy.device)), "emdr2_colsum")
    return None if sink is not None else out.to(dy.dtype)


def _c(t):
    return t if t.stride(-1) == 1 and (t.dim() < 2 or t.stride(0) % 8 == 0) else t.contiguous()


# --------------------------------------------------------------------------------------- linear
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        ctx.save_for_backward(x, weight)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.params = (weight, bias)
        _expect(weight, bias)
        return ops.linear(x, weight, bias, residual=residual)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        wp, bp = ctx.params
        dy = _c(dy)
        dx = ops.gemm_ex(dy, weight, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = _weight_grad(dy, x, _sink(wp)) if ctx.needs_input_grad[1] else None
        db = _bias_grad(dy, _sink(bp)) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        _done(wp, bp)
        return dx, dw, db, dres


def linear(x, weight, bias=None, residual=None):
    if _needs_grad(x, weight, bias, residual):
        return _LinearFn.apply(x, weight, bias, residual)
    return ops.linear(x, weight, bias, residual=residual)


class _PackedLinearFn(torch.autograd.Function):
    """y = x W_k^T + b_k for a fused projection whose PARAMETER keeps the reference's [np, hn, splits] row order
    (transformer.py:232-240, checkpoint compatibility) while the kernels use the [splits, np, hn] copy W_k.  The
    permutation is not part of the autograd graph: the weight gradient is produced per split — dW[:, s, :] =
    dy[:, s-th column block]^T x — straight into the parameter's row order (a strided [h, k] view of the
    parameter-shaped buffer), i.e. into its main-grad sink when there is one."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_k, b_k, heads, hn, splits):
        ctx.save_for_backward(x, w_k)
        ctx.params, ctx.layout = (weight, bias), (heads, hn, splits)
        _expect(weight, bias)
        return ops.linear(x, w_k, b_k)

    @staticmethod
    def backward(ctx, dy):
        x, w_k = ctx.saved_tensors
        wp, bp = ctx.params
        heads, hn, splits = ctx.layout
        h, k = heads * hn, x.shape[1]
        dy = _c(dy)
        dx = ops.gemm_ex(dy, w_k, b_mn=True) if ctx.needs_input_grad[0] else None
        sw, sb = _sink(wp), _sink(bp)
        dw32 = sw if sw is not None else torch.zeros((splits * h, k), dtype=torch.float32, device=dy.device)
        per_split = dw32.view(h, splits, k)
        for s_i in range(splits):
            _weight_grad(dy[:, s_i * h:(s_i + 1) * h], x, per_split[:, s_i, :])
        db_k = _bias_grad(dy).float()                                  # [splits * h] in kernel order
        if sb is not None:
            sb.view(h, splits).add_(db_k.view(splits, h).t())
        if sw is not None and sb is not None:
            _done(wp, bp)
            return (dx, None, None) + (None,) * 5
        dw = None if sw is not None else dw32.to(dy.dtype)
        db = None if sb is not None else db_k.view(splits, h).t().reshape(-1).to(dy.dtype)
        _done(wp, bp)
        return (dx, dw, db) + (None,) * 5


def packed_linear(x, weight, bias, w_k, b_k, heads, hn, splits):
    if _needs_grad(x, weight, bias):
        return _PackedLinearFn.apply(x, weight, bias, w_k, b_k, heads, hn, splits)
    return ops.linear(x, w_k, b_k)


# ------------------------------------------------------------------------------------------ mlp
class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        pre = torch.empty((x.shape[0], w1.shape[0]), dtype=x.dtype, device=x.device)
        act = ops.gemm_ex(x, w1, bias=b1, gelu=True, preact_out=pre)
        ctx.save_for_backward(x, w1, w2, pre, act)
        ctx.has_res = residual is not None
        ctx.params = (w1, b1, w2, b2)
        _expect(w1, b1, w2, b2)
        return ops.linear(act, w2, b2, residual=residual)

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2, pre, act = ctx.saved_tensors
        p1, q1, p2, q2 = ctx.params
        dy = _c(dy)
        du = ops.gemm_ex(dy, w2, b_mn=True, gelu_bwd_aux=pre)        # (dy . W2) * GeLU'(pre)
        dw2 = _weight_grad(dy, act, _sink(p2)) if ctx.needs_input_grad[3] else None
        db2 = _bias_grad(dy, _sink(q2)) if ctx.needs_input_grad[4] else None
        dx = ops.gemm_ex(du, w1, b_mn=True) if ctx.needs_input_grad[0] else None
        dw1 = _weight_grad(du, x, _sink(p1)) if ctx.needs_input_grad[1] else None
        db1 = _bias_grad(du, _sink(q1)) if ctx.needs_input_grad[2] else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[5]) else None
        _done(p1, q1, p2, q2)
        return dx, dw1, db1, dw2, db2, dres


def mlp(x, w1, b1, w2, b2, residual=None):
    """residual + (GeLU(x w1^T + b1) w2^T + b2): ParallelMLP + bias-dropout-add at p = 0."""
    if _needs_grad(x, w1, b1, w2, b2, residual):
        return _MlpFn.apply(x, w1, b1, w2, b2, residual)
    return ops.linear(ops.linear(x, w1, b1, gelu=True), w2, b2, residual=residual)


# ---------------------------------------------------------------------------------- dropout-add
class _DropoutAddFn(torch.autograd.Function):
    """out = residual + dropout(y) (bias_dropout_add once the bias is in y, transformer.py:397-419; with
    residual None the embedding dropout of language_model.py:181).  Backward regenerates the mask."""

    @staticmethod
    def forward(ctx, y, residual, spec):
        ctx.spec, ctx.has_res = spec, residual is not None
        return ops.dropout_add(y, residual, spec)

    @staticmethod
    def backward(ctx, dout):
        dout = _c(dout)
        dy = ops.dropout_add(dout, None, ctx.spec) if ctx.needs_input_grad[0] else None
        dres = dout if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return dy, dres, None


def dropout_add(y, residual, p, spec=None):
    """residual + dropout_p(y); p == 0 degenerates to the plain sum (or y)."""
    if spec is None:
        if not p:
            return y if residual is None else y + residual
        from . import dropout as _dropout
        spec = _dropout.STATE.next(p, y.device, y.shape[1])
    if _needs_grad(y, residual):
        return _DropoutAddFn.apply(y, residual, spec)
    return ops.dropout_add(y, residual, spec)


# ------------------------------------------------------------------------------------ layernorm
class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y, mean, rstd = ops.layernorm(x, gamma, beta, eps, return_stats=True)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.params = (gamma, beta)
        _expect(gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        gp, bp = ctx.params
        dy = _c(dy)
        rows, h = x.shape
        dx = torch.empty_like(x)
        sunk = _sink(gp) is not None and _sink(bp) is not None
        dgamma = _sink(gp) if sunk else torch.zeros(h, dtype=torch.float32, device=x.device)
        dbeta = _sink(bp) if sunk else torch.zeros(h, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with ops._OnDevice(x.device):
            _lib.check(lib.emdr2_layernorm_bwd(
                _DT[x.dtype], ops._ptr(dy), dy.stride(0), ops._ptr(x), x.stride(0), ops._ptr(gamma),
                ops._ptr(mean), ops._ptr(rstd), None, 0, ops._ptr(dx), dx.stride(0), ops._ptr(dgamma),
                ops._ptr(dbeta), rows, h, ops._stream(x.device)), "emdr2_layernorm_bwd")
        if sunk:
            _done(gp, bp)
            return dx, None, None, None
        return dx, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None


def layernorm(x, gamma, beta, eps=1e-5):
    if _needs_grad(x, gamma, beta):
        return _LayerNormFn.apply(x, gamma, beta, eps)
    return ops.layernorm(x, gamma, beta, eps)


class _LayerNormForkFn(torch.autograd.Function):
    """(LN(x), x): the sublayer input AND the residual branch of a pre-LN block (transformer.py:524-560 — the hidden
    state feeds the LayerNorm and, unchanged, the bias-dropout-add behind the sublayer).  With two separate consumers
    autograd sums the two gradients of x in an extra elementwise pass over [tokens, h]; here the backward kernel adds
    the residual-branch gradient while it writes dx (`dres` of emdr2_layernorm_bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y, mean, rstd = ops.layernorm(x, gamma, beta, eps, return_stats=True)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.params = (gamma, beta)
        _expect(gamma, beta)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, gamma, mean, rstd = ctx.saved_tensors
        gp, bp = ctx.params
        rows, h = x.shape
        sunk = _sink(gp) is not None and _sink(bp) is not None
        if dy is None:                      # only the residual branch was used
            if sunk:
                _done(gp, bp)
            return dres, None, None, None
        dy = _c(dy)
        dres = None if dres is None else _c(dres)
        dx = torch.empty_like(x)
        dgamma = _sink(gp) if sunk else torch.zeros(h, dtype=torch.float32, device=x.device)
        dbeta = _sink(bp) if sunk else torch.zeros(h, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with ops._OnDevice(x.device):
            _lib.check(lib.emdr2_layernorm_bwd(
                _DT[x.dtype], ops._ptr(dy), dy.stride(0), ops._ptr(x), x.stride(0), ops._ptr(gamma),
                ops._ptr(mean), ops._ptr(rstd), ops._ptr(dres), dres.stride(0) if dres is not None else 0,
                ops._ptr(dx), dx.stride(0), ops._ptr(dgamma), ops._ptr(dbeta), rows, h, ops._stream(x.device)),
                "emdr2_layernorm_bwd")
        if sunk:
            _done(gp, bp)
            return dx, None, None, None
        return dx, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None


def layernorm_fork(x, gamma, beta, eps=1e-5):
    """(LN(x), x) for a pre-LN residual block; see _LayerNormForkFn."""
    if _needs_grad(x) and x.dim() == 2:
        return _LayerNormForkFn.apply(x, gamma, beta, eps)
    return layernorm(x, gamma, beta, eps), x


# ------------------------------------------------------------------------------------ attention
def _attention_bwd(q, k, v, o, dout, dq, dk, dv, batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, causal,
                   scale, lse, dropout=None):
    dvec = torch.empty((batch, heads, sq), dtype=torch.float32, device=q.device)
    lib = _lib.load()
    p = ops._ptr
    drop = dropout.c_args() if dropout is not None else (ctypes.c_float(0.0), ctypes.c_uint64(0), ctypes.c_uint64(0), None)
    with ops._OnDevice(q.device):
        _lib.check(lib.emdr2_attention_bwd_dropout(
            _DT[q.dtype], p(q), q.stride(0), p(k), k.stride(0), p(v), v.stride(0), p(o), o.stride(0),
            p(dout), dout.stride(0), p(dq), dq.stride(0), p(dk), dk.stride(0), p(dv), dv.stride(0),
            batch, heads, sq, sk, p(q_pad), p(k_pad), p(q_live), p(k_live), 1 if causal else 0, float(scale),
            p(lse), p(dvec), *drop, ops._stream(q.device)), "emdr2_attention_bwd")


def _u8(m, device):
    return None if m is None else m.to(device=device, dtype=torch.uint8).contiguous()


class _SelfAttentionFn(torch.autograd.Function):
    """ctx = attention(q, k, v) with q | k | v the three column blocks of one [tokens, 3h] tensor."""

    @staticmethod
    def forward(ctx, qkv, batch, heads, seq, pad, live, causal, scale, dropout):
        h = heads * 64
        pad, live = _u8(pad, qkv.device), _u8(live, qkv.device)
        out, lse = ops.attention(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], batch, heads, seq, seq, q_pad=pad,
                                 k_pad=pad, causal=causal, scale=scale, return_lse=True, q_live=live, k_live=live,
                                 dropout=dropout)
        ctx.save_for_backward(qkv, out, lse, pad, live)
        ctx.cfg = (batch, heads, seq, causal, scale, dropout)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, pad, live = ctx.saved_tensors
        batch, heads, seq, causal, scale, dropout = ctx.cfg
        h = heads * 64
        dout = _c(dout)
        dqkv = torch.empty_like(qkv)
        _attention_bwd(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], out, dout, dqkv[:, :h], dqkv[:, h:2 * h],
                       dqkv[:, 2 * h:], batch, heads, seq, seq, pad, pad, live, live, causal, scale, lse, dropout)
        return dqkv, None, None, None, None, None, None, None, None


class _CrossAttentionFn(torch.autograd.Function):
    """ctx = attention(q, k, v) with k | v the two column blocks of one [batch*sk, 2h] tensor.

    A long key axis with few (batch, head) pairs (FiD: 8 questions x 12 heads over 25 600 keys) is cut into `splits`
    key ranges that run as extra batch entries in the forward AND the backward kernels: 96 CTAs walking 200 key
    blocks each become ~800 work items.  Forward: partial outputs merged by their log-sum-exp weights; the saved lse
    is the merged one.  Backward: every range sees the merged output / lse (so P is the global softmax), dK / dV
    come out per range — the ranges partition the keys — and dQ is the sum of the ranges' partial dQ.  The dropout
    mask plane is indexed by (range entry, head, query, key within the range) in both directions."""

    @staticmethod
    def forward(ctx, q, kv, batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, scale, dropout):
        h = heads * 64
        q_pad, k_pad = _u8(q_pad, q.device), _u8(k_pad, q.device)
        q_live, k_live = _u8(q_live, q.device), _u8(k_live, q.device)
        splits = _cross_splits(batch, heads, sk)
        if splits > 1:
            q_pad, k_pad, q_live, k_live = _split_masks(q_pad, k_pad, q_live, k_live, batch, sk, splits)
            out_s, lse_s = ops.attention(_rep_rows(q, batch, splits, sq), kv[:, :h], kv[:, h:], batch * splits, heads,
                                         sq, sk // splits, q_pad=q_pad, k_pad=k_pad, scale=scale, return_lse=True,
                                         q_live=q_live, k_live=k_live, dropout=dropout)
            out, lse = _merge_splits(out_s, lse_s, batch, splits, heads, sq, q.dtype)
        else:
            out, lse = ops.attention(q, kv[:, :h], kv[:, h:], batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad,
                                     scale=scale, return_lse=True, q_live=q_live, k_live=k_live, dropout=dropout)
        ctx.save_for_backward(q, kv, out, lse, q_pad, k_pad, q_live, k_live)
        ctx.cfg = (batch, heads, sq, sk, scale, dropout, splits)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, out, lse, q_pad, k_pad, q_live, k_live = ctx.saved_tensors
        batch, heads, sq, sk, scale, dropout, splits = ctx.cfg
        h = heads * 64
        dout = _c(dout)
        dkv = torch.empty_like(kv)
        if splits > 1:
            dq_s = torch.empty((batch * splits * sq, h), dtype=q.dtype, device=q.device)
            lse_s = lse.view(batch, 1, heads, sq).expand(batch, splits, heads, sq).contiguous()
            _attention_bwd(_rep_rows(q, batch, splits, sq), kv[:, :h], kv[:, h:], _rep_rows(out, batch, splits, sq),
                           _rep_rows(dout, batch, splits, sq), dq_s, dkv[:, :h], dkv[:, h:], batch * splits, heads, sq,
                           sk // splits, q_pad, k_pad, q_live, k_live, False, scale, lse_s, dropout)
            dq = dq_s.view(batch, splits, sq * h).sum(dim=1, dtype=torch.float32).to(q.dtype).view(batch * sq, h)
        else:
            dq = torch.empty_like(q)
            _attention_bwd(q, kv[:, :h], kv[:, h:], out, dout, dq, dkv[:, :h], dkv[:, h:], batch, heads, sq, sk,
                           q_pad, k_pad, q_live, k_live, False, scale, lse, dropout)
        return (dq, dkv) + (None,) * 10


def _rep_rows(x, batch, splits, sq):
    """[batch*sq, h] -> [batch*splits*sq, h]: every question's rows once per key range."""
    h = x.shape[1]
    return x.reshape(batch, 1, sq, h).expand(batch, splits, sq, h).reshape(batch * splits * sq, h)


def _split_masks(q_pad, k_pad, q_live, k_live, batch, sk, splits):
    """The masks of a key-split launch: per-query masks repeat for the ranges of a question, per-key masks are
    re-cut.  The kernels want one live key block per entry, so block 0 of every range is marked live; in an
    all-padding range its keys are still masked (-10000) and the range weighs exp(-10000) = 0."""
    sk_s = sk // splits
    q_pad = None if q_pad is None else q_pad.repeat_interleave(splits, dim=0)
    q_live = None if q_live is None else q_live.repeat_interleave(splits, dim=0)
    k_pad = None if k_pad is None else k_pad.reshape(batch * splits, sk_s)
    if k_live is not None:
        k_live = k_live.reshape(batch * splits, sk_s // 128).clone()
        k_live[:, 0] = 1
    return q_pad, k_pad, q_live, k_live


def _merge_splits(out_s, lse_s, batch, splits, heads, sq, dtype):
    """Partial outputs [batch*splits*sq, h] and their lse [batch*splits, heads, sq] -> merged output, merged lse."""
    lse_s = lse_s.view(batch, splits, heads, sq)
    lse = torch.logsumexp(lse_s, dim=1)                                          # fp32 [B, heads, sq]
    w = torch.exp(lse_s - lse.unsqueeze(1)).permute(0, 1, 3, 2).unsqueeze(-1)    # [B, splits, sq, heads, 1]
    merged = (out_s.view(batch, splits, sq, heads, 64).float() * w).sum(dim=1)
    return merged.to(dtype).view(batch * sq, heads * 64), lse.contiguous()


class _GroupedSelfAttentionFn(torch.autograd.Function):
    """Self-attention over a token-packed [T, 3h] projection made of several rectangular groups
    (emdr2_b200/blocks.py: length-bucketed execution).  `groups` is a list of
    (row offset, batch, seq, pad uint8 [batch, seq], live uint8 block map or None); one attention
    launch per group in either direction, all reading / writing slices of the same buffers."""

    @staticmethod
    def forward(ctx, qkv, heads, groups, causal, scale, dropouts):
        h = heads * 64
        out = torch.empty((qkv.shape[0], h), dtype=qkv.dtype, device=qkv.device)
        lses = []
        for (off, b, s, pad, live), drop in zip(groups, dropouts):
            part = qkv[off:off + b * s]
            _, lse = ops.attention(part[:, :h], part[:, h:2 * h], part[:, 2 * h:], b, heads, s, s, q_pad=pad,
                                   k_pad=pad, causal=causal, scale=scale, return_lse=True, q_live=live,
                                   k_live=live, out=out[off:off + b * s], dropout=drop)
            lses.append(lse)
        ctx.save_for_backward(qkv, out, *lses)
        ctx.groups, ctx.cfg = groups, (heads, causal, scale, dropouts)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out = ctx.saved_tensors[:2]
        lses = ctx.saved_tensors[2:]
        heads, causal, scale, dropouts = ctx.cfg
        h = heads * 64
        dout = _c(dout)
        dqkv = torch.empty_like(qkv)
        for (off, b, s, pad, live), lse, drop in zip(ctx.groups, lses, dropouts):
            r = slice(off, off + b * s)
            q, dq = qkv[r], dqkv[r]
            _attention_bwd(q[:, :h], q[:, h:2 * h], q[:, 2 * h:], out[r], dout[r], dq[:, :h], dq[:, h:2 * h],
                           dq[:, 2 * h:], b, heads, s, s, pad, pad, live, live, causal, scale, lse, drop)
        return dqkv, None, None, None, None, None


def _attn_spec(p, device, keys):
    """A fresh DropoutSpec for one attention launch over `keys` key columns (None when p == 0)."""
    from . import dropout as _dropout
    return _dropout.STATE.next(p, device, keys) if p else None


def self_attention_grouped(qkv, heads, groups, causal=False, scale=0.125, dropout_p=0.0):
    """ctx [T, h] for a packed projection qkv [T, 3h] whose rows are the concatenation of the groups'
    [batch*seq] token blocks.  groups: (row offset, batch, seq, pad, live) per group."""
    groups = [(int(off), int(b), int(s), _u8(pad, qkv.device), _u8(live, qkv.device))
              for off, b, s, pad, live in groups]
    dropouts = [_attn_spec(dropout_p, qkv.device, g[2]) for g in groups]     # one mask plane per launch
    if _needs_grad(qkv):
        return _GroupedSelfAttentionFn.apply(qkv, heads, groups, causal, scale, dropouts)
    h = heads * 64
    out = torch.empty((qkv.shape[0], h), dtype=qkv.dtype, device=qkv.device)
    for (off, b, s, pad, live), drop in zip(groups, dropouts):
        part = qkv[off:off + b * s]
        ops.attention(part[:, :h], part[:, h:2 * h], part[:, 2 * h:], b, heads, s, s, q_pad=pad, k_pad=pad,
                      causal=causal, scale=scale, q_live=live, k_live=live, out=out[off:off + b * s], dropout=drop)
    return out


def self_attention(qkv, batch, heads, seq, pad=None, live=None, causal=False, scale=0.125, dropout_p=0.0,
                   dropout=None):
    """dropout_p draws a fresh mask; `dropout` (a DropoutSpec) replays a given one."""
    drop = dropout if dropout is not None else _attn_spec(dropout_p, qkv.device, seq)
    if _needs_grad(qkv):
        return _SelfAttentionFn.apply(qkv, batch, heads, seq, pad, live, causal, scale, drop)
    h = heads * 64
    return ops.attention(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], batch, heads, seq, seq, q_pad=pad, k_pad=pad,
                         causal=causal, scale=scale, q_live=live, k_live=live, dropout=drop)


#: Key-split ("flash-decoding") of long cross-attention in the no-grad path: with few (batch, head)
#: pairs and a very long key axis (FiD: 8 questions x 12 heads over 50 x 512 = 25 600 keys) one CTA per
#: pair would walk 200 key blocks while two thirds of the GPU idle.  The key axis is cut into
#: `splits` contiguous ranges that run as extra batch entries of the same kernel; the partial outputs
#: are merged with their log-sum-exp weights.  Applies when sk >= CROSS_SPLIT_MIN_KEYS.
CROSS_SPLIT_MIN_KEYS = 4096
CROSS_SPLIT_TARGET_ITEMS = 888        # ~3 work items per resident CTA (2 x 148)


def _cross_splits(batch, heads, sk):
    """Number of key ranges (a divisor of the 128-key block count), 1 = do not split."""
    if sk < CROSS_SPLIT_MIN_KEYS or sk % 128 or batch * heads * 2 > CROSS_SPLIT_TARGET_ITEMS:
        return 1
    nblk = sk // 128
    want = max(1, CROSS_SPLIT_TARGET_ITEMS // (batch * heads))
    best = 1
    for s in range(2, min(nblk // 4, 64) + 1):          # keep at least 4 key blocks per range
        if nblk % s == 0 and abs(s - want) < abs(best - want):
            best = s
    return best


def _cross_attention_split(q, kv, batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, scale, splits, dropout=None):
    h = heads * 64
    dev = q.device
    q_pad, k_pad, q_live, k_live = _split_masks(_u8(q_pad, dev), _u8(k_pad, dev), _u8(q_live, dev), _u8(k_live, dev),
                                                batch, sk, splits)
    out, lse = ops.attention(_rep_rows(q, batch, splits, sq), kv[:, :h], kv[:, h:], batch * splits, heads, sq,
                             sk // splits, q_pad=q_pad, k_pad=k_pad, scale=scale, return_lse=True, q_live=q_live,
                             k_live=k_live, dropout=dropout)
    return _merge_splits(out, lse, batch, splits, heads, sq, q.dtype)[0]


def cross_attention(q, kv, batch, heads, sq, sk, q_pad=None, k_pad=None, q_live=None, k_live=None, scale=0.125,
                    dropout_p=0.0, dropout=None):
    """A long key axis runs key-split (see _CrossAttentionFn) with or without autograd; a dropout mask plane is then
    indexed by (range entry, head, query, key within the range)."""
    drop = dropout if dropout is not None else _attn_spec(dropout_p, q.device, sk)
    if _needs_grad(q, kv):
        return _CrossAttentionFn.apply(q, kv, batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, scale, drop)
    h = heads * 64
    splits = _cross_splits(batch, heads, sk)
    if splits > 1:
        return _cross_attention_split(q, kv, batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, scale, splits, drop)
    return ops.attention(q, kv[:, :h], kv[:, h:], batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad, scale=scale,
                         q_live=q_live, k_live=k_live, dropout=drop)


def cross_attention_packed(q, kv, heads, plan, scale=0.125):
    """FiD cross-attention over TOKEN-PACKED encoder states (no-grad forward path): q [n_sets * sq, h], kv [T, 2h]
    = the key/value projection of the packed states, plan = packed.CrossPlan.  Every key set is cut into ranges
    that run as separate work items of one varlen launch; the partial outputs are merged with their
    log-sum-exp weights (an absent range weighs exp(-inf) = 0)."""
    h = heads * 64
    sq, n_slots = plan.sq, plan.n_slots
    part = torch.zeros(((n_slots + 1) * sq, h), dtype=q.dtype, device=q.device)
    lse = torch.full(((n_slots + 1) * heads * sq,), float("-inf"), dtype=torch.float32, device=q.device)
    ops.attention_varlen(q, kv[:, :h], kv[:, h:], heads, plan.items, plan.n_items, scale=scale, out=part, lse=lse,
                         flops=plan.attention_flops)
    if plan.max_chunks == 1 and n_slots == plan.n_sets:
        return part[:n_slots * sq]
    idx = plan.slot_index.reshape(-1).long()
    w = torch.softmax(lse.view(n_slots + 1, heads, sq).index_select(0, idx)
                      .view(plan.n_sets, plan.max_chunks, heads, sq), dim=1)                    # [sets, chunks, heads, sq]
    o = part.view(n_slots + 1, sq, heads, 64).index_select(0, idx).view(plan.n_sets, plan.max_chunks, sq, heads, 64)
    merged = (o.float() * w.permute(0, 1, 3, 2).unsqueeze(-1)).sum(dim=1)
    return merged.to(q.dtype).view(plan.n_sets * sq, h)


# ------------------------------------------------------------------------------------ embedding
class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, word, pos, types, type_emb):
        ctx.save_for_backward(ids, types)
        ctx.shapes = (word.shape, pos.shape, None if type_emb is None else type_emb.shape, word.dtype)
        ctx.params = (word, pos, type_emb)
        _expect(word, pos, type_emb)
        return ops.embedding(ids, word, pos, types, type_emb)

    @staticmethod
    def backward(ctx, dx):
        ids, types = ctx.saved_tensors
        wshape, pshape, tshape, dtype = ctx.shapes
        dx = dx.contiguous()
        dev = dx.device
        wp, pp, tp = ctx.params
        sunk = _sink(wp) is not None and _sink(pp) is not None and (tp is None or _sink(tp) is not None)
        if sunk:
            dword, dpos = _sink(wp), _sink(pp)
            dtyp = _sink(tp) if (tp is not None and types is not None) else None
        else:
            dword = torch.zeros(wshape, dtype=torch.float32, device=dev)
            dpos = torch.zeros(pshape, dtype=torch.float32, device=dev)
            dtyp = torch.zeros(tshape, dtype=torch.float32, device=dev) if (tshape is not None and types is not None) else None
        ids2 = ids.to(torch.int64).contiguous()
        ty2 = None if types is None else types.to(torch.int64).contiguous()
        lib = _lib.load()
        with ops._OnDevice(dev):
            _lib.check(lib.emdr2_embedding_bwd(
                _DT[dtype], ops._ptr(dx), ops._ptr(ids2), ops._ptr(ty2), ops._ptr(dword), ops._ptr(dpos),
                ops._ptr(dtyp), ids2.numel(), ids2.shape[-1], wshape[1], wshape[0],
                0 if tshape is None else tshape[0], ops._stream(dev)), "emdr2_embedding_bwd")
        if sunk:
            _done(wp, pp, tp)
            return None, None, None, None, None
        return (None, dword.to(dtype), dpos.to(dtype), None,
                None if tshape is None else (dtyp.to(dtype) if dtyp is not None else torch.zeros(tshape, dtype=dtype, device=dev)))


def embedding(ids, word, pos, types=None, type_emb=None):
    if _needs_grad(word, pos, type_emb):
        return _EmbeddingFn.apply(ids, word, pos, types, type_emb if types is not None else None)
    return ops.embedding(ids, word, pos, types, type_emb)


# -------------------------------------------------------------------------------- token_logprob
class _TokenLogprobFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lp, lse = ops.token_logprob(logits, labels)
        ctx.save_for_backward(logits, labels, lse)
        return lp

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse = ctx.saved_tensors
        vocab = logits.shape[-1]
        l2 = logits.reshape(-1, vocab)
        if l2.stride(1) != 1:
            l2 = l2.contiguous()
        lab = labels.to(torch.int64).reshape(-1).contiguous()
        gg = g.to(torch.float32).reshape(-1).contiguous()
        dl = torch.empty((l2.shape[0], vocab), dtype=logits.dtype, device=logits.device)
        lib = _lib.load()
        with ops._OnDevice(logits.device):
            _lib.check(lib.emdr2_token_logprob_bwd(
                _DT[logits.dtype], ops._ptr(l2), max(vocab, l2.stride(0)), ops._ptr(lab), ops._ptr(lse.reshape(-1).contiguous()),
                ops._ptr(gg), ops._ptr(dl), vocab, l2.shape[0], vocab, ops._stream(logits.device)),
                "emdr2_token_logprob_bwd")
        return dl.view(logits.shape), None


def token_logprob(logits, labels):
    """log p(label) per row, differentiable w.r.t. the logits."""
    if _needs_grad(logits):
        return _TokenLogprobFn.apply(logits, labels)
    return ops.token_logprob(logits, labels)[0]
