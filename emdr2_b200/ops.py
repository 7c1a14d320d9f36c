"""Torch-tensor front end of the C-ABI transformer-block operators (csrc/gemm.cu, ...).

PyTorch owns the memory and the stream; all arithmetic runs in libemdr2_b200.so.  No CPU path.
"""
import ctypes

import torch

from . import _lib

_DTYPES = {torch.float16: _lib.EMDR2_DTYPE_FP16, torch.bfloat16: _lib.EMDR2_DTYPE_BF16}
GEMM_BIAS, GEMM_GELU, GEMM_RESIDUAL = 1, 2, 4


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    """Raw handle of the calling thread's current stream on `device` (no Stream object: this runs once per launch)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(device.index if device.index is not None
                                                              else torch.cuda.current_device()))


class _OnDevice(object):
    """`with torch.cuda.device(d)` only when d is not already current (the usual case: one device per process,
    one launch = one of these): entering the real context manager costs more host time than the launch itself."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _check_2d(name, t, dtype, device):
    if t.dim() != 2 or t.dtype != dtype or t.device != device or t.stride(1) != 1:
        raise ValueError("%s must be a 2-D %s tensor on %s with unit inner stride" % (name, dtype, device))


def linear(x, weight, bias=None, gelu=False, residual=None, out=None):
    """out[m, n] = (GeLU)(x[m, k] @ weight[n, k].T + bias) + residual, fp32 accumulate.

    x / residual / out may be row-strided views (e.g. a column block of a wider buffer)."""
    if x.dtype not in _DTYPES:
        raise TypeError("linear takes float16/bfloat16 tensors, got %s" % x.dtype)
    if not x.is_cuda:
        raise RuntimeError("emdr2_b200 has no CPU path")
    dtype, device = x.dtype, x.device
    _check_2d("x", x, dtype, device)
    _check_2d("weight", weight, dtype, device)
    m, k = x.shape
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise ValueError("weight is [%d, %d], expected [n, %d]" % (weight.shape[0], weight.shape[1], k))
    if out is None:
        out = torch.empty((m, n), dtype=dtype, device=device)
    _check_2d("out", out, dtype, device)
    flags = 0
    if bias is not None:
        if bias.dtype != dtype or bias.device != device or bias.numel() != n or not bias.is_contiguous():
            raise ValueError("bias must be a contiguous [n] tensor of the same dtype/device")
        flags |= GEMM_BIAS
    if gelu:
        flags |= GEMM_GELU
    ldr = 0
    if residual is not None:
        _check_2d("residual", residual, dtype, device)
        if tuple(residual.shape) != (m, n):
            raise ValueError("residual must be [m, n]")
        flags |= GEMM_RESIDUAL
        ldr = residual.stride(0)
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_gemm(_DTYPES[dtype], _ptr(x), x.stride(0) if m > 1 else max(k, x.stride(0)),
                                  _ptr(weight), weight.stride(0) if n > 1 else max(k, weight.stride(0)),
                                  _ptr(out), out.stride(0) if m > 1 else max(n, out.stride(0)),
                                  _ptr(bias), _ptr(residual), ldr, m, n, k, flags, _stream(device)),
                   "emdr2_gemm")
    return out


def live_blocks(pad, block=128):
    """uint8 [batch, ceil(s/block)]: 1 where the block holds at least one non-padding position of the
    bool/uint8 padding mask [batch, s]; block 0 is always marked live.  This is what
    attention(..., q_live=, k_live=) takes to skip all-padding blocks."""
    b, s = pad.shape
    nblk = -(-s // block)
    live = (pad == 0)
    if nblk * block != s:
        live = torch.nn.functional.pad(live, (0, nblk * block - s))
    live = live.view(b, nblk, block).any(dim=2)
    live[:, 0] = True
    return live.to(torch.uint8).contiguous()


def dropout_add(y, residual, spec, out=None):
    """out = residual + dropout(y) (residual may be None; out may be y): csrc/rowops.cu dropout_add_kernel.
    `spec` is an emdr2_b200.dropout.DropoutSpec; the same spec regenerates the same mask."""
    dtype, device = y.dtype, y.device
    if dtype not in _DTYPES or not y.is_cuda:
        raise TypeError("dropout_add takes CUDA float16/bfloat16 tensors")
    _check_2d("y", y, dtype, device)
    rows, cols = y.shape
    if residual is not None:
        _check_2d("residual", residual, dtype, device)
        if tuple(residual.shape) != (rows, cols):
            raise ValueError("residual must have y's shape")
    if out is None:
        out = torch.empty((rows, cols), dtype=dtype, device=device)
    _check_2d("out", out, dtype, device)
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_dropout_add(
            _DTYPES[dtype], _ptr(y), max(cols, y.stride(0)), _ptr(residual),
            0 if residual is None else max(cols, residual.stride(0)), _ptr(out), max(cols, out.stride(0)), rows, cols,
            *spec.c_args(), _stream(device)), "emdr2_dropout_add")
    return out


def attention(q, k, v, batch, heads, sq, sk, q_pad=None, k_pad=None, causal=False, scale=None,
              out=None, return_lse=False, q_live=None, k_live=None, dropout=None):
    """Fused attention forward (head dim 64).  q/out: [batch*sq, >= heads*64] views, k/v:
    [batch*sk, >= heads*64] views (unit inner stride; e.g. column blocks of a fused QKV buffer).
    q_pad [batch, sq] / k_pad [batch, sk]: uint8/bool, 1 = padding.  Masked scores are replaced by
    -10000 (the reference's attention_mask_func), not -inf.  dropout: a DropoutSpec (attention dropout on the
    normalised probabilities, transformer.py:345-346) or None.  q_live / k_live (uint8 block maps, see
    live_blocks) switch on padding skipping: identical results at non-padding queries, zeros at
    all-padding query blocks."""
    dtype, device = q.dtype, q.device
    if dtype not in _DTYPES or not q.is_cuda:
        raise TypeError("attention takes CUDA float16/bfloat16 tensors")
    width = heads * 64
    for name, t, rows in (("q", q, batch * sq), ("k", k, batch * sk), ("v", v, batch * sk)):
        _check_2d(name, t, dtype, device)
        if t.shape[0] != rows or t.shape[1] != width:
            raise ValueError("%s must be [%d, %d], got %s" % (name, rows, width, tuple(t.shape)))
    if out is None:
        out = torch.empty((batch * sq, width), dtype=dtype, device=device)
    _check_2d("out", out, dtype, device)
    masks = []
    for name, m, n in (("q_pad", q_pad, sq), ("k_pad", k_pad, sk)):
        if m is not None:
            m = m.to(device=device, dtype=torch.uint8).contiguous()
            if tuple(m.shape) != (batch, n):
                raise ValueError("%s must be [batch, %d]" % (name, n))
        masks.append(m)
    # all-padding query blocks (q_live == 0) are stored as zeros and never write their lse: define it (0)
    # so that consumers that combine partial results by lse (autograd._cross_attention_split) stay finite
    lse = None
    if return_lse:
        lse = (torch.zeros if q_live is not None else torch.empty)((batch, heads, sq), dtype=torch.float32, device=device)
    if scale is None:
        scale = 1.0 / 8.0
    lib = _lib.load()
    drop = dropout.c_args() if dropout is not None else (ctypes.c_float(0.0), ctypes.c_uint64(0), ctypes.c_uint64(0), None)
    with _OnDevice(device):
        _lib.check(lib.emdr2_attention_fwd_dropout(
            _DTYPES[dtype], _ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0),
            _ptr(out), out.stride(0), batch, heads, sq, sk, _ptr(masks[0]), _ptr(masks[1]),
            _ptr(q_live), _ptr(k_live), 1 if causal else 0, float(scale), _ptr(lse), *drop, _stream(device)),
            "emdr2_attention_fwd")
    return (out, lse) if return_lse else out


def attention_varlen(q, k, v, heads, items, n_items, scale=0.125, out=None, out_rows=None, lse=None, flops=0.0):
    """Variable-length attention over token-packed matrices (csrc/attention_varlen.cu): q [Tq, >= heads*64],
    k / v [Tk, >= heads*64] (unit inner stride, e.g. column blocks of a fused projection); `items` int32 [n, 8]
    on the device (emdr2_b200/packed.py builds them).  out: [out_rows, heads*64] (default: q's rows).  Rows no
    item covers are left untouched."""
    dtype, device = q.dtype, q.device
    if dtype not in _DTYPES or not q.is_cuda:
        raise TypeError("attention_varlen takes CUDA float16/bfloat16 tensors")
    width = heads * 64
    for name, t in (("q", q), ("k", k), ("v", v)):
        _check_2d(name, t, dtype, device)
        if t.shape[1] != width:
            raise ValueError("%s must have %d columns" % (name, width))
    if k.shape[0] != v.shape[0]:
        raise ValueError("k and v must have the same number of rows")
    if out is None:
        out = torch.empty((q.shape[0] if out_rows is None else out_rows, width), dtype=dtype, device=device)
    _check_2d("out", out, dtype, device)
    if items.dtype != torch.int32 or items.device != device or not items.is_contiguous():
        raise ValueError("items must be a contiguous int32 tensor on %s" % (device,))
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_attention_varlen_fwd(
            _DTYPES[dtype], _ptr(q), q.stride(0), q.shape[0], _ptr(k), k.stride(0), _ptr(v), v.stride(0), k.shape[0],
            _ptr(out), out.stride(0), out.shape[0], heads, _ptr(items), int(n_items), float(scale), _ptr(lse),
            _stream(device)), "emdr2_attention_varlen_fwd")
    if flops:
        timing_add_flops(KIND_ATTENTION, flops)
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None, return_stats=False):
    """Row-wise LayerNorm of a 2-D [rows, h] view (h % 8 == 0, h <= 1024), fp32 statistics."""
    dtype, device = x.dtype, x.device
    if dtype not in _DTYPES or not x.is_cuda:
        raise TypeError("layernorm takes CUDA float16/bfloat16 tensors")
    _check_2d("x", x, dtype, device)
    rows, h = x.shape
    if out is None:
        out = torch.empty((rows, h), dtype=dtype, device=device)
    _check_2d("out", out, dtype, device)
    mean = torch.empty(rows, dtype=torch.float32, device=device) if return_stats else None
    rstd = torch.empty(rows, dtype=torch.float32, device=device) if return_stats else None
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_layernorm_fwd(
            _DTYPES[dtype], _ptr(x), max(h, x.stride(0)), _ptr(gamma.contiguous()), _ptr(beta.contiguous()),
            _ptr(out), max(h, out.stride(0)), rows, h, float(eps), _ptr(mean), _ptr(rstd), _stream(device)),
            "emdr2_layernorm_fwd")
    return (out, mean, rstd) if return_stats else out


def embedding(ids, word, pos, types=None, type_emb=None, seq=None, pos_ids=None):
    """out[b*s + i] = word[ids[b,i]] + pos[i] (+ type_emb[types[b,i]]); ids int64 [b, s].  pos_ids (int32, one per
    token): explicit positions for token-packed sequences instead of i."""
    dtype, device = word.dtype, word.device
    if dtype not in _DTYPES or not word.is_cuda:
        raise TypeError("embedding takes CUDA float16/bfloat16 tables")
    ids = ids.to(device=device, dtype=torch.int64).contiguous()
    if seq is None:
        seq = ids.shape[-1]
    tokens = ids.numel()
    h = word.shape[1]
    if pos_ids is None and pos.shape[0] < seq:
        raise ValueError("sequence length %d exceeds the position table (%d rows)" % (seq, pos.shape[0]))
    if pos_ids is not None and (pos_ids.dtype != torch.int32 or pos_ids.numel() != tokens or not pos_ids.is_contiguous()):
        raise ValueError("pos_ids must be a contiguous int32 tensor with one entry per token")
    if types is not None:
        types = types.to(device=device, dtype=torch.int64).contiguous()
    out = torch.empty((tokens, h), dtype=dtype, device=device)
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_embedding_fwd_pos(
            _DTYPES[dtype], _ptr(ids), _ptr(types), _ptr(word.contiguous()), _ptr(pos.contiguous()),
            _ptr(type_emb.contiguous()) if type_emb is not None else None, _ptr(out), tokens, int(seq), h,
            word.shape[0], type_emb.shape[0] if type_emb is not None else 0, _ptr(pos_ids), pos.shape[0],
            _stream(device)), "emdr2_embedding_fwd")
    return out


def token_logprob(logits, labels):
    """(logprob, lse) fp32, shaped like `labels`: logprob = logits[..., label] - logsumexp(logits).
    logits [..., V] 16-bit CUDA (last dim contiguous), labels int64 [...]."""
    dtype, device = logits.dtype, logits.device
    if dtype not in _DTYPES or not logits.is_cuda:
        raise TypeError("token_logprob takes CUDA float16/bfloat16 logits")
    vocab = logits.shape[-1]
    l2 = logits.reshape(-1, vocab)
    if l2.stride(1) != 1:
        l2 = l2.contiguous()
    lab = labels.to(device=device, dtype=torch.int64).reshape(-1).contiguous()
    rows = l2.shape[0]
    if lab.numel() != rows:
        raise ValueError("labels must have one entry per logits row")
    lp = torch.empty(rows, dtype=torch.float32, device=device)
    lse = torch.empty(rows, dtype=torch.float32, device=device)
    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_token_logprob(_DTYPES[dtype], _ptr(l2), max(vocab, l2.stride(0)), _ptr(lab),
                                           _ptr(lp), _ptr(lse), rows, vocab, _stream(device)),
                   "emdr2_token_logprob")
    return lp.view(labels.shape), lse.view(labels.shape)


KIND_GEMM, KIND_ATTENTION, KIND_ROWOP = 0, 1, 2


def set_option(name, value):
    """Process-wide switch of the block operators (include/emdr2_b200.h: emdr2_ops_set_option),
    e.g. set_option("gemm_pair", 1)."""
    _lib.check(_lib.load().emdr2_ops_set_option(name.encode(), int(value)), "emdr2_ops_set_option")


def get_option(name):
    v = ctypes.c_int64()
    _lib.check(_lib.load().emdr2_ops_get_option(name.encode(), ctypes.byref(v)), "emdr2_ops_get_option")
    return v.value


def timing(enable):
    """Bracket every block-operator launch of this thread with CUDA events (measurement aid)."""
    _lib.check(_lib.load().emdr2_ops_timing(1 if enable else 0), "emdr2_ops_timing")


def timing_add_flops(kind, flops):
    """Credit algorithmic work to a kernel kind whose entry point cannot know it (varlen attention: the work is
    in the item list)."""
    _lib.check(_lib.load().emdr2_ops_timing_add_flops(kind, ctypes.c_double(flops)), "emdr2_ops_timing_add_flops")


def timing_read(kind):
    """(seconds, launches, algorithmic flops) of one kernel kind since the last read."""
    ns, n, fl = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    _lib.check(_lib.load().emdr2_ops_timing_read(kind, ctypes.byref(ns), ctypes.byref(n), ctypes.byref(fl)),
               "emdr2_ops_timing_read")
    return ns.value * 1e-9, n.value, fl.value


GEMM_ACCUM_F32, GEMM_GELU_BWD, GEMM_PREACT = 8, 16, 32


def gemm_ex(a, b, a_mn=False, b_mn=False, out=None, bias=None, gelu=False, residual=None,
            gelu_bwd_aux=None, preact_out=None, accumulate_into=None, splits=1):
    """General product out[m, n] = a . b^T with fp32 accumulation.

    a is [m, k] (or [k, m] with a_mn=True), b is [n, k] (or [k, n] with b_mn=True): the *_mn forms
    read a tensor whose ROW index is the contraction index, in place (no transpose copy):
        dX = gemm_ex(dY, W, b_mn=True)                                   # [m,n] . [n,k]
        gemm_ex(dY, X, a_mn=True, b_mn=True, accumulate_into=dW32, splits=8)   # dW += dY^T . X
    accumulate_into: fp32 [m, n] tensor that receives atomic adds (split-K / grad accumulation).
    gelu_bwd_aux: saved pre-activation u [m, n]; out = (a . b^T) * GeLU'(u).
    preact_out: [m, n] 16-bit tensor that also receives a . b^T + bias before the GeLU."""
    dtype, device = a.dtype, a.device
    if dtype not in _DTYPES or not a.is_cuda:
        raise TypeError("gemm_ex takes CUDA float16/bfloat16 tensors")
    _check_2d("a", a, dtype, device)
    _check_2d("b", b, dtype, device)
    k, m = (a.shape if a_mn else (a.shape[1], a.shape[0]))
    kb, n = (b.shape if b_mn else (b.shape[1], b.shape[0]))
    if k != kb:
        raise ValueError("contraction sizes differ: %d vs %d" % (k, kb))
    flags = 0
    aux, ld_aux = None, 0
    if bias is not None:
        flags |= GEMM_BIAS
    if gelu:
        flags |= GEMM_GELU
    if residual is not None:
        _check_2d("residual", residual, dtype, device)
        flags |= GEMM_RESIDUAL
        aux, ld_aux = residual, residual.stride(0)
    if gelu_bwd_aux is not None:
        _check_2d("gelu_bwd_aux", gelu_bwd_aux, dtype, device)
        flags |= GEMM_GELU_BWD
        aux, ld_aux = gelu_bwd_aux, gelu_bwd_aux.stride(0)
    pre, ld_pre = None, 0
    if preact_out is not None:
        _check_2d("preact_out", preact_out, dtype, device)
        flags |= GEMM_PREACT
        pre, ld_pre = preact_out, preact_out.stride(0)
    if accumulate_into is not None:
        if accumulate_into.dtype != torch.float32 or tuple(accumulate_into.shape) != (m, n) \
                or accumulate_into.stride(1) != 1 or accumulate_into.device != device:
            raise ValueError("accumulate_into must be a float32 [m, n] tensor on the same device")
        flags |= GEMM_ACCUM_F32
        out = accumulate_into
    elif out is None:
        out = torch.empty((m, n), dtype=dtype, device=device)
    else:
        _check_2d("out", out, dtype, device)

    def ld(t, inner):
        return t.stride(0) if t.shape[0] > 1 else max(inner, t.stride(0))

    lib = _lib.load()
    with _OnDevice(device):
        _lib.check(lib.emdr2_gemm_ex(
            _DTYPES[dtype], _ptr(a), ld(a, a.shape[1]), 1 if a_mn else 0, _ptr(b), ld(b, b.shape[1]),
            1 if b_mn else 0, _ptr(out), ld(out, n), _ptr(bias), _ptr(aux), ld_aux, _ptr(pre), ld_pre,
            m, n, k, flags, int(splits), _stream(device)), "emdr2_gemm_ex")
    return out
