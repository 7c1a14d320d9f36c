"""Numerics of the transformer-block operators against plain PyTorch fp32 references of the same
op (floating-point kernels: tolerance stated per test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


def _ref_linear(x, w, bias, gelu, residual):
    y = x.float() @ w.float().T
    if bias is not None:
        y = y + bias.float()
    if gelu:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


GEMM_SHAPES = [
    # m, n, k
    (128, 256, 64), (128, 256, 768), (256, 768, 768), (1000, 2304, 768), (4096, 3072, 768),
    (4096, 768, 3072), (300, 520, 72), (1, 8, 8), (129, 264, 200), (2048, 30720, 768),
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_plain(m, n, k, dtype):
    from emdr2_b200.ops import linear
    x, w = _rand((m, k), dtype, 1), _rand((n, k), dtype, 2, scale=k ** -0.5)
    got = linear(x, w).float()
    want = _ref_linear(x, w, None, False, None)
    # fp32 accumulation, one rounding to 16 bits on store: half an ulp of the output format
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.allclose(got, want, rtol=tol, atol=tol), (got - want).abs().max().item()


@pytest.mark.parametrize("bias,gelu,res", [(True, False, False), (True, True, False),
                                            (True, False, True), (False, False, True),
                                            (True, True, True)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_fused_epilogues(bias, gelu, res, dtype):
    from emdr2_b200.ops import linear
    m, n, k = 777, 3072, 768
    x, w = _rand((m, k), dtype, 3), _rand((n, k), dtype, 4, scale=k ** -0.5)
    b = _rand((n,), dtype, 5) if bias else None
    r = _rand((m, n), dtype, 6) if res else None
    got = linear(x, w, bias=b, gelu=gelu, residual=r).float()
    want = _ref_linear(x, w, b, gelu, r)
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.allclose(got, want, rtol=tol, atol=tol), (got - want).abs().max().item()


def test_gemm_exact_on_integer_inputs_and_strided_views():
    from emdr2_b200.ops import linear
    g = torch.Generator(device=DEV).manual_seed(7)
    big = torch.randint(-4, 5, (513, 1024), generator=g, device=DEV).to(torch.float16)
    x = big[:, 128:128 + 320]                       # row-strided view, k = 320
    w = torch.randint(-4, 5, (264, 320), generator=g, device=DEV).to(torch.float16)
    outbuf = torch.zeros((513, 512), dtype=torch.float16, device=DEV)
    out = outbuf[:, 8:8 + 264]
    linear(x, w, out=out)
    want = x.float() @ w.float().T                  # |values| <= 16*320: exact in fp16? not all
    assert torch.equal(out.float(), want.to(torch.float16).float())
    assert (outbuf[:, :8] == 0).all() and (outbuf[:, 8 + 264:] == 0).all()


def test_gemm_rejects_bad_arguments():
    from emdr2_b200 import _lib
    from emdr2_b200.ops import linear
    x = torch.zeros(4, 12, dtype=torch.float16, device=DEV)
    w = torch.zeros(8, 12, dtype=torch.float16, device=DEV)
    with pytest.raises(_lib.Emdr2Error):
        linear(x, w)                                 # row pitch of 12 elements is not 16-byte aligned
    with pytest.raises(TypeError):
        linear(x.float(), w.float())


# ---------------------------------------------------------------- attention
def _ref_attention(q, k, v, batch, heads, sq, sk, q_pad, k_pad, causal, scale=0.125):
    """fp32 restatement of transformer.py:301-383 with the -10000 masked_fill of bert/t5 mask funcs."""
    qf = q.float().view(batch, sq, heads, 64).permute(0, 2, 1, 3)
    kf = k.float().view(batch, sk, heads, 64).permute(0, 2, 1, 3)
    vf = v.float().view(batch, sk, heads, 64).permute(0, 2, 1, 3)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    mask = torch.zeros(batch, 1, sq, sk, dtype=torch.bool, device=q.device)
    if q_pad is not None:
        mask |= q_pad.bool()[:, None, :, None]
    if k_pad is not None:
        mask |= k_pad.bool()[:, None, None, :]
    if causal:
        mask |= torch.ones(sq, sk, dtype=torch.bool, device=q.device).triu(1)[None, None]
    s = s.masked_fill(mask, -10000.0)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf).permute(0, 2, 1, 3).reshape(batch * sq, heads * 64)
    return o, torch.logsumexp(s, dim=-1)


ATTN_CASES = [
    # batch, heads, sq, sk, pad, causal
    (2, 3, 128, 128, False, False),
    (2, 12, 256, 256, True, False),      # BERT tower shape
    (3, 12, 512, 512, True, False),      # T5 encoder shape
    (2, 4, 32, 32, True, True),          # T5 decoder self-attention
    (2, 12, 32, 1536, True, False),      # decoder cross-attention over concatenated passages
    (1, 2, 200, 333, True, False),       # ragged: neither a multiple of 128
    (1, 1, 1, 1, False, False),
    (2, 2, 130, 130, False, True),
    (26, 12, 300, 260, True, False),     # 936 work items: every persistent CTA walks several items
    (40, 8, 130, 130, False, True),      # 640 items, causal, ragged last query block
    (5, 2, 64, 2000, True, False),       # long key axis: many blocks per item (running-max rescales)
]


@pytest.mark.parametrize("batch,heads,sq,sk,pad,causal", ATTN_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attention_forward(batch, heads, sq, sk, pad, causal, dtype):
    from emdr2_b200.ops import attention
    w = heads * 64
    qkv = _rand((batch * max(sq, sk), 3 * w), dtype, 11)
    q = qkv[:batch * sq, :w]                       # strided views, as out of a fused QKV projection
    k = qkv[:batch * sk, w:2 * w]
    v = qkv[:batch * sk, 2 * w:]
    q_pad = k_pad = None
    if pad:
        g = torch.Generator().manual_seed(5)
        qlen = torch.randint(1, sq + 1, (batch,), generator=g)
        klen = torch.randint(1, sk + 1, (batch,), generator=g)
        q_pad = (torch.arange(sq)[None] >= qlen[:, None]).to(DEV)
        k_pad = (torch.arange(sk)[None] >= klen[:, None]).to(DEV)
        if batch > 1:
            q_pad[-1, :] = True if sq > 1 else q_pad[-1, :]     # a fully masked batch entry
    got, lse = attention(q, k, v, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad, causal=causal,
                         return_lse=True)
    want, want_lse = _ref_attention(q, k, v, batch, heads, sq, sk, q_pad, k_pad, causal)
    # probabilities are rounded to 16 bits before P.V and the output once more
    tol = 2 ** -7 if dtype == torch.bfloat16 else 2 ** -10
    err = (got.float() - want).abs().max().item()
    assert torch.allclose(got.float(), want, rtol=tol, atol=tol), err
    assert torch.allclose(lse, want_lse, rtol=1e-4, atol=1e-3), (lse - want_lse).abs().max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attention_running_max_rises_across_blocks(dtype):
    """Keys grow along the key axis, so later blocks beat the running maximum by far more than the
    lazy-rescale slack (2^8) and O must be rescaled in tensor memory several times per row; a second
    half with shrinking keys exercises the no-rescale path with large stale-max probabilities."""
    from emdr2_b200.ops import attention
    batch, heads, sq, sk = 2, 3, 160, 1024
    w = heads * 64
    q = _rand((batch * sq, w), dtype, 61, scale=2.0)
    ramp = torch.cat([torch.linspace(0.2, 6.0, sk // 2), torch.linspace(6.0, 0.2, sk // 2)]).to(DEV)
    k = (_rand((batch * sk, w), dtype, 62).float().view(batch, sk, w) * ramp[None, :, None]).to(dtype).view(batch * sk, w)
    v = _rand((batch * sk, w), dtype, 63)
    got, lse = attention(q, k, v, batch, heads, sq, sk, return_lse=True)
    want, want_lse = _ref_attention(q, k, v, batch, heads, sq, sk, None, None, False)
    tol = 2 ** -6 if dtype == torch.bfloat16 else 2 ** -9
    assert torch.allclose(got.float(), want, rtol=tol, atol=tol), (got.float() - want).abs().max().item()
    assert torch.allclose(lse, want_lse, rtol=1e-4, atol=1e-3), (lse - want_lse).abs().max().item()


# ---------------------------------------------------------------- layernorm / embedding
@pytest.mark.parametrize("rows,h", [(1, 8), (1000, 768), (4097, 1024), (33, 64), (16, 128)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_layernorm_forward(rows, h, dtype):
    from emdr2_b200.ops import layernorm
    x = _rand((rows, h), dtype, 21, scale=3.0) + 1.5
    gamma, beta = _rand((h,), dtype, 22) + 1.0, _rand((h,), dtype, 23)
    got, mean, rstd = layernorm(x, gamma, beta, eps=1e-5, return_stats=True)
    want = torch.nn.functional.layer_norm(x.float(), (h,), gamma.float(), beta.float(), 1e-5)
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    assert torch.allclose(got.float(), want, rtol=tol, atol=tol), (got.float() - want).abs().max().item()
    assert torch.allclose(mean, x.float().mean(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(rstd, (x.float().var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_embedding_forward(dtype):
    from emdr2_b200.ops import embedding
    vocab, h, b, s = 1000, 768, 5, 77
    word, pos, typ = _rand((vocab, h), dtype, 31), _rand((128, h), dtype, 32), _rand((2, h), dtype, 33)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, vocab, (b, s), generator=g).to(DEV)
    types = torch.randint(0, 2, (b, s), generator=g).to(DEV)
    got = embedding(ids, word, pos, types, typ)
    want = (word[ids] + pos[:s][None]) + typ[types]          # 16-bit adds, like the reference module
    assert torch.equal(got.view(b, s, h), want)
    got2 = embedding(ids, word, pos)
    assert torch.equal(got2.view(b, s, h), word[ids] + pos[:s][None])


def test_attention_padding_skip_changes_only_padding_rows():
    """q_live / k_live block maps: identical outputs at non-padding queries, zeros where a whole
    128-query block is padding (FiD-style interleaved padding on the key side)."""
    from emdr2_b200.ops import attention, live_blocks
    dtype = torch.bfloat16
    batch, heads, sq, sk = 3, 4, 384, 1024
    w = heads * 64
    q, k, v = _rand((batch * sq, w), dtype, 41), _rand((batch * sk, w), dtype, 42), _rand((batch * sk, w), dtype, 43)
    q_pad = torch.zeros(batch, sq, dtype=torch.bool, device=DEV)
    q_pad[0, 100:] = True          # blocks 1, 2 dead
    q_pad[1, 300:] = True          # block 2 partially live
    k_pad = torch.zeros(batch, sk, dtype=torch.bool, device=DEV)
    for p0 in range(0, sk, 256):   # 4 "passages" of 256 keys, each padded after 90..150 real tokens
        k_pad[:, p0 + 90 + p0 // 16:p0 + 256] = True
    k_pad[2, :] = True             # a batch entry with no real key at all
    exact = attention(q, k, v, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad)
    fast = attention(q, k, v, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad,
                     q_live=live_blocks(q_pad), k_live=live_blocks(k_pad))
    live_rows = (~q_pad).view(-1)
    live_rows[2 * sq:] = False     # batch 2: every key is padding -> all rows are "uniform" rows
    assert torch.equal(fast[live_rows], exact[live_rows])
    assert (fast.view(batch, sq, w)[0, 128:] == 0).all()
    want, _ = _ref_attention(q, k, v, batch, heads, sq, sk, q_pad, k_pad, False)
    assert torch.allclose(exact.float(), want, rtol=2 ** -7, atol=2 ** -7)


# ---------------------------------------------------------------- general GEMM (training path)
@pytest.mark.parametrize("m,n,k", [(512, 768, 256), (1000, 520, 328), (4096, 3072, 768), (136, 72, 64)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_ex_mn_major_operands(m, n, k, dtype):
    """dX = dY . W (b stored [k_contract, n]) and dW = dY^T . X (both operands stored with the
    contraction index as the row index), vs fp32 matmul."""
    from emdr2_b200.ops import gemm_ex
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    a = _rand((m, k), dtype, 51)
    b_kn = _rand((k, n), dtype, 52, scale=k ** -0.5)
    got = gemm_ex(a, b_kn, b_mn=True).float()
    want = a.float() @ b_kn.float()
    assert torch.allclose(got, want, rtol=tol, atol=tol), (got - want).abs().max().item()
    # a stored [k, m], b stored [k, n]: out[m, n] = a^T . b, accumulated in fp32 with split-K
    a_km = _rand((k, m), dtype, 53, scale=k ** -0.5)
    want2 = a_km.float().T @ b_kn.float()
    for splits in (1, 3):
        acc = torch.ones((m, n), dtype=torch.float32, device=DEV)
        gemm_ex(a_km, b_kn, a_mn=True, b_mn=True, accumulate_into=acc, splits=splits)
        assert torch.allclose(acc - 1.0, want2, rtol=1e-4, atol=1e-4), (acc - 1 - want2).abs().max().item()


def test_gemm_ex_weight_gradient_shape_split_k():
    """dW[n_out, k_in] += dY[tokens, n_out]^T . X[tokens, k_in] with tokens = 20000 split 16 ways."""
    from emdr2_b200.ops import gemm_ex
    dtype = torch.bfloat16
    dy, x = _rand((20000, 768), dtype, 61, scale=0.05), _rand((20000, 3072), dtype, 62)
    acc = torch.zeros((768, 3072), dtype=torch.float32, device=DEV)
    gemm_ex(dy, x, a_mn=True, b_mn=True, accumulate_into=acc, splits=16)
    want = dy.float().T @ x.float()
    assert torch.allclose(acc, want, rtol=2e-4, atol=2e-3), (acc - want).abs().max().item()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_gemm_ex_preactivation_and_gelu_backward(dtype):
    from emdr2_b200.ops import gemm_ex
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    m, n, k = 900, 1024, 256
    x, w, b = _rand((m, k), dtype, 71), _rand((n, k), dtype, 72, scale=k ** -0.5), _rand((n,), dtype, 73)
    pre = torch.empty((m, n), dtype=dtype, device=DEV)
    act = gemm_ex(x, w, bias=b, gelu=True, preact_out=pre)
    u = x.float() @ w.float().T + b.float()
    assert torch.allclose(pre.float(), u, rtol=tol, atol=tol)
    assert torch.allclose(act.float(), torch.nn.functional.gelu(u), rtol=tol, atol=tol)
    # backward through the activation fused into dA = dY . W2: out = (dy . w2) * gelu'(pre)
    dy, w2 = _rand((m, k), dtype, 74), _rand((k, n), dtype, 75, scale=k ** -0.5)
    got = gemm_ex(dy, w2, b_mn=True, gelu_bwd_aux=pre).float()
    uu = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(dy.float() @ w2.float())
    assert torch.allclose(got, uu.grad, rtol=2 * tol, atol=2 * tol), (got - uu.grad).abs().max().item()


# ---------------------------------------------------------------- key-split cross-attention
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_key_split_cross_attention_matches_the_single_pass(dtype):
    """FiD-shaped cross-attention (few questions, very long key axis) runs as several key ranges merged
    by their log-sum-exp weights (autograd.py: _cross_attention_split); one whole range is padding."""
    from emdr2_b200 import autograd as ag
    from emdr2_b200 import ops
    batch, heads, sq, sk = 3, 2, 20, 40 * 128
    w = heads * 64
    q = _rand((batch * sq, w), dtype, 71, scale=1.5)
    kv = _rand((batch * sk, 2 * w), dtype, 72)
    k_pad = torch.zeros(batch, sk, dtype=torch.bool, device=DEV)
    k_pad[0, 3000:] = True                     # question 0: the last ranges are padding only
    k_pad[1, 100:1500] = True                  # question 1: a hole early on
    q_pad = torch.zeros(batch, sq, dtype=torch.bool, device=DEV)
    q_pad[2, 15:] = True
    splits = ag._cross_splits(batch, heads, sk)
    assert splits > 1 and (sk // 128) % splits == 0
    for live in (False, True):
        k_live = ops.live_blocks(k_pad) if live else None
        q_live = ops.live_blocks(q_pad) if live else None
        got = ag.cross_attention(q, kv, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad, q_live=q_live, k_live=k_live)
        want = ops.attention(q, kv[:, :w], kv[:, w:], batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad,
                             q_live=q_live, k_live=k_live)
        ref, _ = _ref_attention(q, kv[:, :w].contiguous(), kv[:, w:].contiguous(), batch, heads, sq, sk, q_pad, k_pad, False)
        keep = ~q_pad.reshape(-1)
        tol = 2 ** -6 if dtype == torch.bfloat16 else 2 ** -9
        assert torch.allclose(got.float()[keep], want.float()[keep], rtol=tol, atol=tol)
        assert torch.allclose(got.float()[keep], ref[keep], rtol=tol, atol=tol)
