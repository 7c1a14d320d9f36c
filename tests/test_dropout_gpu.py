"""Dropout on the sm_100a kernels (reference: torch dropout, p = 0.1, at megatron/model/transformer.py:345-346,
397-419, 511-515 and language_model.py:181).  The masks are counter-based and regenerated in the backward
kernels, so the tests (a) pin the kernels' mask to the numpy restatement of csrc/dropout.cuh bit for bit,
(b) REPLAY that mask in a plain fp32 PyTorch reference of each op and compare forward and gradients, and
(c) check the module-level behaviour (train vs eval, determinism under a seed, the reference's switches).

Tolerances: relative Frobenius error 2e-2 (bf16) / 4e-3 (fp16) on attention outputs and gradients, as in
test_backward_gpu.py (16-bit P / dS operands of the tensor-core products)."""
import math

import numpy as np
import pytest
import torch

from helpers import TINY, seeded_weights
from test_dropout import keep_mask

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = {torch.bfloat16: 2e-2, torch.float16: 4e-3}


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _spec(p, cols, seed=77):
    from emdr2_b200 import dropout
    state = dropout.DropoutState(seed)
    return state.next(p, torch.device(DEV), cols)


def test_kernel_mask_is_the_numpy_restatement_bit_for_bit():
    from emdr2_b200 import dropout
    for p, rows, cols, seed in [(0.1, 300, 768, 1234), (0.5, 64, 25600, 5), (0.1, 5000, 64, (1 << 40) + 3)]:
        state = dropout.DropoutState(seed)
        spec = state.next(p, torch.device(DEV), cols)
        spec2 = state.next(p, torch.device(DEV), cols)
        assert (spec.offset, spec2.offset) == (1, 2)
        got = dropout.mask(spec, rows, cols).cpu().numpy().astype(bool)
        assert np.array_equal(got, keep_mask(seed, 1, rows, cols, p))
        assert np.array_equal(dropout.mask(spec2, rows, cols).cpu().numpy().astype(bool), keep_mask(seed, 2, rows, cols, p))
    assert abs(dropout.keep_scale(0.1) - 1 / 0.9) < 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_dropout_add_forward_and_backward_replay_the_same_mask(dtype):
    from emdr2_b200 import autograd as ag, dropout
    rows, cols, p = 777, 768, 0.1
    g = torch.Generator().manual_seed(3)
    y = torch.randn(rows, cols, generator=g).to(dtype).to(DEV).requires_grad_(True)
    res = torch.randn(rows, cols, generator=g).to(dtype).to(DEV).requires_grad_(True)
    spec = _spec(p, cols)
    keep = dropout.mask(spec, rows, cols).float()
    scale = dropout.keep_scale(p)
    out = ag.dropout_add(y, res, p, spec=spec)
    want = (res.float() + y.float() * keep * scale).to(dtype)
    assert torch.equal(out, want)                                    # fp32 arithmetic, one rounding
    gout = torch.randn(rows, cols, generator=g).to(dtype).to(DEV)
    out.backward(gout)
    assert torch.equal(res.grad, gout)
    assert torch.equal(y.grad, (gout.float() * keep * scale).to(dtype))
    # no residual (embedding dropout), in the no-grad path too
    with torch.no_grad():
        alone = ag.dropout_add(y, None, p, spec=spec)
    assert torch.equal(alone, (y.float() * keep * scale).to(dtype))
    assert 0.08 < 1.0 - keep.mean().item() < 0.12


def _reference_attention(q, k, v, heads, mask, keep, scale_keep, scale):
    """fp32 torch: softmax(masked_fill(q k^T * scale, -10000)) -> * keep / (1 - p) -> . v, per head."""
    b, sq, h = q.shape
    sk = k.shape[1]
    qh, kh, vh = (t.view(b, -1, heads, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(mask[:, None], -10000.0)
    pr = torch.softmax(s, dim=-1) * keep.view(b, heads, sq, sk) * scale_keep
    return torch.matmul(pr, vh).permute(0, 2, 1, 3).reshape(b, sq, h)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
def test_self_attention_with_dropout_forward_and_backward_vs_replayed_reference(dtype, causal):
    from emdr2_b200 import autograd as ag, dropout
    b, heads, s, p = 3, 2, 200, 0.1                   # two 128-row blocks per sequence, ragged padding
    h = heads * 64
    g = torch.Generator().manual_seed(11)
    qkv = (torch.randn(b * s, 3 * h, generator=g) * 0.7).to(dtype).to(DEV).requires_grad_(True)
    lens = [200, 131, 57]
    pad = torch.zeros(b, s, dtype=torch.bool)
    for i, n in enumerate(lens):
        pad[i, n:] = True
    spec = _spec(p, s)
    keep = dropout.mask(spec, b * heads * s, s).float()
    out = ag.self_attention(qkv, b, heads, s, pad=pad.to(DEV), causal=causal, scale=0.125, dropout=spec)
    ref_in = qkv.detach().float().requires_grad_(True)
    q, k, v = (ref_in[:, i * h:(i + 1) * h].view(b, s, h) for i in range(3))
    mask = pad[:, :, None] | pad[:, None, :]
    if causal:
        mask = mask | (torch.arange(s)[None, :] > torch.arange(s)[:, None])[None]
    want = _reference_attention(q, k, v, heads, mask.to(DEV), keep, dropout.keep_scale(p), 0.125).view(b * s, h)
    live = (~pad).view(-1).to(DEV)
    assert _rel(out[live], want[live]) < RTOL[dtype]
    gout = (torch.randn(b * s, h, generator=g)).to(dtype).to(DEV) * live[:, None]
    out.backward(gout)
    want.backward(gout.float())
    for i, name in enumerate(("dq", "dk", "dv")):
        got_g, want_g = qkv.grad[:, i * h:(i + 1) * h], ref_in.grad[:, i * h:(i + 1) * h]
        assert _rel(got_g[live], want_g[live]) < 2 * RTOL[dtype], (name, _rel(got_g[live], want_g[live]))
    # dropout really happened, and with the expected strength: the undropped output differs
    plain = ag.self_attention(qkv.detach(), b, heads, s, pad=pad.to(DEV), causal=causal, scale=0.125)
    assert _rel(out[live], plain[live]) > 0.05


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_cross_attention_with_dropout_forward_and_backward_vs_replayed_reference(dtype):
    from emdr2_b200 import autograd as ag, dropout
    b, heads, sq, sk, p = 2, 2, 40, 300, 0.1
    h = heads * 64
    g = torch.Generator().manual_seed(12)
    q = (torch.randn(b * sq, h, generator=g) * 0.7).to(dtype).to(DEV).requires_grad_(True)
    kv = (torch.randn(b * sk, 2 * h, generator=g) * 0.7).to(dtype).to(DEV).requires_grad_(True)
    k_pad = torch.zeros(b, sk, dtype=torch.bool)
    k_pad[0, 250:] = True
    k_pad[1, 129:] = True
    spec = _spec(p, sk)
    keep = dropout.mask(spec, b * heads * sq, sk).float()
    out = ag.cross_attention(q, kv, b, heads, sq, sk, k_pad=k_pad.to(DEV), scale=0.125, dropout=spec)
    rq, rkv = q.detach().float().requires_grad_(True), kv.detach().float().requires_grad_(True)
    mask = k_pad[:, None, :].expand(b, sq, sk)
    want = _reference_attention(rq.view(b, sq, h), rkv[:, :h].reshape(b, sk, h), rkv[:, h:].reshape(b, sk, h), heads,
                                mask.to(DEV), keep, dropout.keep_scale(p), 0.125).view(b * sq, h)
    assert _rel(out, want) < RTOL[dtype]
    gout = torch.randn(b * sq, h, generator=g).to(dtype).to(DEV)
    out.backward(gout)
    want.backward(gout.float())
    live_k = (~k_pad).view(-1).to(DEV)
    assert _rel(q.grad, rq.grad) < 2 * RTOL[dtype]
    assert _rel(kv.grad[live_k], rkv.grad[live_k]) < 2 * RTOL[dtype]


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_key_split_cross_attention_trains_like_the_unsplit_reference(p):
    """A long key axis (>= CROSS_SPLIT_MIN_KEYS) runs as key ranges in the forward AND backward kernels
    (autograd._CrossAttentionFn): outputs merged by lse, dK / dV per range, dQ summed over the ranges.  The mask
    plane of the split launch is (range entry, head, query, key within the range); the fp32 reference replays it."""
    from emdr2_b200 import autograd as ag, dropout
    dtype = torch.bfloat16
    b, heads, sq, sk = 2, 2, 40, 4096
    h = heads * 64
    splits = ag._cross_splits(b, heads, sk)
    assert splits > 1 and sk % (128 * splits) == 0
    sk_s = sk // splits
    g = torch.Generator().manual_seed(13)
    q = (torch.randn(b * sq, h, generator=g) * 0.7).to(dtype).to(DEV).requires_grad_(True)
    kv = (torch.randn(b * sk, 2 * h, generator=g) * 0.7).to(dtype).to(DEV).requires_grad_(True)
    k_pad = torch.zeros(b, sk, dtype=torch.bool)
    k_pad[0, 3000:] = True                    # the last two ranges of question 0 are all padding
    k_pad[1, 129:700] = True
    k_live = ~k_pad.view(b, sk // 128, 128).all(dim=2)
    spec = _spec(p, sk) if p else None
    if p:
        keep = dropout.mask(spec, b * splits * heads * sq, sk_s).float()
        keep = keep.view(b, splits, heads, sq, sk_s).permute(0, 2, 3, 1, 4).reshape(b * heads * sq, sk)
    else:
        keep = torch.ones(b * heads * sq, sk, device=DEV)
    out = ag.cross_attention(q, kv, b, heads, sq, sk, k_pad=k_pad.to(DEV), k_live=k_live.to(DEV), scale=0.125,
                             dropout=spec)
    rq, rkv = q.detach().float().requires_grad_(True), kv.detach().float().requires_grad_(True)
    mask = k_pad[:, None, :].expand(b, sq, sk)
    want = _reference_attention(rq.view(b, sq, h), rkv[:, :h].reshape(b, sk, h), rkv[:, h:].reshape(b, sk, h), heads,
                                mask.to(DEV), keep, dropout.keep_scale(p) if p else 1.0, 0.125).view(b * sq, h)
    assert _rel(out, want) < RTOL[dtype]
    gout = torch.randn(b * sq, h, generator=g).to(dtype).to(DEV)
    out.backward(gout)
    want.backward(gout.float())
    live_k = (~k_pad).view(-1).to(DEV)
    assert _rel(q.grad, rq.grad) < 2 * RTOL[dtype]
    assert _rel(kv.grad[live_k], rkv.grad[live_k]) < 2 * RTOL[dtype]
    # padding keys receive no gradient (their probability is exactly 0)
    assert kv.grad[~live_k].float().abs().max().item() == 0.0
    # the no-grad path takes the same split and, with the same spec, the same mask
    with torch.no_grad():
        again = ag.cross_attention(q.detach(), kv.detach(), b, heads, sq, sk, k_pad=k_pad.to(DEV),
                                   k_live=k_live.to(DEV), scale=0.125, dropout=spec)
    assert torch.equal(again, out.detach())


def test_modules_apply_dropout_in_training_mode_only_and_deterministically():
    """train(): embedding, attention-probability and bias-dropout-add masks are drawn (different outputs from
    eval, identical under the same seed, gradients flow); eval(): exactly the dropout-free path.
    EMDR2Model's --disable-retriever-dropout puts the towers in eval mode (emdr2_model.py:69-77)."""
    from emdr2_b200 import dropout
    from emdr2_b200.blocks import BertTower
    dtype = torch.float16
    cfg = dict(TINY, dtype=dtype, hidden_dropout=0.1, attention_dropout=0.1)
    model = BertTower(cfg).to(DEV)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            prm.copy_(seeded_weights(name, tuple(prm.shape)).to(dtype))
    rng = np.random.RandomState(2)
    ids = torch.from_numpy(rng.randint(1, TINY["vocab"], size=(5, 40)).astype(np.int64))
    ids[1, 30:] = 0
    ids[3, 9:] = 0
    types = torch.zeros_like(ids)
    model.eval()
    with torch.no_grad():
        clean = model(ids.to(DEV), None, types.to(DEV))
    calls_before = dropout.STATE.counter
    with torch.no_grad():
        assert torch.equal(model(ids.to(DEV), None, types.to(DEV)), clean)
    assert dropout.STATE.counter == calls_before               # eval draws no masks
    model.train()
    dropout.manual_seed(42)
    a = model(ids.to(DEV), None, types.to(DEV))
    per_layer = 3                                               # attention probs, attention dense, mlp
    assert dropout.STATE.counter == 1 + per_layer * TINY["layers"]
    dropout.manual_seed(42)
    b = model(ids.to(DEV), None, types.to(DEV))
    dropout.manual_seed(43)
    c = model(ids.to(DEV), None, types.to(DEV))
    assert torch.equal(a, b) and not torch.equal(a, c)
    rel = _rel(a, clean)
    assert 0.02 < rel < 1.0, rel                                # noisy, not destroyed
    a.float().sum().backward()
    grads = [prm.grad for prm in model.parameters() if prm.grad is not None]
    assert len(grads) > 10 and all(torch.isfinite(gr.float()).all() for gr in grads)
    # averaged over many masks the dropped-out forward is centred on the clean one (inverted dropout)
    acc = torch.zeros_like(clean, dtype=torch.float32)
    with torch.no_grad():
        for i in range(48):
            dropout.manual_seed(1000 + i)
            acc += model(ids.to(DEV), None, types.to(DEV)).float()
    assert _rel(acc / 48, clean) < 0.5 * rel
    dropout.manual_seed(1234)
