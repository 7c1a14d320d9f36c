"""Backward passes of the block operators and of the BERT tower / T5 reader, against torch autograd
over fp32 restatements of the same ops (floating point: tolerances stated per test).  Gradients
come back in the 16-bit parameter/activation dtype after fp32 accumulation."""
import math

import numpy as np
import pytest
import torch

from helpers import TINY, seeded_weights, tiny_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(shape, generator=g, device=DEV) * scale).to(dtype)


def _rel(got, want):
    got, want = got.float(), want.float()
    return ((got - want).norm() / (want.norm() + 1e-12)).item()


RTOL = {torch.bfloat16: 1.2e-2, torch.float16: 2e-3}     # relative Frobenius error of a 16-bit gradient


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_linear_and_mlp_backward(dtype):
    from emdr2_b200 import autograd as ag
    m, k, n = 1500, 256, 512
    x = _rand((m, k), dtype, 1).requires_grad_(True)
    w1 = _rand((n, k), dtype, 2, k ** -0.5).requires_grad_(True)
    b1 = _rand((n,), dtype, 3).requires_grad_(True)
    w2 = _rand((k, n), dtype, 4, n ** -0.5).requires_grad_(True)
    b2 = _rand((k,), dtype, 5).requires_grad_(True)
    res = _rand((m, k), dtype, 6).requires_grad_(True)
    gout = _rand((m, k), dtype, 7)
    y = ag.mlp(x, w1, b1, w2, b2, residual=res)
    y.backward(gout)
    got = [t.grad.clone() for t in (x, w1, b1, w2, b2, res)]
    ref = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2, res)]
    yr = ref[5] + torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(ref[0], ref[1], ref[2])),
                                             ref[3], ref[4])
    yr.backward(gout.float())
    assert _rel(y, yr) < RTOL[dtype]
    for g, r, name in zip(got, ref, ["x", "w1", "b1", "w2", "b2", "res"]):
        assert g.dtype == dtype and _rel(g, r.grad) < RTOL[dtype], (name, _rel(g, r.grad))
    # plain linear with residual
    for t in (x, w1, b1):
        t.grad = None
    r2 = _rand((m, n), dtype, 8).requires_grad_(True)
    z = ag.linear(x, w1, b1, residual=r2)
    g2 = _rand((m, n), dtype, 9)
    z.backward(g2)
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w1, b1))
    (torch.nn.functional.linear(xr, wr, br) + r2.detach().float()).backward(g2.float())
    assert _rel(x.grad, xr.grad) < RTOL[dtype] and _rel(w1.grad, wr.grad) < RTOL[dtype]
    assert _rel(b1.grad, br.grad) < RTOL[dtype] and torch.equal(r2.grad, g2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_layernorm_embedding_logprob_backward(dtype):
    from emdr2_b200 import autograd as ag
    rows, h = 777, 768
    x = (_rand((rows, h), dtype, 11, 2.0) + 0.5).requires_grad_(True)
    gam = (_rand((h,), dtype, 12) + 1).requires_grad_(True)
    bet = _rand((h,), dtype, 13).requires_grad_(True)
    g = _rand((rows, h), dtype, 14)
    ag.layernorm(x, gam, bet, 1e-5).backward(g)
    xr, gr, br = (t.detach().float().requires_grad_(True) for t in (x, gam, bet))
    torch.nn.functional.layer_norm(xr, (h,), gr, br, 1e-5).backward(g.float())
    assert _rel(x.grad, xr.grad) < RTOL[dtype] and _rel(gam.grad, gr.grad) < RTOL[dtype]
    assert _rel(bet.grad, br.grad) < RTOL[dtype]

    vocab, b, s = 200, 3, 40
    word, pos, typ = (_rand(sh, dtype, sd).requires_grad_(True) for sh, sd in (((vocab, h), 15), ((64, h), 16), ((2, h), 17)))
    gen = torch.Generator().manual_seed(1)
    ids = torch.randint(0, vocab, (b, s), generator=gen).to(DEV)
    types = torch.randint(0, 2, (b, s), generator=gen).to(DEV)
    ge = _rand((b * s, h), dtype, 18)
    ag.embedding(ids, word, pos, types, typ).backward(ge)
    wr, pr, tr = (t.detach().float().requires_grad_(True) for t in (word, pos, typ))
    (wr[ids] + pr[:s][None] + tr[types]).view(b * s, h).backward(ge.float())
    assert _rel(word.grad, wr.grad) < RTOL[dtype] and _rel(pos.grad, pr.grad) < RTOL[dtype]
    assert _rel(typ.grad, tr.grad) < RTOL[dtype]

    logits = _rand((50, 1024), dtype, 19, 2.0).requires_grad_(True)
    labels = torch.randint(0, 1024, (50,), generator=gen).to(DEV)
    wgt = torch.rand(50, generator=gen).to(DEV)
    (ag.token_logprob(logits, labels) * wgt).sum().backward()
    lr = logits.detach().float().requires_grad_(True)
    (torch.log_softmax(lr, -1).gather(1, labels[:, None]).squeeze(1) * wgt).sum().backward()
    assert _rel(logits.grad, lr.grad) < RTOL[dtype]


def _ref_attn(q, k, v, batch, heads, sq, sk, q_pad, k_pad, causal):
    qf = q.view(batch, sq, heads, 64).permute(0, 2, 1, 3)
    kf = k.view(batch, sk, heads, 64).permute(0, 2, 1, 3)
    vf = v.view(batch, sk, heads, 64).permute(0, 2, 1, 3)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * 0.125
    mask = torch.zeros(batch, 1, sq, sk, dtype=torch.bool, device=q.device)
    if q_pad is not None:
        mask = mask | q_pad[:, None, :, None]
    if k_pad is not None:
        mask = mask | k_pad[:, None, None, :]
    if causal:
        mask = mask | torch.ones(sq, sk, dtype=torch.bool, device=q.device).triu(1)[None, None]
    p = torch.softmax(s.masked_fill(mask, -10000.0), dim=-1)
    return torch.matmul(p, vf).permute(0, 2, 1, 3).reshape(batch * sq, heads * 64)


@pytest.mark.parametrize("batch,heads,sq,sk,causal,cross", [
    (2, 2, 128, 128, False, False), (3, 4, 256, 256, False, False), (2, 3, 200, 200, False, False),
    (2, 4, 32, 32, True, False), (2, 4, 32, 640, False, True), (1, 2, 130, 300, False, True)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_attention_backward(batch, heads, sq, sk, causal, cross, dtype):
    from emdr2_b200 import autograd as ag
    from emdr2_b200.ops import live_blocks
    w = heads * 64
    gen = torch.Generator().manual_seed(3)
    qlen = torch.randint(max(1, sq // 3), sq + 1, (batch,), generator=gen)
    klen = torch.randint(max(1, sk // 3), sk + 1, (batch,), generator=gen)
    q_pad = (torch.arange(sq)[None] >= qlen[:, None]).to(DEV)
    k_pad = q_pad if not cross else (torch.arange(sk)[None] >= klen[:, None]).to(DEV)
    gout = _rand((batch * sq, w), dtype, 24)
    gout = gout * (~q_pad).view(-1, 1)        # no gradient arrives at padding rows (see attention_bwd.cu)
    if cross:
        q = _rand((batch * sq, w), dtype, 21).requires_grad_(True)
        kv = _rand((batch * sk, 2 * w), dtype, 22).requires_grad_(True)
        out = ag.cross_attention(q, kv, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad)
        out.backward(gout)
        qr, kvr = q.detach().float().requires_grad_(True), kv.detach().float().requires_grad_(True)
        ref = _ref_attn(qr, kvr[:, :w], kvr[:, w:], batch, heads, sq, sk, q_pad, k_pad, False)
        ref.backward(gout.float())
        pairs = [(q.grad, qr.grad, "dq"), (kv.grad[:, :w], kvr.grad[:, :w], "dk"), (kv.grad[:, w:], kvr.grad[:, w:], "dv")]
        # skip mode: same gradients
        q2, kv2 = q.detach().clone().requires_grad_(True), kv.detach().clone().requires_grad_(True)
        ag.cross_attention(q2, kv2, batch, heads, sq, sk, q_pad=q_pad, k_pad=k_pad, q_live=live_blocks(q_pad),
                           k_live=live_blocks(k_pad)).backward(gout)
        live_q = (~q_pad).view(-1)
        assert _rel(q2.grad[live_q], q.grad[live_q]) < 1e-6 and _rel(kv2.grad, kv.grad) < 2e-3
    else:
        qkv = _rand((batch * sq, 3 * w), dtype, 23).requires_grad_(True)
        out = ag.self_attention(qkv, batch, heads, sq, pad=q_pad, causal=causal)
        out.backward(gout)
        r = qkv.detach().float().requires_grad_(True)
        ref = _ref_attn(r[:, :w], r[:, w:2 * w], r[:, 2 * w:], batch, heads, sq, sq, q_pad, q_pad, causal)
        ref.backward(gout.float())
        pairs = [(qkv.grad[:, :w], r.grad[:, :w], "dq"), (qkv.grad[:, w:2 * w], r.grad[:, w:2 * w], "dk"),
                 (qkv.grad[:, 2 * w:], r.grad[:, 2 * w:], "dv")]
    assert _rel(out, ref) < RTOL[dtype]
    for g, rg, name in pairs:
        assert _rel(g, rg) < 2 * RTOL[dtype], (name, _rel(g, rg))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_t5_reader_and_bert_tower_parameter_gradients(dtype):
    """d(loss)/d(parameters) through the whole tiny BERT tower and the FiD reader vs fp32 autograd of
    the block oracle on the same 16-bit-rounded weights."""
    from emdr2_b200.blocks import BertTower, T5Reader
    from oracle import blocks as ob
    cfg = dict(TINY, dtype=dtype)
    inp = tiny_inputs()
    tol = 4 * RTOL[dtype]

    def fill(model):
        w32 = {}
        with torch.no_grad():
            for name, p in model.named_parameters():
                w = seeded_weights(name, tuple(p.shape)).to(dtype)
                p.copy_(w)
                w32[name] = w.float().requires_grad_(True)
        return w32

    bert = BertTower(cfg).to(DEV)
    w32 = fill(bert)
    ids, types = torch.from_numpy(inp["bert_ids"]), torch.from_numpy(inp["bert_types"])
    proj = torch.randn(ids.shape[0], TINY["hidden"], generator=torch.Generator().manual_seed(5))
    (bert(ids.to(DEV), None, types.to(DEV)).float() * proj.to(DEV)).sum().backward()
    (ob.bert_pooled(ids, types, w32, TINY["heads"], TINY["layers"]) * proj).sum().backward()
    def worst_of(model, w32):
        errs = []
        for n, p in model.named_parameters():
            if w32[n].grad is None:             # parameter the loss does not depend on (e.g. unused token types)
                assert p.grad is None or float(p.grad.float().abs().max()) == 0.0, n
                continue
            errs.append((_rel(p.grad.cpu(), w32[n].grad), n))
        return max(errs)

    worst = worst_of(bert, w32)
    assert worst[0] < tol, worst

    t5 = T5Reader(cfg).to(DEV)
    w32 = fill(t5)
    enc, dec = torch.from_numpy(inp["t5_enc_ids"]), torch.from_numpy(inp["t5_dec_ids"])
    b, k, s = inp["fid_shape"]
    labels = dec[:b].clone()
    enc_states = t5(enc.to(DEV), dec.to(DEV), output_enc_hidden=True)
    fid_ids = enc.reshape(b, k * s).to(DEV)
    loss, _ = t5(fid_ids[:, :s], dec[:b].to(DEV), enc_hidden_states=enc_states.reshape(b, k * s, -1),
                 enc_ids_for_mask=fid_ids, lm_labels=labels.to(DEV))
    mask = (labels > 0).float()
    (loss * mask.to(DEV)).sum().backward()
    oenc = ob.t5_encode(enc, w32, TINY["heads"], TINY["layers"])
    ologits = ob.t5_decode(dec[:b], oenc.reshape(b, k * s, -1), enc.reshape(b, k * s), w32, TINY["heads"], TINY["layers"])
    oloss = torch.nn.functional.cross_entropy(ologits.view(-1, ologits.shape[-1]), labels.view(-1), reduction="none")
    (oloss * mask.view(-1)).sum().backward()
    worst = worst_of(t5, w32)
    assert worst[0] < tol, worst


def test_main_grad_sinks_equal_the_autograd_gradients():
    """data_parallel.GradientBuckets(main_grad=True): the backward kernels accumulate straight into fp32 flat
    buffers and autograd sees no parameter gradients; the result must equal the ordinary `.grad` path (which casts
    each gradient to 16 bits) to that cast's rounding, for every parameter incl. the fused projections whose
    parameters keep the reference's [np, hn, 3] row order.  Tolerance: relative Frobenius 1e-2 (bf16 cast)."""
    from emdr2_b200.blocks import T5Reader
    from emdr2_b200.data_parallel import GradientBuckets
    dtype = torch.bfloat16
    cfg = dict(TINY, dtype=dtype)
    inp = tiny_inputs()
    t5 = T5Reader(cfg).to(DEV)
    with torch.no_grad():
        for name, p in t5.named_parameters():
            p.copy_(seeded_weights(name, tuple(p.shape)).to(dtype))
    enc, dec = torch.from_numpy(inp["t5_enc_ids"]).to(DEV), torch.from_numpy(inp["t5_dec_ids"]).to(DEV)

    def run():
        loss, _ = t5(enc, dec, lm_labels=dec)
        (loss * (dec > 0).float()).sum().backward()

    run()
    plain = {n: p.grad.float().clone() for n, p in t5.named_parameters() if p.grad is not None}
    for p in t5.parameters():
        p.grad = None
    gb = GradientBuckets(list(t5.parameters()), bucket_bytes=1 << 20, main_grad=True)
    for step in range(2):                                      # sinks are re-zeroed and reused
        gb.start_step()
        run()
        gb.finish()
        assert all(p.grad is None for p in t5.parameters())
        assert all(getattr(p, "_pending_main_grads", 0) == 0 for p in t5.parameters())
        worst = 0.0
        for n, p in t5.named_parameters():
            if n not in plain:
                assert float(p.main_grad.abs().max()) == 0.0, n
                continue
            worst = max(worst, _rel(p.main_grad, plain[n]))
        assert worst < 1e-2, worst
    gb.close()
