"""BERT tower / T5 reader forward on the sm_100a block kernels against (a) the committed outputs
of the reference's own modules (tests/golden/blocks_ref_*.npz) and (b) the CPU block oracle.

Tolerances (stated, floating point): the CUDA path keeps weights and activations in 16 bits with
fp32 accumulation; the reference outputs are fp32.  Against the oracle evaluated on the SAME
16-bit-rounded weights the only difference is activation rounding (every intermediate of a layer
is stored in 16 bits): |err| <= 4e-2 for bf16 (2.5 ulp at |x| in [2,4), the range of post-LayerNorm
states) / 5e-3 for fp16, and a relative Frobenius error <= 8e-3 / 1e-3.  Against the reference's
own fp32-weight outputs the bound is 4x that (weight rounding adds in)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, TINY, seeded_weights, tiny_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {torch.bfloat16: (4e-2, 8e-3), torch.float16: (5e-3, 1e-3)}


def _cfg(dtype, **over):
    cfg = dict(TINY, dtype=dtype)
    cfg.update(over)
    return cfg


def _fill(module, dtype):
    w32 = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            w = seeded_weights(name, tuple(p.shape)).to(dtype)
            p.copy_(w)
            w32[name] = w.float()
    return w32


def _close(got, want, dtype, scale=1.0):
    abs_tol, rel_tol = TOL[dtype]
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs().max().item()
    rel = ((got - want).norm() / want.norm()).item()
    assert err <= abs_tol * scale and rel <= rel_tol * scale, (err, rel)


def _golden(name):
    with np.load(os.path.join(GOLDEN, "blocks_ref_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_bert_tower_tiny_vs_reference_and_oracle(dtype):
    from emdr2_b200.blocks import BertTower
    from oracle import blocks as ob
    model = BertTower(_cfg(dtype)).to(DEV)
    model.language_model.skip_padding = False        # reference values at padding positions too
    w32 = _fill(model, dtype)
    assert sorted(w32) == sorted(str(n) for n in _golden("bert")["names"])      # same parameter names
    inp = tiny_inputs()
    ids, types = torch.from_numpy(inp["bert_ids"]), torch.from_numpy(inp["bert_types"])
    hidden = model.hidden_states(ids.to(DEV), types.to(DEV))
    pooled = model(ids.to(DEV), None, types.to(DEV))
    want = ob.bert_hidden(ids, types, w32, TINY["heads"], TINY["layers"])
    _close(hidden, want, dtype)
    _close(pooled, want[:, 0], dtype)
    # the reference's fp32 outputs (fp32 weights there): add the weight-rounding error
    g = _golden("bert")
    _close(hidden, torch.from_numpy(g["hidden"]), dtype, scale=4.0)
    _close(pooled, torch.from_numpy(g["pooled"]), dtype, scale=4.0)
    # padding skip (the default): every non-padding position is unchanged
    model.language_model.skip_padding = True
    fast = model.hidden_states(ids.to(DEV), types.to(DEV))
    live = (ids > 0)
    assert torch.equal(fast.cpu()[live], hidden.cpu()[live])
    assert torch.equal(model(ids.to(DEV), None, types.to(DEV)), pooled)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_t5_reader_tiny_vs_reference_and_oracle(dtype):
    from emdr2_b200.blocks import T5Reader
    from oracle import blocks as ob
    model = T5Reader(_cfg(dtype)).to(DEV)
    model.language_model.skip_padding = False
    w32 = _fill(model, dtype)
    assert sorted(w32) == sorted(str(n) for n in _golden("t5")["names"])
    inp = tiny_inputs()
    enc, dec = torch.from_numpy(inp["t5_enc_ids"]), torch.from_numpy(inp["t5_dec_ids"])
    logits, enc_out = model(enc.to(DEV), dec.to(DEV))
    want_logits, want_enc = ob.t5_forward(enc, dec, w32, TINY["heads"], TINY["layers"])
    _close(enc_out, want_enc, dtype)
    _close(logits, want_logits, dtype)
    g = _golden("t5")
    _close(enc_out, torch.from_numpy(g["enc_out"]), dtype, scale=4.0)
    _close(logits, torch.from_numpy(g["logits"]), dtype, scale=4.0)
    # FiD: K passages' encoder states concatenated along the key axis, encoder bypassed
    b, k, s = inp["fid_shape"]
    only_enc = model(enc.to(DEV), dec.to(DEV), output_enc_hidden=True)
    assert torch.equal(only_enc, enc_out)
    fid_states = enc_out.reshape(b, k * s, -1)
    fid_ids = enc.reshape(b, k * s).to(DEV)
    fid_logits, _ = model(fid_ids[:, :s], dec[:b].to(DEV), enc_hidden_states=fid_states,
                          enc_ids_for_mask=fid_ids)
    want_fid = ob.t5_decode(dec[:b], want_enc.reshape(b, k * s, -1), enc.reshape(b, k * s), w32,
                            TINY["heads"], TINY["layers"])
    _close(fid_logits, want_fid, dtype)
    _close(fid_logits, torch.from_numpy(g["fid_logits"]), dtype, scale=4.0)
    # padding skip (the default): logits at non-padding decoder positions and encoder states at
    # non-padding positions are bit-identical
    model.language_model.skip_padding = True
    f_logits, f_enc = model(enc.to(DEV), dec.to(DEV))
    assert torch.equal(f_enc.cpu()[enc > 0], enc_out.cpu()[enc > 0])
    assert torch.equal(f_logits.cpu()[dec > 0], logits.cpu()[dec > 0])
    f_fid, _ = model(fid_ids[:, :s], dec[:b].to(DEV), enc_hidden_states=f_enc.reshape(b, k * s, -1),
                     enc_ids_for_mask=fid_ids)
    assert torch.equal(f_fid.cpu()[dec[:b] > 0], fid_logits.cpu()[dec[:b] > 0])
    model.language_model.skip_padding = False
    loss, _ = model(enc.to(DEV), dec.to(DEV), lm_labels=dec.to(DEV))
    want_loss = torch.nn.functional.cross_entropy(want_logits.reshape(-1, want_logits.shape[-1]),
                                                  dec.reshape(-1), reduction="none").view_as(dec)
    assert torch.allclose(loss.cpu(), want_loss, rtol=5e-2, atol=5e-2)


def test_bert_base_shape_vs_oracle():
    """Full-width stack: hidden 768, 12 heads, 12 layers, ffn 3072, 4 x 256 tokens."""
    from emdr2_b200.blocks import BertTower
    from oracle import blocks as ob
    dtype = torch.bfloat16
    cfg = dict(hidden=768, heads=12, layers=12, ffn=3072, vocab=1024, max_pos=256, dtype=dtype, hidden_dropout=0.0,
               attention_dropout=0.0)
    model = BertTower(cfg).to(DEV)
    model.language_model.skip_padding = False
    w32 = _fill(model, dtype)
    rng = np.random.RandomState(0)
    ids = torch.from_numpy(rng.randint(1, 1024, size=(4, 256)).astype(np.int64))
    for i, n in enumerate([256, 180, 101, 17]):
        ids[i, n:] = 0
    types = torch.zeros_like(ids)
    got = model.hidden_states(ids.to(DEV), types.to(DEV))
    want = ob.bert_hidden(ids, types, w32, 12, 12)
    live = (ids > 0)
    # positions that hold real tokens (padding rows are defined too, but nobody consumes them)
    _close(got.cpu()[live], want[live], dtype, scale=2.0)
    _close(got.cpu(), want, dtype, scale=2.0)


def test_reference_checkpoint_layout_loads():
    from emdr2_b200.blocks import BertTower, load_reference_state_dict
    model = BertTower(_cfg(torch.float16)).to(DEV)
    flat = {n: seeded_weights(n, tuple(p.shape)) for n, p in model.named_parameters()}
    nested = {}
    for k, v in flat.items():
        node = nested
        parts = k.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = v
    load_reference_state_dict(model, nested)
    for n, p in model.named_parameters():
        assert torch.equal(p.detach().cpu(), flat[n].to(torch.float16))
    with pytest.raises(KeyError):
        load_reference_state_dict(model, {"bogus.weight": torch.zeros(1)})


# ------------------------------------------------------------------ length-bucketed execution
def _ragged_batch(b, s, vocab, seed, lo=3):
    rng = np.random.RandomState(seed)
    ids = rng.randint(1, vocab, size=(b, s)).astype(np.int64)
    lens = rng.randint(lo, s + 1, size=b)
    lens[rng.randint(0, b)] = s
    for i, n in enumerate(lens):
        ids[i, n:] = 0
    types = (rng.rand(b, s) < 0.5).astype(np.int64) * (ids > 0)
    return torch.from_numpy(ids), torch.from_numpy(types), lens


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_length_bucketed_encoder_equals_the_rectangular_run_at_every_token(dtype):
    """Sorting rows by length into token-packed buckets (blocks.py: encode) leaves every non-padding
    position bit-identical (cut columns are padding in every member of the bucket; GEMM and LayerNorm
    rows are independent); columns beyond a bucket's width come back as zeros."""
    from emdr2_b200.blocks import BertTower
    model = BertTower(_cfg(dtype)).to(DEV)
    _fill(model, dtype)
    lm = model.language_model
    ids, types, lens = _ragged_batch(70, 64, TINY["vocab"], 5)
    lm.bucket_min_rows = 1 << 30
    plain = lm.encode(ids.to(DEV), types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
    lm.bucket_min_rows, lm.length_buckets = 8, 3
    assert lm._bucket_plan(70, 64, lens) is not None
    with torch.no_grad():
        fast = lm.encode(ids.to(DEV), types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
        cls = model(ids.to(DEV), None, types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
    live = ids > 0
    assert fast.shape == plain.shape
    assert torch.equal(fast.cpu()[live], plain.detach().cpu()[live])
    # columns beyond a row's bucket width were never computed: they come back as zeros
    plan = lm._bucket_plan(70, 64, lens)
    for idx, w in plan:
        assert (fast.cpu()[torch.from_numpy(idx), w:] == 0).all()
    assert torch.equal(cls, plain.detach()[:, 0, :])


def test_length_bucketed_encoder_backward_matches_the_rectangular_run():
    from emdr2_b200.blocks import BertTower
    dtype = torch.float16
    model = BertTower(_cfg(dtype)).to(DEV)
    _fill(model, dtype)
    lm = model.language_model
    ids, types, lens = _ragged_batch(40, 48, TINY["vocab"], 9)
    weight = torch.randn(40, TINY["hidden"], generator=torch.Generator().manual_seed(1)).to(DEV)

    def grads(bucketed):
        lm.bucket_min_rows, lm.length_buckets = (8, 4) if bucketed else (1 << 30, 4)
        for p in model.parameters():
            p.grad = None
        out = model(ids.to(DEV), None, types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
        (out.float() * weight).sum().backward()
        return out.detach(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}

    out_a, g_a = grads(False)
    out_b, g_b = grads(True)
    assert torch.equal(out_a, out_b)
    assert sorted(g_a) == sorted(g_b) and len(g_a) > 10
    for n in g_a:        # same products, different summation order over the (re-ordered) tokens
        denom = g_a[n].norm().item() + 1e-6
        assert (g_a[n] - g_b[n]).norm().item() / denom < 5e-3, n
