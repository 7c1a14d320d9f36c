"""Token-packed (variable-length) forward path: csrc/attention_varlen.cu and emdr2_b200/packed.py against the
rectangular kernels on the same ragged batches.  Packed results at real tokens must be BIT-IDENTICAL to the
rectangular run wherever the same products are summed in the same order (self-attention, the towers); the FiD
cross-attention cuts the key axis differently, so it agrees to fp32 merge rounding (tolerance stated)."""
import numpy as np
import pytest
import torch

from helpers import TINY, seeded_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ragged(b, s, seed, lo=1, vocab=None):
    rng = np.random.RandomState(seed)
    ids = rng.randint(1, vocab or TINY["vocab"], size=(b, s)).astype(np.int64)
    lens = rng.randint(lo, s + 1, size=b)
    lens[rng.randint(0, b)] = s
    lens[rng.randint(0, b)] = lo
    for i, n in enumerate(lens):
        ids[i, n:] = 0
    return torch.from_numpy(ids), lens


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,s,heads", [(7, 300, 2), (33, 129, 3), (5, 64, 1), (3, 513, 2)])
def test_varlen_self_attention_equals_the_rectangular_kernel_at_real_tokens(dtype, b, s, heads):
    from emdr2_b200 import ops
    from emdr2_b200.packed import PackedBatch
    h = heads * 64
    ids, lens = _ragged(b, s, seed=b * 1000 + s)
    g = torch.Generator().manual_seed(b + s)
    qkv = (torch.randn(b * s, 3 * h, generator=g) * 0.8).to(dtype).to(DEV)
    pad = (ids == 0).to(DEV)
    rect = ops.attention(qkv[:, :h], qkv[:, h:2 * h], qkv[:, 2 * h:], b, heads, s, s, q_pad=pad, k_pad=pad, scale=0.125)
    pb = PackedBatch(lens, s, heads, torch.device(DEV))
    assert pb.T == int(lens.sum())
    packed_qkv = qkv.index_select(0, pb.gather)
    out = torch.full((pb.T, h), float("nan"), dtype=dtype, device=DEV)
    lse = torch.empty(heads * pb.T, dtype=torch.float32, device=DEV)
    ops.attention_varlen(packed_qkv[:, :h], packed_qkv[:, h:2 * h], packed_qkv[:, 2 * h:], heads, pb.items, pb.n_items,
                         scale=0.125, out=out)
    assert torch.isfinite(out.float()).all()                     # every real token was written exactly by its tile
    assert torch.equal(out, rect.index_select(0, pb.gather))
    # a second launch into the same buffer writes the same bits (items are independent)
    again = ops.attention_varlen(packed_qkv[:, :h], packed_qkv[:, h:2 * h], packed_qkv[:, 2 * h:], heads, pb.items,
                                 pb.n_items, scale=0.125)
    assert torch.equal(again, out)
    del lse


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_packed_towers_equal_the_rectangular_towers(dtype):
    """BertTower (CLS states) and T5Reader (encoder states + FiD logits) with per-row lengths handed over: the
    packed path (default under no_grad/eval) against the rectangular path of the same modules."""
    from emdr2_b200.blocks import BertTower, T5Reader
    cfg = dict(TINY, dtype=dtype)
    bert = BertTower(cfg).to(DEV).eval()
    t5 = T5Reader(cfg).to(DEV).eval()
    with torch.no_grad():
        for m in (bert, t5):
            for name, p in m.named_parameters():
                p.copy_(seeded_weights(name, tuple(p.shape)).to(dtype))
    ids, lens = _ragged(37, 64, seed=3, lo=2)
    types = ((torch.arange(64)[None, :] % 3 == 0) & (ids > 0)).long()
    with torch.no_grad():
        bert.language_model.packed_varlen = False
        want = bert(ids.to(DEV), None, types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
        bert.language_model.packed_varlen = True
        got = bert(ids.to(DEV), None, types.to(DEV), max_len=int(lens.max()), row_lengths=lens)
    assert torch.equal(got, want)

    b, k, s, L = 3, 4, 64, 12
    enc_ids, elens = _ragged(b * k, s, seed=9, lo=5)
    dec = torch.from_numpy(np.random.RandomState(1).randint(1, TINY["vocab"], size=(b, L)).astype(np.int64))
    dec[1, 7:] = 0
    lm = t5.language_model
    with torch.no_grad():
        lm.packed_varlen = False
        enc_rect = t5(enc_ids.to(DEV), dec.to(DEV), output_enc_hidden=True, enc_max_len=int(elens.max()), enc_row_lengths=elens)
        fid_ids = enc_ids[:, :enc_rect.shape[1]].reshape(b, -1).to(DEV)
        want_logits, _ = t5(fid_ids[:, :1], dec.to(DEV), enc_hidden_states=enc_rect.reshape(b, k * enc_rect.shape[1], -1),
                            enc_ids_for_mask=fid_ids)
        lm.packed_varlen = True
        enc_packed = t5(enc_ids.to(DEV), dec.to(DEV), output_enc_hidden=True, enc_max_len=int(elens.max()), enc_row_lengths=elens)
        assert hasattr(enc_packed, "cross_plan") and enc_packed.states.shape[0] == int(elens.sum())
        live = (enc_ids[:, :enc_rect.shape[1]] > 0).to(DEV)
        assert torch.equal(enc_packed.to_padded(enc_rect.shape[1])[live], enc_rect[live])
        got_logits, _ = t5(enc_ids[:, :1].to(DEV), dec.to(DEV), enc_hidden_states=enc_packed, fid_group=k)
    dlive = (dec > 0).to(DEV)
    # same products; the key axis is merged from differently cut ranges: fp32 merge + one 16-bit rounding of the context
    err = (got_logits.float()[dlive] - want_logits.float()[dlive]).abs().max().item()
    scale = want_logits.float()[dlive].abs().max().item()
    assert err <= (2e-2 if dtype == torch.bfloat16 else 3e-3) * max(1.0, scale), (err, scale)


def test_cross_attention_packed_many_ranges_vs_fp32_reference():
    """8 key sets of 50 short sequences each (~9 000 keys per set, cut into ranges) attended by 32 queries: against a
    plain fp32 torch softmax over each set's keys.  Tolerance: relative Frobenius 2e-2 (bf16 P operand)."""
    from emdr2_b200 import autograd as ag
    from emdr2_b200.packed import PackedStates
    heads, sq, n_sets, group = 12, 32, 8, 50
    h = heads * 64
    rng = np.random.RandomState(4)
    lens = rng.randint(120, 250, size=n_sets * group)
    cu = np.concatenate([[0], np.cumsum(lens)])
    T = int(cu[-1])
    g = torch.Generator().manual_seed(8)
    kv = (torch.randn(T, 2 * h, generator=g) * 0.6).to(torch.bfloat16).to(DEV)
    q = (torch.randn(n_sets * sq, h, generator=g) * 0.6).to(torch.bfloat16).to(DEV)
    states = PackedStates(torch.empty(T, 8, device=DEV), lens, cu)
    plan = states.cross_plan(group, sq, heads)
    assert plan.max_chunks > 1 and plan.n_items >= n_sets * heads * 2
    got = ag.cross_attention_packed(q, kv, heads, plan, scale=0.125)
    for s in range(n_sets):
        k0, k1 = int(cu[s * group]), int(cu[(s + 1) * group])
        qs = q[s * sq:(s + 1) * sq].float().view(sq, heads, 64).permute(1, 0, 2)
        ks = kv[k0:k1, :h].float().view(-1, heads, 64).permute(1, 0, 2)
        vs = kv[k0:k1, h:].float().view(-1, heads, 64).permute(1, 0, 2)
        want = torch.matmul(torch.softmax(torch.matmul(qs, ks.transpose(1, 2)) * 0.125, dim=-1), vs)
        want = want.permute(1, 0, 2).reshape(sq, h)
        rel = ((got[s * sq:(s + 1) * sq].float() - want).norm() / want.norm()).item()
        assert rel < 2e-2, (s, rel)
