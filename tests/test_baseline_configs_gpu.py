"""Oracle parity at BASELINE.json's own sizes ("parity gaps"): the per-GPU shard of
config 3 (2 625 000 x 768 fp16) and a shard whose row numbers pass 2^24 against the C oracle on a
query subset, k = 100 through the row-range refinement on the GPU (the recall evaluator's setting),
and the T5-base-shaped reader (hidden 768, 12 layers, S = 512, FiD over 4096 keys with the key-split
cross-attention and the length-bucketed encoder on, V = 30 720) against oracle/blocks.py.

Tolerances are stated next to each assertion; ids are bit-exact outside the oracle's numerical-tie
groups (neighbouring scores within 2^-20 relative — fp32 tensor-core accumulation order vs the
oracle's fp64 sum)."""
import numpy as np
import pytest
import torch

from helpers import assert_ids_equal_outside_ties, seeded_weights, to_oracle_input

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _oracle():
    from oracle import mips
    return mips


def _gauss_shard(n, d, dtype, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    rows = torch.empty((n, d), dtype=dtype, device=DEV)
    for r0 in range(0, n, 1 << 20):
        r1 = min(n, r0 + (1 << 20))
        rows[r0:r1] = (torch.randn(r1 - r0, d, generator=g, device=DEV) / d ** 0.5).to(dtype)
    return rows


def _check_against_oracle(rows, queries, k, subset, id_base=1):
    from emdr2_b200.mips import ShardSearcher
    s = ShardSearcher(rows.shape[1], rows.dtype, DEV)
    s.set_shard(rows, None, id_base=id_base)
    got_s, got_i = s.search(queries, k)
    got_s, got_i = got_s.cpu().numpy(), got_i.cpu().numpy()
    s.close()
    host_rows = to_oracle_input(rows)
    want_s, want_i, ties = _oracle().mips_topk(host_rows, to_oracle_input(queries[subset]), k, id_base=id_base,
                                               want_ties=True)
    sub = np.asarray(subset)
    assert_ids_equal_outside_ties(got_i[sub], want_i, ties, "shard %d x %d" % tuple(rows.shape))
    assert np.allclose(got_s[sub], want_s, rtol=1e-5, atol=1e-6)
    for j in range(len(sub)):                 # inside a tie group the same ids come back, in some order
        assert sorted(got_i[sub[j]].tolist()) == sorted(want_i[j].tolist()) or ties[j].any()
    return got_s, got_i


def test_config3_per_gpu_shard_2625000_rows_vs_oracle():
    """BASELINE configs[2] at W = 8: one rank's 2 625 000 x 768 fp16 shard, 64 queries, top-50; the C
    oracle (scalar fp64, ~10 s on 8 host threads) checks 8 of the 64 queries."""
    n, d, nq, k = 2_625_000, 768, 64, 50
    rows = _gauss_shard(n, d, torch.float16, 1234)
    g = torch.Generator(device=DEV).manual_seed(99)
    queries = torch.randn(nq, d, generator=g, device=DEV).to(torch.float16)
    _check_against_oracle(rows, queries, k, subset=list(range(0, 64, 8)), id_base=1)


def test_shard_with_more_than_2_pow_24_rows_vs_oracle():
    """17 000 000 rows (> 2^24 = 16 777 216: row numbers no longer fit an fp32 mantissa or a 24-bit
    field) at d = 64 so the host copy the oracle reads stays at 2.2 GB.  Rows planted beyond 2^24 must
    come back with their exact row number + id base."""
    n, d, nq, k = 17_000_000, 64, 16, 50
    rows = _gauss_shard(n, d, torch.float16, 7)
    g = torch.Generator(device=DEV).manual_seed(5)
    queries = torch.randn(nq, d, generator=g, device=DEV).to(torch.float16)
    planted = (1 << 24) + 1 + torch.arange(nq, device=DEV) * 13001          # all > 2^24, odd offsets
    rows[planted] = (queries.float() * 0.5).to(torch.float16)
    got_s, got_i = _check_against_oracle(rows, queries, k, subset=[0, 5, 10, 15], id_base=1)
    assert np.array_equal(got_i[:, 0], planted.cpu().numpy() + 1)


@pytest.mark.parametrize("clustered", [False, True])
def test_k100_row_range_refinement_on_the_gpu_vs_oracle(clustered):
    """k = 100 > the kernel's 64 (examples/helper-scripts/create_wiki_indexes_and_evaluate.sh:67) through
    B200BruteForceIndex._search_local_large on the real scan kernel; the clustered case puts most of the
    answer into one row range and forces the halving loop."""
    from emdr2_b200.index import B200BruteForceIndex, B200FaissMIPSIndex
    rng = np.random.RandomState(5)
    n, d, nq, k = 60000, 128, 9, 100
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(nq, d)) / 64).astype(np.float16)
    if clustered:
        # (k/64) * (3k'/128 + k''/256): every product is an integer / 2^14 and the 128-term sums stay below 2^24 of
        # those units, so fp32 accumulation is exact in any order
        rows[1000:1400] = (queries[0].astype(np.float32) * 1.5).astype(np.float16) + rows[1000:1400] / 4
    ids = np.arange(1, n + 1, dtype=np.int64)          # ascending ids: row order == id order
    index = B200BruteForceIndex(d, device=DEV)
    calls = []
    orig = index._search_range
    index._search_range = lambda q, lo, hi, kk: (calls.append((lo, hi)), orig(q, lo, hi, kk))[1]
    index.add_arrays(ids, rows)
    got_s, got_i = index.search(torch.from_numpy(queries).to(DEV), k)
    want_s, want_i, ties = _oracle().mips_topk(rows, queries, k, ids=ids, want_ties=True)
    assert np.array_equal(got_s.cpu().numpy(), want_s)                 # exact-arithmetic inputs: bit-exact
    assert_ids_equal_outside_ties(got_i.cpu().numpy(), want_i, ties)
    assert len(calls) >= 4 and (len(calls) > 4) == clustered
    d16, i32 = index.search_mips_index(torch.from_numpy(queries).to(DEV), k)
    assert d16.dtype == torch.float16 and i32.dtype == torch.int32 and tuple(i32.shape) == (nq, k)
    fa = B200FaissMIPSIndex(d, device=DEV)
    fa.add_arrays(ids, rows)
    fd, fi = fa.search_mips_index(torch.from_numpy(queries), k, reconstruct=False)
    assert fd.dtype == np.float32 and fi.dtype == np.int64 and np.array_equal(fd, want_s)


def test_t5_base_shape_reader_vs_oracle():
    """T5-base-shaped reader at the recipe's shapes: 2 questions x 8 passages, S = 512 (ragged, NQ-like
    lengths), 12 layers / 12 heads / hidden 768 / ffn 3072, V = 30 720, decoder L = 32, FiD cross-attention
    over 8 x 512 = 4096 keys per question (>= CROSS_SPLIT_MIN_KEYS: the key-split path runs) with the
    length-bucketed, token-packed encoder on.  Reference: oracle/blocks.py in fp32 on the same
    bf16-rounded weights.  Tolerance (bf16 activations through 12 + 12 layers): encoder states at real
    tokens max|err| <= 0.12 (3 ulp at |x| in [4,8)), relative Frobenius error <= 1.6e-2; logits relative
    Frobenius error <= 2.5e-2 and the greedy token agrees wherever the oracle's top-2 logit margin
    exceeds 0.05."""
    from emdr2_b200 import autograd as ag
    from emdr2_b200.blocks import T5Reader
    from oracle import blocks as ob
    dtype = torch.bfloat16
    cfg = dict(hidden=768, heads=12, layers=12, ffn=3072, vocab=30720, max_pos=512, dtype=dtype, hidden_dropout=0.0,
               attention_dropout=0.0)
    model = T5Reader(cfg).to(DEV)
    w32 = {}
    with torch.no_grad():
        for name, p in model.named_parameters():
            w = seeded_weights(name, tuple(p.shape))
            if p.dim() >= 2:
                w = w * 0.25                   # N(0, 0.02): the recipe's init scale (model/utils.py:25-40)
            w = w.to(dtype)
            p.copy_(w)
            w32[name] = w.float()
    b, kp, s, L = 2, 8, 512, 32
    rng = np.random.RandomState(11)
    enc = rng.randint(1000, 30000, size=(b * kp, s)).astype(np.int64)
    lens = rng.randint(150, 420, size=b * kp)
    lens[3] = s
    for i, n in enumerate(lens):
        enc[i, n:] = 0
    dec = np.zeros((b, L), dtype=np.int64)
    for i in range(b):
        m = int(rng.randint(3, 9))
        dec[i, 0] = 30522
        dec[i, 1:1 + m] = rng.randint(1000, 30000, size=m)
    enc_t, dec_t = torch.from_numpy(enc), torch.from_numpy(dec)

    lm = model.language_model
    lm.bucket_min_rows, lm.length_buckets = 8, 4                 # 16 rows: make the bucketed path run
    assert lm._bucket_plan(b * kp, s, lens) is not None
    assert ag._cross_splits(b, 12, kp * s) > 1                   # the key-split path is taken
    with torch.no_grad():
        enc_out = model(enc_t.to(DEV), dec_t.to(DEV), output_enc_hidden=True, enc_max_len=int(lens.max()),
                        enc_row_lengths=lens)
        assert enc_out.shape == (b * kp, s, 768)
        fid_states = enc_out.reshape(b, kp * s, 768)
        fid_ids = enc_t.reshape(b, kp * s).to(DEV)
        logits, _ = model(fid_ids[:, :1], dec_t.to(DEV), enc_hidden_states=fid_states, enc_ids_for_mask=fid_ids)
    assert logits.shape == (b, L, 30720)

    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        want_enc = ob.t5_encode(enc_t, w32, 12, 12)
        want_logits = ob.t5_decode(dec_t, want_enc.reshape(b, kp * s, 768), enc_t.reshape(b, kp * s), w32, 12, 12)
    live = enc_t > 0
    ge, we = enc_out.float().cpu()[live], want_enc[live]
    assert (ge - we).abs().max().item() <= 0.12, (ge - we).abs().max().item()
    assert ((ge - we).norm() / we.norm()).item() <= 1.6e-2
    dlive = dec_t > 0
    gl, wl = logits.float().cpu()[dlive], want_logits[dlive]
    assert ((gl - wl).norm() / wl.norm()).item() <= 2.5e-2, ((gl - wl).norm() / wl.norm()).item()
    top2 = wl.topk(2, dim=-1).values
    sure = (top2[:, 0] - top2[:, 1]) > 0.05
    assert sure.any() and torch.equal(gl.argmax(-1)[sure], wl.argmax(-1)[sure])
