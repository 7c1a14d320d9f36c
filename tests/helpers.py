"""Shared checkers for the MIPS parity tests (tie-aware top-k validity, seeded inputs)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, "mips_ref_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


def exact_scores(rows, queries):
    """fp32(fp64-accumulated) score matrix [nq, N] of float-valued inputs (small cases only)."""
    return (np.asarray(queries, np.float64) @ np.asarray(rows, np.float64).T).astype(np.float32)


def assert_valid_topk(scores, ids, full_scores, all_ids, k, what=""):
    """(scores, ids) [nq,k] is *a* correct top-k of full_scores [nq,N] — the tie-order-free
    criterion: the returned score vector equals the k largest values, every returned id carries
    exactly the score reported at its rank, and no id repeats."""
    full_scores = np.asarray(full_scores)
    nq, n = full_scores.shape
    pos = {int(v): i for i, v in enumerate(np.asarray(all_ids))}
    kk = min(k, n)
    want = -np.sort(-full_scores.astype(np.float64), axis=1)[:, :kk]
    got = np.asarray(scores, dtype=np.float64)
    assert np.array_equal(got[:, :kk], want), "%s: score vectors differ from the k largest" % what
    for q in range(nq):
        row_ids = [int(v) for v in np.asarray(ids)[q, :kk]]
        assert len(set(row_ids)) == kk, "%s: duplicate ids for query %d" % (what, q)
        for r, i in enumerate(row_ids):
            assert i in pos, "%s: unknown id %d" % (what, i)
            assert float(full_scores[q, pos[i]]) == got[q, r], \
                "%s: id %d at rank %d of query %d does not carry the reported score" % (what, i, r, q)


def assert_ids_equal_outside_ties(ids, oracle_ids, tie_mask, what=""):
    ids, oracle_ids = np.asarray(ids), np.asarray(oracle_ids)
    free = np.asarray(tie_mask) == 0
    bad = (ids != oracle_ids) & free
    assert not bad.any(), "%s: %d ids differ from the oracle outside tie groups (first at %s)" % (
        what, int(bad.sum()), np.argwhere(bad)[:3].tolist())


def synth(n, d, nq, kind, seed=1234, dtype="float16"):
    """Seeded synthetic evidence/queries: kind 'X' = exact-arithmetic k/64 values,
    'G' = Gaussian.  Returns float32 arrays already rounded to the 16-bit dtype's grid."""
    import torch
    g = torch.Generator().manual_seed(seed)
    tdt = {"float16": torch.float16, "bfloat16": torch.bfloat16}[dtype]
    if kind == "X":
        rows = torch.randint(-127, 128, (n, d), generator=g).float() / 64
        queries = torch.randint(-127, 128, (nq, d), generator=g).float() / 64
    else:
        rows = torch.randn(n, d, generator=g) / d ** 0.5
        queries = torch.randn(nq, d, generator=g)
    return rows.to(tdt), queries.to(tdt)


def to_oracle_input(t):
    """torch fp16/bf16 tensor -> what oracle.mips.mips_topk takes (float16 array / uint16 bf16 bits)."""
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float16:
        return t.numpy()
    return t.view(torch.int16).numpy().view(np.uint16)


# ------------------------------------------------------------------ tiny BERT/T5 config (blocks)
# dropout off unless a test turns it on: the parity tests compare against dropout-free references
TINY = dict(hidden=128, heads=2, layers=2, ffn=256, vocab=128, max_pos=64, hidden_dropout=0.0,
            attention_dropout=0.0)


def seeded_weights(name, shape):
    """Deterministic parameter values keyed by the parameter's name (shared by the script that ran
    the reference and by the tests, so fixtures hold no weights).  CPU generator: identical bits in
    the build container and on the GPU box (same image)."""
    import zlib
    import torch
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    w = torch.randn(shape, generator=g)
    if "layernorm" in name and name.endswith("weight"):
        return 1.0 + 0.1 * w
    if len(shape) >= 2:
        return 0.08 * w
    return 0.05 * w


def tiny_inputs():
    """Seeded token ids with ragged padding (pad id 0) for the tiny BERT tower and T5 reader."""
    rng = np.random.RandomState(4321)
    v = TINY["vocab"]

    def ragged(b, s, lo):
        ids = rng.randint(1, v, size=(b, s)).astype(np.int64)
        lens = rng.randint(lo, s + 1, size=b)
        lens[0] = s
        for i, n in enumerate(lens):
            ids[i, n:] = 0
        return ids

    bert_ids = ragged(6, 48, 5)
    bert_types = (rng.rand(6, 48) < 0.5).astype(np.int64) * (bert_ids > 0)
    b, k, s = 2, 3, 40
    return dict(bert_ids=bert_ids, bert_types=bert_types, t5_enc_ids=ragged(b * k, s, 7),
                t5_dec_ids=ragged(b * k, 12, 2), fid_shape=(b, k, s))
