"""Pin the CPU oracle against outputs of the reference's own DistributedBruteForceIndex
(tests/golden/mips_ref_*.npz, produced by tests/golden/make_mips_golden.py from
/root/reference/megatron/data/emdr2_index.py:200-305)."""
import numpy as np
import pytest

from oracle import mips as oracle
from helpers import assert_valid_topk, exact_scores, load_golden

CASES = ["c1_exact", "c1_gauss", "shard3"]


@pytest.mark.parametrize("name", CASES)
def test_reference_output_is_a_valid_topk_of_the_oracle_scores(name):
    g = load_golden(name)
    k = int(g["k"])
    full = exact_scores(g["rows"], g["queries"]).astype(np.float16)   # fp16 score storage (:284)
    assert_valid_topk(g["ref_distances"], g["ref_indices"], full, g["ids"], k, what=name)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp16_mode_reproduces_reference_distances(name):
    g = load_golden(name)
    k = int(g["k"])
    s, i = oracle.mips_topk(g["rows"], g["queries"], k, ids=g["ids"], round_fp16=True)
    assert np.array_equal(s.astype(np.float16), g["ref_distances"])
    full = exact_scores(g["rows"], g["queries"]).astype(np.float16)
    assert_valid_topk(s, i, full, g["ids"], k, what=name + "/oracle")
    # ids agree with the reference wherever the fp16 score is unique among the candidates
    for q in range(s.shape[0]):
        col = full[q].astype(np.float32)
        for r in range(k):
            if (col == s[q, r]).sum() == 1:
                assert i[q, r] == g["ref_indices"][q, r]


@pytest.mark.parametrize("name", CASES)
def test_end_to_end_restatement_matches_reference(name):
    """brute_force_index_search restates the whole class incl. the torch.chunk split."""
    g = load_golden(name)
    k, ngpu = int(g["k"]), int(g["ngpu"])
    order = np.argsort(g["ids"])            # restatement wants ascending ids per shard
    if name == "shard3":
        pytest.skip("shuffled ids: covered by the shard-wise test below")
    items = [(int(g["ids"][j]), g["rows"][j]) for j in order]
    dist, idx = oracle.brute_force_index_search(items, g["queries"], k, world=ngpu)
    assert dist.dtype == np.float16 and idx.dtype == np.int32
    assert np.array_equal(dist, g["ref_distances"])


def test_shardwise_oracle_equals_reference_on_shuffled_ids():
    g = load_golden("shard3")
    k, ngpu = int(g["k"]), int(g["ngpu"])
    parts_s, parts_i = [], []
    for lo, hi in oracle.chunk_rows(len(g["ids"]), ngpu):
        s, i = oracle.mips_topk(g["rows"][lo:hi], g["queries"], k, ids=g["ids"][lo:hi],
                                round_fp16=True)
        parts_s.append(s)
        parts_i.append(i)
    s, i = oracle.merge_topk(np.stack(parts_s), np.stack(parts_i))
    assert np.array_equal(s.astype(np.float16), g["ref_distances"])
    full = exact_scores(g["rows"], g["queries"]).astype(np.float16)
    assert_valid_topk(s, i, full, g["ids"], k, what="shard3/merged")


def test_chunk_rows_is_torch_chunk():
    import torch
    for n in [0, 1, 5, 8, 9, 777, 1000, 21_000_000 // 1000]:
        for w in [1, 2, 3, 4, 8]:
            got = [hi - lo for lo, hi in oracle.chunk_rows(n, w)]
            want = [c.shape[0] for c in torch.chunk(torch.empty(n, 1), w, dim=0)] if n else []
            want = want + [0] * (w - len(want))
            assert got == want, (n, w, got, want)


def test_c_oracle_agrees_with_numpy_second_opinion():
    rng = np.random.RandomState(7)
    rows = (rng.randint(-127, 128, size=(3000, 96)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(9, 96)) / 64).astype(np.float16)
    ids = np.arange(10, 3010, dtype=np.int64)
    s1, i1 = oracle.mips_topk(rows, queries, 50, ids=ids)
    s2, i2 = oracle.mips_topk_numpy(rows.astype(np.float32), queries.astype(np.float32), 50, ids=ids)
    assert np.array_equal(s1, s2) and np.array_equal(i1, i2)


def test_oracle_bf16_and_short_shard_padding():
    from oracle.mips import f32_to_bf16_bits, bf16_bits_to_f32
    rng = np.random.RandomState(3)
    rows = f32_to_bf16_bits(rng.randn(3, 16).astype(np.float32))
    queries = f32_to_bf16_bits(rng.randn(2, 16).astype(np.float32))
    s, i = oracle.mips_topk(rows, queries, 5, id_base=1)
    full = exact_scores(bf16_bits_to_f32(rows), bf16_bits_to_f32(queries))
    assert np.array_equal(s[:, :3], -np.sort(-full, axis=1))
    assert np.all(np.isneginf(s[:, 3:])) and np.all(i[:, 3:] == -1)
    s0, i0 = oracle.mips_topk(rows[:0], queries, 5)
    assert np.all(np.isneginf(s0)) and np.all(i0 == -1)


def test_flat_ip_port_agrees_with_the_c_oracle():
    """The CPU-baseline port (FAISS IndexFlatIP algorithm) ranks like the pinned oracle."""
    import torch
    from oracle.flat_ip import flat_ip_search
    rng = np.random.RandomState(9)
    rows = (rng.randint(-127, 128, size=(70000, 64)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(20, 64)) / 64).astype(np.float16)
    ids = np.arange(1, 70001, dtype=np.int64)
    s, i = flat_ip_search(torch.from_numpy(rows).float(), ids, torch.from_numpy(queries).float(), 50)
    ws, wi = oracle.mips_topk(rows, queries, 50, ids=ids)
    assert np.array_equal(s, ws) and np.array_equal(i, wi)
    s, i = flat_ip_search(torch.from_numpy(rows[:3]).float(), None,
                          torch.from_numpy(queries).float(), 5)
    assert np.all(i[:, 3:] == -1) and np.all(np.isneginf(s[:, 3:]))
