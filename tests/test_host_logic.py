"""Host-side logic on CPU: evidence store formats, the torch.chunk split rule, and the
world_size-2 exchange (gloo) with the CUDA searcher replaced by an oracle-backed test double."""
import os
import pickle
import socket
import sys

import numpy as np
import pytest
import torch

from emdr2_b200.index import B200BruteForceIndex, chunk_range
from emdr2_b200.store import EvidenceStore, dict_to_arrays, load_flat, save_flat
from oracle import mips as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_range_is_torch_chunk():
    for n in [0, 1, 7, 8, 9, 1000, 2625]:
        for w in [1, 2, 3, 8]:
            sizes = [c.shape[0] for c in torch.chunk(torch.empty(n, 1), w, dim=0)] if n else []
            sizes += [0] * (w - len(sizes))
            got = [chunk_range(n, w, r) for r in range(w)]
            assert [hi - lo for lo, hi in got] == sizes
            assert got[0][0] == 0 and all(got[r][1] == got[r + 1][0] for r in range(w - 1))
            assert got == oracle.chunk_rows(n, w)


def test_store_pickle_shards_merge_and_flat_roundtrip(tmp_path):
    path = str(tmp_path / "evidence.pkl")
    rng = np.random.RandomState(0)
    rows = rng.randn(10, 8).astype(np.float32)
    for rank, sl in enumerate([slice(0, 4), slice(4, 10)]):
        st = EvidenceStore(path, load_from_path=False, rank=rank)
        st.add_block_data(range(1 + sl.start, 1 + sl.stop), rows[sl])
        with pytest.raises(ValueError):
            st.add_block_data([1 + sl.start], rows[:1])
        st.save_shard()
    st0 = EvidenceStore(path, load_from_path=False, rank=0)
    st0.add_block_data(range(1, 5), rows[:4])
    st0.merge_shards_and_save()
    assert not os.path.exists(st0.temp_dir_name)
    with open(path, "rb") as f:
        state = pickle.load(f)                       # the reference's on-disk format (:33-36)
    assert list(state) == ["embed_data"] and len(state["embed_data"]) == 10
    assert all(v.dtype == np.float16 for v in state["embed_data"].values())
    loaded = EvidenceStore(path)
    ids, arr = loaded.to_arrays()
    assert sorted(ids.tolist()) == list(range(1, 11))
    assert np.array_equal(arr[np.argsort(ids)], rows.astype(np.float16))
    loaded.save_flat()
    fids, frows = load_flat(path)
    assert np.array_equal(fids, ids) and np.array_equal(np.asarray(frows), arr)
    sids, srows = load_flat(path, row_range=(3, 7))
    assert np.array_equal(sids, ids[3:7]) and np.array_equal(np.asarray(srows), arr[3:7])
    loaded.clear()
    assert loaded.embed_data == {}


@pytest.mark.skipif(not os.path.isdir("/root/reference/megatron"), reason="reference not mounted")
def test_store_is_interchangeable_with_the_reference_class(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_mips_golden
    make_mips_golden.install_shims()
    from megatron.data.emdr2_index import OpenRetreivalDataStore
    path = str(tmp_path / "e.pkl")
    ours = EvidenceStore(path, load_from_path=False, rank=0)
    ours.add_block_data([5, 3, 9], np.arange(12, dtype=np.float32).reshape(3, 4))
    ours.save_shard()
    ours.merge_shards_and_save()
    theirs = OpenRetreivalDataStore(path, load_from_path=True, rank=0)
    assert list(theirs.embed_data) == [5, 3, 9]
    for k in ours.embed_data:
        assert np.array_equal(theirs.embed_data[k], ours.embed_data[k])
    theirs.add_block_data([11], np.ones((1, 4), np.float32))
    theirs.save_shard()
    theirs.merge_shards_and_save()
    again = EvidenceStore(path)
    assert list(again.embed_data) == [5, 3, 9, 11]


def test_store_blocks_flat_format_and_overwrite_semantics(tmp_path):
    """The array-backed store: blocks are appended without per-row work, `embed_data` is the reference's
    dict on demand, overwrites follow dict semantics, and format='flat' shards merge into flat files."""
    path = str(tmp_path / "ev.pkl")
    rng = np.random.RandomState(3)
    rows = rng.randn(12, 8).astype(np.float16)
    for rank, sl in enumerate([slice(0, 5), slice(5, 12)]):
        st = EvidenceStore(path, load_from_path=False, rank=rank, format="flat")
        st.add_block_data(np.arange(1 + sl.start, 1 + sl.stop), rows[sl])
        assert len(st) == sl.stop - sl.start and len(st._id_blocks) == 1
        st.save_shard()
    main = EvidenceStore(path, load_from_path=False, rank=0, format="flat")
    main.add_block_data(np.arange(1, 6), rows[:5])
    main.merge_shards_and_save()
    assert not os.path.exists(path) and os.path.exists(str(tmp_path / "ev.rows.f16"))
    whole = EvidenceStore(path)                               # finds the flat files
    ids, arr = whole.to_arrays()
    assert ids.tolist() == list(range(1, 13)) and np.array_equal(np.asarray(arr), rows)
    part = EvidenceStore(path, load_from_path=False)
    part.load_from_file(row_range=(4, 9))
    assert part.loaded_range == (4, 9, 12) and part.to_arrays()[0].tolist() == [5, 6, 7, 8, 9]
    assert list(whole.embed_data) == list(range(1, 13)) and whole.embed_data[7].dtype == np.float16
    with pytest.raises(ValueError):
        whole.add_block_data([3], rows[:1])                   # the reference's overwrite guard (:59-60)
    whole.add_block_data([3, 40], rows[:2] * 0 + 1, allow_overwrite=True)
    assert list(whole.embed_data)[2] == 3 and list(whole.embed_data)[-1] == 40 and len(whole) == 13
    assert np.all(np.asarray(whole.embed_data[3]) == 1)
    overlap = EvidenceStore(path, load_from_path=False, rank=1, format="flat")
    overlap.add_block_data([1], rows[:1])
    overlap.save_shard()
    clash = EvidenceStore(path, load_from_path=False, rank=0, format="flat")
    clash.add_block_data([1, 2], rows[:2])
    clash.save_shard()
    with pytest.raises(AssertionError):
        clash.merge_shards_and_save()                         # :88-90


# ---------------------------------------------------------------- world_size-2 exchange on gloo
class OracleSearcher(object):
    """CPU test double with ShardSearcher's surface, backed by the oracle (tests only)."""

    def __init__(self, d, dtype, device):
        self.d = d

    def set_shard(self, rows, ids, id_base=0):
        self.rows = rows.cpu().numpy()
        self.ids = None if ids is None else ids.cpu().numpy()
        self.id_base = id_base

    def search(self, q, k):
        s, i = oracle.mips_topk(self.rows, q.cpu().numpy(), k, ids=self.ids, id_base=self.id_base)
        return torch.from_numpy(s), torch.from_numpy(i)

    def close(self):
        pass


def oracle_merge(scores, ids):
    s, i = oracle.merge_topk(scores.numpy(), ids.numpy())
    return torch.from_numpy(s), torch.from_numpy(i)


class CpuDoubleIndex(B200BruteForceIndex):
    searcher_factory = OracleSearcher
    merge_fn = staticmethod(oracle_merge)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    from emdr2_b200.retriever import B200EvidenceRetriever
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world)
    rng = np.random.RandomState(11)
    rows = (rng.randint(-127, 128, size=(n, 32)) / 64).astype(np.float16)
    ids = np.arange(1, n + 1, dtype=np.int64)
    local_q = torch.from_numpy((np.random.RandomState(100 + rank).randint(-127, 128, size=(3, 32)) / 64)
                               .astype(np.float16))
    store = type("S", (), {})()
    store.embed_data = {int(i): r for i, r in zip(ids, rows)}
    store.embedding_path = "unused"
    store.clear = lambda: None

    class R(B200EvidenceRetriever):
        index_cls = CpuDoubleIndex

    ret = R(topk=6, embedding_size=32, store=store, allow_trivial_doc=False)
    assert ret.topk == 7
    lo, hi = chunk_range(n, world, rank)
    assert (ret.mips_index.row_lo, ret.mips_index.row_hi) == (lo, hi)
    scores, idx = ret.search_all(local_q)
    topk_data, distance = ret.get_topk(local_q)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), scores=scores.numpy(), ids=idx.numpy(),
             mine=np.array([t[0] for t in topk_data]), dist=distance.numpy(), q=local_q.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 1])
def test_sharded_search_world2_gloo_equals_unsharded_oracle(tmp_path, n):
    import torch.multiprocessing as mp
    world = 2
    for attempt in range(2):          # one retry on a fresh port (bind-and-release can lose the port)
        try:
            mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt:
                raise
    r = [np.load(str(tmp_path / ("r%d.npz" % i))) for i in range(world)]
    assert np.array_equal(r[0]["scores"], r[1]["scores"]) and np.array_equal(r[0]["ids"], r[1]["ids"])
    rng = np.random.RandomState(11)
    rows = (rng.randint(-127, 128, size=(n, 32)) / 64).astype(np.float16)
    allq = np.concatenate([r[0]["q"], r[1]["q"]])
    s, i = oracle.mips_topk(rows, allq, 7, id_base=1)
    assert np.array_equal(r[0]["scores"], s) and np.array_equal(r[0]["ids"], i)
    for rank in range(world):
        assert np.array_equal(r[rank]["mine"], i[3 * rank:3 * rank + 3].astype(np.int32))
        assert r[rank]["dist"].dtype == np.float16


def test_bucket_plan_is_a_partition_sorted_by_length():
    from emdr2_b200.blocks import TransformerLanguageModel
    lm = TransformerLanguageModel.__new__(TransformerLanguageModel)      # plan is pure host arithmetic
    lens = np.random.RandomState(0).randint(1, 200, size=333)
    plan = lm._bucket_plan(333, 256, lens)
    rows = np.concatenate([idx for idx, _ in plan])
    assert sorted(rows.tolist()) == list(range(333))
    assert all(w % 8 == 0 and w <= 256 and lens[idx].max() <= w for idx, w in plan)
    assert [w for _, w in plan] == sorted(w for _, w in plan) and 1 < len(plan) <= lm.length_buckets
    assert lm._bucket_plan(32, 256, lens[:32]) is None                   # small batches run as one rectangle
    assert lm._bucket_plan(333, 256, None) is None
    assert lm._bucket_plan(100, 256, np.full(100, 77)) is None           # equal lengths: nothing to bucket


@pytest.mark.parametrize("clustered", [False, True])
def test_k_above_the_kernel_limit_is_exact_by_range_refinement(clustered):
    """k = 100 (the recall evaluator's setting) on the oracle-backed double: the answer equals the
    oracle's top-100 even when most of it sits in one row range (forces the refinement loop)."""
    rng = np.random.RandomState(5)
    n, d, nq, k = 3000, 16, 5, 100
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(nq, d)) / 64).astype(np.float16)
    if clustered:
        rows[100:400] = (queries[0].astype(np.float32) * 1.5).astype(np.float16) + rows[100:400] / 16
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    index = CpuDoubleIndex(d, device="cpu")
    calls = []
    orig = index._search_range
    index._search_range = lambda q, lo, hi, kk: (calls.append((lo, hi)), orig(q, lo, hi, kk))[1]
    index.add_arrays(ids, rows)
    got_s, got_i = index.search(torch.from_numpy(queries), k)
    want_s, want_i, ties = oracle.mips_topk(rows, queries, k, ids=ids, want_ties=True)
    assert np.array_equal(got_s.numpy(), want_s)
    free = ties == 0
    assert np.array_equal(got_i.numpy()[free], want_i[free])
    for qi in range(nq):                                     # tie groups: same id sets
        assert sorted(got_i.numpy()[qi].tolist()) == sorted(want_i[qi].tolist())
    assert len(calls) >= 4 and (len(calls) > 4) == clustered
    dist16, idx32 = index.search_mips_index(torch.from_numpy(queries), k)
    assert dist16.dtype == torch.float16 and idx32.dtype == torch.int32 and idx32.shape == (nq, k)
    small_s, small_i = index.search(torch.from_numpy(queries), 7)
    assert np.array_equal(small_i.numpy(), got_i.numpy()[:, :7]) or not free[:, :7].all()


def _worker_flat_refresh(rank, world, port, out_dir):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    path = os.path.join(out_dir, "ev.pkl")
    index = CpuDoubleIndex(16, EvidenceStore(path, load_from_path=False), device="cpu", group=dist.group.WORLD)
    loads = []
    orig = EvidenceStore.load_from_file
    EvidenceStore.load_from_file = lambda self, row_range=None: (loads.append(row_range), orig(self, row_range))[1]
    index.update_index()                                   # each rank maps ONLY its torch.chunk range
    rng = np.random.RandomState(77)
    queries = torch.from_numpy((rng.randint(-127, 128, size=(4, 16)) / 64).astype(np.float16))
    s1, i1 = index.search(queries, 7)
    index.reset_index()
    s2, i2 = index.search(queries, 7)
    np.savez(os.path.join(out_dir, "f%d.npz" % rank), s=s1.numpy(), i=i1.numpy(), s2=s2.numpy(), i2=i2.numpy(),
             lo=index.row_lo, hi=index.row_hi, loads=np.array([list(r) for r in loads]))
    dist.barrier()
    dist.destroy_process_group()


def test_index_refresh_from_the_flat_store_loads_only_the_local_row_range(tmp_path):
    """f-2: B200BruteForceIndex.update_index / reset_index over the flat store — every rank memory-maps
    just its own row range (emdr2_index.py:232-266 reloads the whole 32 GB pickle on the index owner)."""
    import torch.multiprocessing as mp
    rng = np.random.RandomState(21)
    n, d = 1001, 16
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    save_flat(str(tmp_path / "ev.pkl"), ids, rows)
    for attempt in range(2):
        try:
            mp.spawn(_worker_flat_refresh, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
            break
        except Exception:
            if attempt:
                raise
    queries = (np.random.RandomState(77).randint(-127, 128, size=(4, d)) / 64).astype(np.float16)
    want_s, want_i, ties = oracle.mips_topk(rows, queries, 7, ids=ids, want_ties=True)
    for rank in range(2):
        r = np.load(str(tmp_path / ("f%d.npz" % rank)))
        assert (int(r["lo"]), int(r["hi"])) == chunk_range(n, 2, rank)
        assert r["loads"].tolist() == [list(chunk_range(n, 2, rank))] * 2
        for s, i in ((r["s"], r["i"]), (r["s2"], r["i2"])):
            assert np.array_equal(s, want_s) and np.array_equal(i[ties == 0], want_i[ties == 0])


def _worker_large_k(rank, world, port, out_dir):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.RandomState(31)
    n, d, k = 1100, 16, 100
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    queries = torch.from_numpy((rng.randint(-127, 128, size=(4, d)) / 64).astype(np.float16))
    index = CpuDoubleIndex(d, device="cpu", group=dist.group.WORLD)
    index.add_arrays(np.arange(1, n + 1, dtype=np.int64), rows)
    scores, ids = index.search(queries, k)
    np.savez(os.path.join(out_dir, "k%d.npz" % rank), scores=scores.numpy(), ids=ids.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_k_above_the_kernel_limit_row_sharded_world2_gloo(tmp_path):
    """k = 100 on two ranks: each rank refines its own row range, one all-gather of [nq, 100] pairs, merge."""
    import torch.multiprocessing as mp
    for attempt in range(2):
        try:
            mp.spawn(_worker_large_k, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
            break
        except Exception:
            if attempt:
                raise
    rng = np.random.RandomState(31)
    rows = (rng.randint(-127, 128, size=(1100, 16)) / 64).astype(np.float16)
    queries = (rng.randint(-127, 128, size=(4, 16)) / 64).astype(np.float16)
    want_s, want_i, ties = oracle.mips_topk(rows, queries, 100, id_base=1, want_ties=True)
    for rank in range(2):
        r = np.load(str(tmp_path / ("k%d.npz" % rank)))
        assert np.array_equal(r["scores"], want_s)
        assert np.array_equal(r["ids"][ties == 0], want_i[ties == 0])


def test_split_k_factor_fills_whole_waves_of_the_persistent_grid():
    """autograd._splits: the weight-gradient product hands (tile, split) items round-robin to `ctas` persistent CTAs;
    the factor must fill whole waves (items / (waves * ctas) high) without buying balance with many splits (each split
    is a pass of fp32 atomics over the output tile), and keep slices of >= 4096 tokens."""
    from emdr2_b200.autograd import _splits

    def occupancy(tiles, s, ctas):
        items = tiles * s
        return items / float(-(-items // ctas) * ctas)

    # the step's tile counts on 148 SMs, and on the 144 a trainer leaves when NCCL holds 4
    for ctas in (148, 144):
        assert _splits(204800, 18, ctas) == 8 and _splits(204800, 54, ctas) == 8 and _splits(204800, 72, ctas) == 2
    for tiles in (9, 18, 36, 54, 72, 720):
        for ctas in (148, 144, 140, 132):
            s = _splits(204800, tiles, ctas)
            assert 1 <= s <= 32
            assert occupancy(tiles, s, ctas) >= 0.85, (tiles, ctas, s)
            assert occupancy(tiles, s, ctas) + 1e-9 >= occupancy(tiles, 1, ctas) - 0.015 * s   # never worse than no split
    assert _splits(204800, 72, 140) <= 8          # not the 15 splits a balance-only rule picks
    assert _splits(2048, 18) == 1 and _splits(12800, 18) <= 4      # slices of at least 4096 tokens


def test_key_split_merge_reproduces_the_unsplit_softmax():
    """autograd._merge_splits / _split_masks / _rep_rows (pure torch, runs on the CPU): attention over a long key axis
    computed range by range — each range with its own softmax normalisation and log-sum-exp, as the kernels return
    them — and merged by the lse weights equals attention over the whole axis; the merged lse is the global one.  One
    range is all padding (weight exp(-10000 - lse) = 0)."""
    from emdr2_b200 import autograd as ag
    torch.manual_seed(3)
    b, heads, sq, sk, splits = 2, 3, 5, 1024, 4
    h, sk_s = heads * 64, sk // splits
    q = torch.randn(b * sq, h)
    k, v = torch.randn(b * sk, h), torch.randn(b * sk, h)
    k_pad = torch.zeros(b, sk, dtype=torch.uint8)
    k_pad[0, 700:] = 1                           # question 0: the last range is all padding
    k_pad[1, 100:300] = 1
    k_live = (k_pad.view(b, sk // 128, 128) == 0).any(dim=2).to(torch.uint8)

    def attend(qx, kx, vx, nb, nk, pad):        # fp32 reference of one launch: masked_fill(-10000) semantics
        qh = qx.view(nb, sq, heads, 64).permute(0, 2, 1, 3)
        kh = kx.view(nb, nk, heads, 64).permute(0, 2, 1, 3)
        vh = vx.view(nb, nk, heads, 64).permute(0, 2, 1, 3)
        s = torch.matmul(qh, kh.transpose(-1, -2)) * 0.125
        s = s.masked_fill(pad.bool()[:, None, None, :], -10000.0)
        lse = torch.logsumexp(s, dim=-1)                                      # [nb, heads, sq]
        out = torch.matmul(torch.softmax(s, dim=-1), vh).permute(0, 2, 1, 3).reshape(nb * sq, h)
        return out, lse

    want, want_lse = attend(q, k, v, b, sk, k_pad)
    q_pad_s, k_pad_s, q_live_s, k_live_s = ag._split_masks(None, k_pad, None, k_live, b, sk, splits)
    assert q_pad_s is None and q_live_s is None
    assert k_pad_s.shape == (b * splits, sk_s) and k_live_s.shape == (b * splits, sk_s // 128)
    assert bool((k_live_s[:, 0] == 1).all())                                  # one live block per entry, always
    q_rep = ag._rep_rows(q, b, splits, sq)
    assert q_rep.shape == (b * splits * sq, h) and torch.equal(q_rep[sq:2 * sq], q[:sq])
    out_s, lse_s = attend(q_rep, k, v, b * splits, sk_s, k_pad_s)
    got, got_lse = ag._merge_splits(out_s, lse_s, b, splits, heads, sq, torch.float32)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(got_lse, want_lse, rtol=1e-6, atol=1e-5)
    # the number of ranges is a divisor of the 128-key block count, >= 4 blocks per range, 1 below the threshold
    assert ag._cross_splits(8, 12, 25600) in (8, 10) and 200 % ag._cross_splits(8, 12, 25600) == 0
    assert ag._cross_splits(8, 12, 2048) == 1 and ag._cross_splits(400, 12, 25600) == 1
