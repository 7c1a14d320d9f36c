"""Parity of the sm_100a MIPS path (through the C ABI) against the CPU oracle and the committed
outputs of the reference's own DistributedBruteForceIndex.  Bit-exact ids and scores on
exact-arithmetic inputs; on Gaussian inputs ids must match wherever the oracle's neighbouring
scores are further apart than 2^-20*|score| (the oracle's tie_mask) and scores within 1e-5 rel."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import (assert_ids_equal_outside_ties, assert_valid_topk, exact_scores, load_golden,
                     synth, to_oracle_input)

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _oracle():
    from oracle import mips
    return mips


def _searcher(d, dtype):
    from emdr2_b200.mips import ShardSearcher
    return ShardSearcher(d, dtype, DEV)


def run_case(n, d, nq, k, kind, dtype=torch.float16, ids="base", opts=None, seed=1234):
    rows, queries = synth(n, d, nq, kind, seed=seed, dtype=str(dtype).split(".")[-1])
    id_arr = None
    if ids == "shuffled":
        id_arr = torch.from_numpy(np.random.RandomState(seed).permutation(n).astype(np.int64) + 1)
    s = _searcher(d, dtype)
    for name, v in (opts or {}).items():
        s.set_option(name, v)
    s.set_shard(rows.to(DEV), None if id_arr is None else id_arr.to(DEV), id_base=1)
    got_s, got_i = s.search(queries.to(DEV), k)
    got_s, got_i = got_s.cpu().numpy(), got_i.cpu().numpy()
    s.close()
    want_s, want_i, ties = _oracle().mips_topk(
        to_oracle_input(rows), to_oracle_input(queries), k,
        ids=None if id_arr is None else id_arr.numpy(), id_base=1, want_ties=True)
    return got_s, got_i, want_s, want_i, ties


EXACT_SHAPES = [
    # n, d, nq, k
    (1000, 128, 32, 5),       # BASELINE configs[0]
    (1, 64, 1, 1),
    (127, 64, 64, 5), (128, 64, 64, 5), (129, 64, 64, 50),
    (5000, 768, 64, 50),
    (40000, 768, 64, 50),     # > 2 tiles per CTA: the probe pass is active
    (3000, 8, 7, 3),          # smallest d
    (3000, 1024, 64, 64),     # largest d, largest k
    (2000, 72, 5, 17),        # d not a multiple of the 64-element K block
    (20000, 768, 65, 50),     # two query passes
    (2500, 256, 130, 10),     # three query passes
]


@pytest.mark.parametrize("n,d,nq,k", EXACT_SHAPES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_exact_arithmetic_inputs_are_bit_exact(n, d, nq, k, dtype):
    got_s, got_i, want_s, want_i, _ = run_case(n, d, nq, k, "X", dtype)
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_s, want_s)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_shuffled_doc_ids_go_through_the_id_map(dtype):
    got_s, got_i, want_s, want_i, ties = run_case(30000, 768, 64, 50, "X", dtype, ids="shuffled")
    assert np.array_equal(got_s, want_s)
    # rows tie-break by row inside the kernel, the oracle by id: compare outside exact ties
    assert_ids_equal_outside_ties(got_i, want_i, ties)


@pytest.mark.parametrize("n,d,nq,k", [(100000, 768, 64, 50), (30000, 128, 64, 50), (777, 64, 16, 7)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gaussian_inputs_match_outside_numerical_ties(n, d, nq, k, dtype):
    got_s, got_i, want_s, want_i, ties = run_case(n, d, nq, k, "G", dtype)
    assert_ids_equal_outside_ties(got_i, want_i, ties)
    assert np.allclose(got_s, want_s, rtol=1e-5, atol=1e-6)   # fp32 accumulation-order noise


@pytest.mark.parametrize("opts", [{"probe": 0}, {"share": 0}, {"max_ctas": 7}, {"max_ctas": 1},
                                  {"probe_timeout_ns": 0}])
def test_threshold_sharing_options_never_change_results(opts):
    got_s, got_i, want_s, want_i, _ = run_case(50000, 768, 64, 50, "X", torch.bfloat16, opts=opts)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s)


def test_short_and_empty_shards_pad_with_minus_inf_and_minus_one():
    got_s, got_i, want_s, want_i, _ = run_case(3, 64, 4, 5, "X")
    assert np.array_equal(got_i, want_i) and np.array_equal(got_s, want_s)
    assert np.all(got_i[:, 3:] == -1) and np.all(np.isneginf(got_s[:, 3:]))
    s = _searcher(64, torch.float16)
    s.set_shard(torch.empty(0, 64, dtype=torch.float16, device=DEV))
    es, ei = s.search(torch.ones(4, 64, dtype=torch.float16, device=DEV), 5)
    assert torch.isneginf(es).all() and (ei == -1).all()
    zs, zi = s.search(torch.ones(0, 64, dtype=torch.float16, device=DEV), 5)
    assert zs.shape == (0, 5) and zi.shape == (0, 5)
    s.close()


def test_nan_rows_are_never_returned_and_duplicates_rank_by_row():
    d = 64
    rows = torch.zeros(300, d, dtype=torch.float16)
    rows[:, 0] = 1.0                       # every row scores exactly q[0]: a 300-way tie
    rows[17, 1] = float("nan")
    q = torch.ones(2, d, dtype=torch.float16)
    s = _searcher(d, torch.float16)
    s.set_shard(rows.to(DEV), None, id_base=1)
    sc, ids = s.search(q.to(DEV), 20)
    s.close()
    want = [i for i in range(1, 40) if i != 18][:20]
    assert ids.cpu().tolist() == [want, want]
    assert (sc == 1.0).all()


def test_error_codes_and_messages():
    from emdr2_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.emdr2_mips_create(100, 0, 0, ctypes.byref(h)) == -1          # d % 8
    assert b"multiple of 8" in lib.emdr2_last_error()
    assert lib.emdr2_mips_create(128, 7, 0, ctypes.byref(h)) == -1          # dtype
    assert lib.emdr2_mips_create(128, 0, 99, ctypes.byref(h)) == -1         # device
    assert lib.emdr2_mips_create(128, 0, 0, ctypes.byref(h)) == 0
    q = torch.zeros(4, 128, dtype=torch.float16, device=DEV)
    out_s = torch.empty(4, 5, device=DEV)
    out_i = torch.empty(4, 5, dtype=torch.int64, device=DEV)
    args = (ctypes.c_void_p(q.data_ptr()), 4, 5, ctypes.c_void_p(out_s.data_ptr()),
            ctypes.c_void_p(out_i.data_ptr()), None)
    assert lib.emdr2_mips_search(h, *args) == -4                            # before set_shard
    rows = torch.zeros(256, 128, dtype=torch.float16, device=DEV)
    assert lib.emdr2_mips_set_shard(h, ctypes.c_void_p(rows.data_ptr() + 2), None, 10, 0) == -1
    assert b"aligned" in lib.emdr2_last_error()
    assert lib.emdr2_mips_set_shard(h, ctypes.c_void_p(rows.data_ptr()), None, 256, 0) == 0
    assert lib.emdr2_mips_search(h, args[0], 4, 0, *args[3:]) == -1         # k < 1
    assert lib.emdr2_mips_search(h, args[0], 4, _lib.MAX_K + 1, *args[3:]) == -1
    assert lib.emdr2_mips_search(h, None, 4, 5, *args[3:]) == -1            # NULL queries
    assert lib.emdr2_mips_search(h, *args) == 0
    torch.cuda.synchronize()
    assert lib.emdr2_mips_destroy(h) == 0
    assert lib.emdr2_mips_search(ctypes.c_void_p(), *args) == -1            # NULL handle
    from emdr2_b200.mips import ShardSearcher
    with pytest.raises(_lib.Emdr2Error):
        ShardSearcher(128, torch.float16, DEV).search(q, 65)


def test_host_buffer_entry_point_equals_device_entry_point():
    rows, queries = synth(20000, 768, 64, "G")
    s = _searcher(768, torch.float16)
    s.set_shard(rows.to(DEV), None, id_base=1)
    ds, di = s.search(queries.to(DEV), 50)
    hs, hi = s.search_host(queries, 50)
    assert not hs.is_cuda and torch.equal(hs, ds.cpu()) and torch.equal(hi, di.cpu())
    s.close()


def test_merge_kernel_equals_oracle_merge():
    from emdr2_b200.mips import merge_topk
    rng = np.random.RandomState(5)
    for parts, nq, k in [(1, 3, 5), (8, 64, 50), (3, 7, 64), (40, 2, 50), (16, 64, 51)]:
        scores = rng.randint(-20, 20, size=(parts, nq, k)).astype(np.float32) / 4   # many ties
        ids = rng.permutation(parts * nq * k).reshape(parts, nq, k).astype(np.int64)
        ids[rng.rand(parts, nq, k) < 0.1] = -1                                         # padding
        scores[rng.rand(parts, nq, k) < 0.02] = np.nan
        gs, gi = merge_topk(torch.from_numpy(scores).to(DEV), torch.from_numpy(ids).to(DEV))
        ws, wi = _oracle().merge_topk(scores, ids)
        assert np.array_equal(gi.cpu().numpy(), wi) and np.array_equal(gs.cpu().numpy(), ws)


# ---------------------------------------------------------------- the reference's own outputs
@pytest.mark.parametrize("name", ["c1_exact", "c1_gauss", "shard3"])
def test_index_class_against_reference_golden(name):
    """B200BruteForceIndex on the inputs the reference's DistributedBruteForceIndex was run on."""
    from emdr2_b200.index import B200BruteForceIndex, B200FaissMIPSIndex
    g = load_golden(name)
    k = int(g["k"])

    class Store(object):
        embedding_path = "unused"

        def __init__(self):
            self.embed_data = {int(i): r for i, r in zip(g["ids"], g["rows"])}

        def clear(self):
            self.embed_data = {}

    index = B200BruteForceIndex(embed_size=g["rows"].shape[1], embed_data=Store())
    q = torch.from_numpy(g["queries"]).to(DEV)
    dist, idx = index.search_mips_index(q, k, reconstruct=False)
    assert dist.dtype == torch.float16 and idx.dtype == torch.int32 and dist.is_cuda and idx.is_cuda
    full32 = exact_scores(g["rows"], g["queries"])
    raw_s, raw_i = index.search(q, k)
    if name == "c1_exact":       # integer scores: fp32, fp16 and the reference all agree exactly
        assert_valid_topk(raw_s.cpu().numpy(), raw_i.cpu().numpy(), full32, g["ids"], k, name)
        assert np.array_equal(dist.cpu().numpy(), g["ref_distances"])
    # our fp16-cast distances equal the reference's (it ranks fp16-rounded scores; values agree
    # except where fp32 accumulation order moves a score across an fp16 rounding boundary)
    assert np.allclose(dist.cpu().numpy().astype(np.float32),
                       g["ref_distances"].astype(np.float32), rtol=2e-3, atol=1e-3)
    # ids equal the reference's wherever its fp16 score is unique among all candidates
    full16 = full32.astype(np.float16)
    ours = idx.cpu().numpy()
    agree = total = 0
    for qi in range(ours.shape[0]):
        for r in range(k):
            if (full16[qi] == g["ref_distances"][qi, r]).sum() == 1 and \
                    (r + 1 == k or g["ref_distances"][qi, r] != g["ref_distances"][qi, k - 1]):
                total += 1
                agree += int(g["ref_indices"][qi, r] in ours[qi])
    assert agree == total and total > 0
    fa = B200FaissMIPSIndex(embed_size=g["rows"].shape[1], embed_data=Store())
    d2, i2 = fa.search_mips_index(torch.from_numpy(g["queries"]), k, reconstruct=False)
    assert isinstance(d2, np.ndarray) and d2.dtype == np.float32 and i2.dtype == np.int64
    assert np.array_equal(i2, raw_i.cpu().numpy())
    d3, i3, rec = fa.search_mips_index(torch.from_numpy(g["queries"]), k, reconstruct=True)
    pos = {int(v): j for j, v in enumerate(g["ids"])}
    assert np.array_equal(rec[0, 0], g["rows"][pos[int(i3[0, 0])]].astype(np.float32))


def test_update_and_reset_index_reload_from_disk(tmp_path):
    from emdr2_b200.index import B200BruteForceIndex
    from emdr2_b200.store import EvidenceStore
    path = str(tmp_path / "ev.pkl")
    rows, queries = synth(600, 128, 8, "X")
    st = EvidenceStore(path, load_from_path=False, rank=0)
    st.add_block_data(range(1, 601), rows.numpy())
    st.save_shard()
    st.merge_shards_and_save()
    index = B200BruteForceIndex(128, EvidenceStore(path))
    a = index.search(queries.to(DEV), 5)[1].cpu()
    st2 = EvidenceStore(path, load_from_path=False, rank=0)      # refreshed index: rows reversed
    st2.add_block_data(range(1, 601), rows.flip(0).numpy())
    st2.save_shard()
    st2.merge_shards_and_save()
    index.update_index()
    b = index.search(queries.to(DEV), 5)[1].cpu()
    assert torch.equal(b, 601 - a)
    index.reset_index()
    assert torch.equal(index.search(queries.to(DEV), 5)[1].cpu(), b)


# ---------------------------------------------------------------- BASELINE full size (configs[1])
def test_full_size_c2_properties():
    """1M x 768 bf16, 64 queries, top-50: too large for the CPU oracle in seconds, so check
    size-independent properties — (a) returned scores are the exact dot products of the returned
    rows (fp64 on the host), sorted; (b) no row outside the result beats the k-th score (independent
    fp32 matmul on the GPU); (c) shard invariance: merging the top-k of two halves reproduces the
    whole; (d) planted rows with known scores come back at the top; (e) idempotence."""
    from emdr2_b200.mips import merge_topk
    n, d, nq, k = 1_000_000, 768, 64, 50
    g = torch.Generator(device=DEV).manual_seed(1234)
    rows = (torch.randn(n, d, generator=g, device=DEV) / d ** 0.5).to(torch.bfloat16)
    queries = torch.randn(nq, d, generator=g, device=DEV).to(torch.bfloat16)
    planted = torch.arange(0, nq, device=DEV) * 15013 + 7
    rows[planted] = (queries.float() * 0.25).to(torch.bfloat16)         # score = |q|^2/4 >> others
    s = _searcher(d, torch.bfloat16)
    s.set_shard(rows, None, id_base=1)
    sc, ids = s.search(queries, k)
    sc2, ids2 = s.search(queries, k)
    assert torch.equal(sc, sc2) and torch.equal(ids, ids2)                         # (e)
    assert torch.equal(ids[:, 0], planted + 1)                                     # (d)
    picked = rows[(ids - 1).view(-1)].view(nq, k, d).double().cpu()
    want = torch.einsum("qkd,qd->qk", picked, queries.double().cpu())
    assert torch.allclose(sc.cpu().double(), want, rtol=1e-5, atol=1e-6)           # (a)
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    kth = sc[:, -1:]
    beat = torch.zeros(nq, dtype=torch.int64, device=DEV)
    for c0 in range(0, n, 1 << 18):
        S = queries.float() @ rows[c0:c0 + (1 << 18)].float().T
        beat += (S > kth * (1 + 1e-5) + 1e-6).sum(1)
    assert (beat <= k - 1).all()                                                   # (b)
    half = n // 2 + 13
    s1, s2 = _searcher(d, torch.bfloat16), _searcher(d, torch.bfloat16)
    s1.set_shard(rows[:half], None, id_base=1)
    s2.set_shard(rows[half:], None, id_base=1 + half)
    a, b = s1.search(queries, k), s2.search(queries, k)
    ms, mi = merge_topk(torch.stack([a[0], b[0]]), torch.stack([a[1], b[1]]))
    assert torch.equal(mi, ids) and torch.equal(ms, sc)                            # (c)
    for x in (s, s1, s2):
        x.close()


# ---------------------------------------------------------------- multi-GPU (NCCL) when present
def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank,
                            world_size=world, device_id=torch.device("cuda", rank))
    from emdr2_b200.index import B200BruteForceIndex
    n, d, nq, k = 50001, 768, 64, 50
    rows, queries = synth(n, d, nq, "X")
    index = B200BruteForceIndex(d)
    index.add_arrays(np.arange(1, n + 1, dtype=np.int64), rows)
    s, i = index.search(queries.cuda(), k)
    np.savez("%s/r%d.npz" % (out_dir, rank), s=s.cpu().numpy(), i=i.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_search_over_nccl_equals_oracle(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows, queries = synth(50001, 768, 64, "X")
    ws, wi = _oracle().mips_topk(to_oracle_input(rows), to_oracle_input(queries), 50, id_base=1)
    for r in range(world):
        z = np.load(str(tmp_path / ("r%d.npz" % r)))
        assert np.array_equal(z["i"], wi) and np.array_equal(z["s"], ws)
