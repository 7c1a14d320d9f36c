"""Asynchronous index refresh (BASELINE configs[4]) on CPU: the two-flag Gloo hand-shake between
trainer and indexer ranks (reference tasks/openqa/e2eqa/async_indexer.py:116-144,
train_e2eqa.py:436-505), the reference-compatible pickle hand-over and the direct shard hand-over.
The context tower is replaced by a deterministic stand-in whose output depends on the "weights
version", so the test can tell which checkpoint every published index was built from."""
import os
import pickle
import socket

import numpy as np
import pytest
import torch

from emdr2_b200.async_indexer import AsyncIndexBuilder, RefreshProtocol, ShardReceiver, owner_of_rows
from emdr2_b200.index import chunk_range

N_DOCS, DIM = 37, 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class StubTower(torch.nn.Module):
    """context_model stand-in: embedding = f(first token, weights version)."""

    def __init__(self):
        super().__init__()
        self.version = torch.nn.Parameter(torch.zeros(1))

    def forward(self, tokens, mask, types, max_len=None, row_lengths=None):
        base = tokens[:, :1].float() * torch.arange(1, DIM + 1).float()[None]
        return (base + 1000.0 * self.version.detach()).to(torch.float16)


class StubDual(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.context_model = StubTower()


def _expected_rows(doc_ids, version):
    return (np.asarray(doc_ids, np.float32)[:, None] * np.arange(1, DIM + 1, dtype=np.float32)[None]
            + 1000.0 * version).astype(np.float16)


def _batches(index_rank, index_world, batch=5):
    """This indexer's share of the evidence, in DistributedBatchSampler order (each global batch is
    cut into per-rank slices, indexer_emdr2.py:16-35): doc ids are 1-based."""
    def gen():
        ids = np.arange(1, N_DOCS + 1)
        for lo in range(0, N_DOCS, batch * index_world):
            chunk = ids[lo:lo + batch * index_world]
            mine = chunk[index_rank * batch:(index_rank + 1) * batch]
            if mine.size == 0:
                continue
            tokens = torch.from_numpy(np.stack([mine, mine + 1], axis=1))
            yield torch.from_numpy(mine), tokens, torch.zeros_like(tokens)
    return gen


def _run(rank, world, n_train, port, tmp, mode, rounds, interval):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    flags = dist.new_group(list(range(world)), backend="gloo")                    # initialize.py:261
    train_group = dist.new_group(list(range(n_train)), backend="gloo")
    index_group = dist.new_group(list(range(n_train, world)), backend="gloo")
    proto = RefreshProtocol(rank, world, n_train, group=flags)
    ckpt = os.path.join(tmp, "ckpt.pt")
    path = os.path.join(tmp, "evidence.pkl")
    log = []
    if proto.is_trainer:
        weights_version = [0.0]

        def save_checkpoint(iteration):
            if rank == 0:
                torch.save({"version": weights_version[0], "iteration": iteration}, ckpt)

        receiver = ShardReceiver(N_DOCS, DIM, rank, n_train) if mode == "direct" else None

        class FakeIndex(object):
            def add_local_shard(self, ids, rows, num_rows=None, row_lo=0):
                log.append(("swap", ids.clone().numpy(), rows.clone().numpy(), row_lo))

        fake = FakeIndex()

        def update_index():
            if mode == "store":
                with open(path, "rb") as f:
                    data = pickle.load(f)["embed_data"]
                log.append(("reload", sorted(data), np.stack([data[k] for k in sorted(data)])))
            else:
                receiver.swap_into(fake)

        def receive_index():
            receiver.receive_from(list(range(n_train, world)))

        save_checkpoint(0)
        dist.barrier(train_group)
        proto.trainer_start(iteration=0)
        iteration = 0
        while proto.reloads < rounds:
            weights_version[0] = float(iteration + 1)             # "training" changes the weights
            happened = proto.trainer_maybe_reload(iteration, interval, save_checkpoint, update_index,
                                                  poll_seconds=0.01,
                                                  receive_index=receive_index if mode == "direct" else None)
            if happened:
                log.append(("handover", iteration, weights_version[0]))
            iteration += 1
        assert proto.last_reload_iteration >= interval
    else:
        model = StubDual()

        def load_weights(m):
            state = torch.load(ckpt)
            with torch.no_grad():
                m.context_model.version.fill_(state["version"])
            log.append(("loaded", state["version"]))

        builder = AsyncIndexBuilder(model, _batches(rank - n_train, world - n_train), proto,
                                    load_weights=load_weights, embedding_path=path,
                                    index_rank=rank - n_train, index_world=world - n_train,
                                    index_group=index_group, mode=mode, num_rows=N_DOCS)
        builder.run_async(max_rounds=rounds)
        assert builder.rounds == rounds
    with open(os.path.join(tmp, "log%d.pkl" % rank), "wb") as f:
        pickle.dump(log, f)
    dist.barrier()
    dist.destroy_process_group()


def _spawn(world, n_train, tmp, mode, rounds, interval, attempts=2):
    """mp.spawn with one retry on a fresh port: the rendezvous port is picked by bind-and-release, which
    another process on the machine can win in between."""
    import torch.multiprocessing as mp
    for attempt in range(attempts):
        try:
            mp.spawn(_run, args=(world, n_train, _free_port(), tmp, mode, rounds, interval), nprocs=world, join=True)
            return
        except Exception:
            if attempt + 1 == attempts:
                raise
            for name in os.listdir(tmp):          # leave no half-written state behind
                path = os.path.join(tmp, name)
                if os.path.isfile(path):
                    os.remove(path)


def _logs(tmp, world):
    out = []
    for r in range(world):
        with open(os.path.join(tmp, "log%d.pkl" % r), "rb") as f:
            out.append(pickle.load(f))
    return out


def test_store_handover_world3_gloo(tmp_path):
    """1 trainer + 2 indexers, the reference's pickle hand-over: every published index holds every
    document once, built with the weights of the checkpoint saved at the previous hand-over."""
    world, n_train, rounds = 3, 1, 3
    _spawn(world, n_train, str(tmp_path), "store", rounds, 2)
    logs = _logs(str(tmp_path), world)
    reloads = [e for e in logs[0] if e[0] == "reload"]
    handovers = [e for e in logs[0] if e[0] == "handover"]
    assert len(reloads) == rounds == len(handovers)
    built_with = 0.0                                       # first build: the initial checkpoint
    for (_, ids, rows), (_, iteration, version) in zip(reloads, handovers):
        assert ids == list(range(1, N_DOCS + 1))
        assert np.array_equal(rows, _expected_rows(ids, built_with))
        built_with = version                               # the checkpoint saved at this hand-over
    for r in (1, 2):                                       # each indexer reloaded weights after each round
        assert [e[1] for e in logs[r] if e[0] == "loaded"] == [h[2] for h in handovers]
    assert all(h[1] >= 2 for h in handovers) and handovers[1][1] - handovers[0][1] >= 2   # reload interval


def test_direct_handover_world4_gloo(tmp_path):
    """2 trainers + 2 indexers, direct hand-over: each trainer ends up with exactly its torch.chunk
    row range, ids attached, without any file."""
    world, n_train, rounds = 4, 2, 2
    _spawn(world, n_train, str(tmp_path), "direct", rounds, 1)
    logs = _logs(str(tmp_path), world)
    assert not os.path.exists(str(tmp_path / "evidence.pkl"))
    for t in range(n_train):
        swaps = [e for e in logs[t] if e[0] == "swap"]
        handovers = [e for e in logs[t] if e[0] == "handover"]
        assert len(swaps) == rounds
        lo, hi = chunk_range(N_DOCS, n_train, t)
        built_with = 0.0
        for (_, ids, rows, row_lo), (_, _, version) in zip(swaps, handovers):
            assert row_lo == lo and ids.tolist() == list(range(lo + 1, hi + 1))
            assert np.array_equal(rows, _expected_rows(ids, built_with))
            built_with = version
    own = owner_of_rows(N_DOCS, n_train)
    assert [int((own == t).sum()) for t in range(n_train)] == [hi - lo for lo, hi in
                                                               (chunk_range(N_DOCS, n_train, t) for t in range(n_train))]


def test_protocol_argument_checks():
    with pytest.raises(ValueError):
        RefreshProtocol(0, 2, 2)
    with pytest.raises(ValueError):
        RefreshProtocol(0, 2, 0)
    with pytest.raises(ValueError):
        AsyncIndexBuilder(StubDual(), None, RefreshProtocol(1, 2, 1), mode="direct")


# ------------------------------------------------ same devices, second stream (ConcurrentShardRefresher)
class _ScaledTower(torch.nn.Module):
    """Context tower stand-in: embedding = scale * onehot-ish function of the first token (exact arithmetic)."""

    def __init__(self, dim, scale):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor([float(scale)]))
        self.dim = dim

    def forward(self, tokens, mask, types, max_len=None, row_lengths=None):
        t = tokens[:, :1].float()
        cols = torch.arange(1, self.dim + 1).float()[None]
        return (((t * cols) % 17.0 - 8.0) / 8.0 * self.scale.detach()).to(torch.float16)


def _refresh_setup(n, d, world=1, rank=0, group=None, delay=0.0):
    from test_host_logic import CpuDoubleIndex
    from emdr2_b200.async_indexer import ConcurrentShardRefresher
    import time as _time
    tower = _ScaledTower(d, 1.0)
    doc_ids = np.arange(1, n + 1, dtype=np.int64)
    old_rows = tower(torch.from_numpy(doc_ids)[:, None], None, None).numpy()
    index = CpuDoubleIndex(d, device="cpu", group=group)
    index.add_arrays(doc_ids, old_rows)
    lo, hi = index.row_lo, index.row_hi

    def make_batches():
        for a in range(lo, hi, 7):
            ids = doc_ids[a:min(hi, a + 7)]
            if delay:
                _time.sleep(delay)
            yield torch.from_numpy(ids), torch.from_numpy(ids)[:, None].repeat(1, 4), torch.zeros(len(ids), 4, dtype=torch.int64)

    refresher = ConcurrentShardRefresher(index, tower, make_batches, group=group)
    return index, tower, refresher, doc_ids, old_rows


def test_concurrent_refresh_swaps_whole_shards_and_searches_never_see_a_mix():
    """c5 on one process: the refresher thread re-encodes the shard with the tower's CURRENT weights into a standby
    buffer while another thread keeps searching; every search result equals the old index's answer or the new
    one's, never a blend, and after the swap the index serves exactly the new embeddings."""
    import threading
    from oracle import mips as oracle
    n, d, k = 400, 16, 9
    index, tower, refresher, doc_ids, old_rows = _refresh_setup(n, d, delay=0.002)
    queries = torch.from_numpy((np.random.RandomState(3).randint(-8, 9, size=(5, d)) / 8).astype(np.float16))
    with torch.no_grad():
        tower.scale.fill_(-1.0)                               # "training" moved the weights: new rows = -old rows
    new_rows = (-old_rows.astype(np.float32)).astype(np.float16)
    want_old = oracle.mips_topk(old_rows, queries.numpy(), k, ids=doc_ids)
    want_new = oracle.mips_topk(new_rows, queries.numpy(), k, ids=doc_ids)
    assert not np.array_equal(want_old[1], want_new[1])
    seen, stop = [], threading.Event()

    def hammer():
        while not stop.is_set():
            s, i = index.search(queries, k)
            seen.append((s.numpy().copy(), i.numpy().copy(), index.generation))

    th = threading.Thread(target=hammer)
    th.start()
    refresher.start()
    with torch.no_grad():
        tower.scale.fill_(5.0)                                # later training steps must not leak into this refresh
    swapped = False
    for _ in range(2000):
        if refresher.maybe_swap():
            swapped = True
            break
        import time as _time
        _time.sleep(0.005)
    for _ in range(20):
        index.search(queries, k)
    stop.set()
    th.join()
    assert swapped and refresher.rounds == 1 and refresher.rows_done == n
    olds = news = 0
    for s, i, _gen in seen:
        if np.array_equal(i, want_old[1]) and np.array_equal(s, want_old[0]):
            olds += 1
        elif np.array_equal(i, want_new[1]) and np.array_equal(s, want_new[0]):
            news += 1
        else:
            raise AssertionError("a search saw a mixture of the old and the new shard")
    assert olds > 0 and news > 0
    s, i = index.search(queries, k)
    assert np.array_equal(i.numpy(), want_new[1]) and np.array_equal(s.numpy(), want_new[0])
    # second round reuses the retired buffer as the standby
    retired = refresher.standby
    assert np.array_equal(retired.numpy(), old_rows)
    refresher.start()
    while not refresher.maybe_swap():
        pass
    assert refresher.rounds == 2 and refresher.standby is not retired


def _worker_concurrent(rank, world, port, out_dir):
    import torch.distributed as dist
    import time as _time
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n, d, k = 300, 16, 7
    # rank 1 re-encodes more slowly: rank 0 must NOT swap before rank 1 is ready too
    index, tower, refresher, doc_ids, old_rows = _refresh_setup(n, d, world, rank, dist.group.WORLD,
                                                                delay=0.001 if rank == 0 else 0.01)
    queries = torch.from_numpy((np.random.RandomState(3).randint(-8, 9, size=(4, d)) / 8).astype(np.float16))
    with torch.no_grad():
        tower.scale.fill_(-1.0)
    refresher.start()
    results, swaps_at = [], None
    for step in range(400):
        s, i = index.search(queries, k)                       # collective: both ranks, same step
        results.append(i.numpy().copy())
        if refresher.maybe_swap():
            swaps_at = step
            break
        _time.sleep(0.002)
    s, i = index.search(queries, k)
    np.savez(os.path.join(out_dir, "c%d.npz" % rank), results=np.stack(results), final=i.numpy(), swaps_at=swaps_at,
             finals=s.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_concurrent_refresh_world2_all_ranks_swap_between_the_same_two_searches(tmp_path):
    import torch.multiprocessing as mp
    from oracle import mips as oracle
    for attempt in range(2):
        try:
            mp.spawn(_worker_concurrent, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
            break
        except Exception:
            if attempt:
                raise
    n, d, k = 300, 16, 7
    doc_ids = np.arange(1, n + 1, dtype=np.int64)
    old_rows = _ScaledTower(d, 1.0)(torch.from_numpy(doc_ids)[:, None], None, None).numpy()
    queries = (np.random.RandomState(3).randint(-8, 9, size=(4, d)) / 8).astype(np.float16)
    want_old = oracle.mips_topk(old_rows, queries, k, ids=doc_ids)[1]
    want_new = oracle.mips_topk((-old_rows.astype(np.float32)).astype(np.float16), queries, k, ids=doc_ids)[1]
    r0, r1 = (np.load(str(tmp_path / ("c%d.npz" % r)), allow_pickle=True) for r in range(2))
    assert int(r0["swaps_at"]) == int(r1["swaps_at"])        # agreed swap point
    assert np.array_equal(r0["results"], r1["results"])      # both ranks saw the same merged answers throughout
    for res in r0["results"]:
        assert np.array_equal(res, want_old)                  # ... and all of them came from the OLD index
    assert np.array_equal(r0["final"], want_new) and np.array_equal(r1["final"], want_new)
