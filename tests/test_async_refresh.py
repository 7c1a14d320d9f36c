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
