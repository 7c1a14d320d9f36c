"""Bucketed, backward-overlapped gradient all-reduce (emdr2_b200/data_parallel.py) on CPU: a world-size-2
Gloo run must leave every rank with the SUM of the per-rank gradients in its flat buffers (reference
megatron/model/distributed.py:35-63 computes the same sum with one flat all-reduce after backward), gradients
must be views of the flat buffers, and parameters re-homed by `flatten_parameters` must keep their values."""
import os
import socket

import numpy as np
import torch

from emdr2_b200.data_parallel import GradientBuckets, flatten_parameters


def _model():
    torch.manual_seed(5)
    m = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Tanh(), torch.nn.Linear(40, 40), torch.nn.Tanh(),
                            torch.nn.Linear(40, 8))
    unused = torch.nn.Parameter(torch.ones(13))          # a parameter no loss depends on (e.g. T5 token types)
    m.register_parameter("unused", unused)
    return m


def _loss(m, rank, step):
    g = torch.Generator().manual_seed(100 * step + rank)
    x = torch.randn(6, 24, generator=g)
    return (m(x) ** 2).sum() * (rank + 1)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    m = _model()
    before = [p.detach().clone() for p in m.parameters()]
    gb = GradientBuckets(list(m.parameters()), bucket_bytes=4096)       # several small buckets
    flats = flatten_parameters(gb)
    assert len(gb.buckets) >= 3 and len(flats) == len(gb.buckets)
    for p, q in zip(m.parameters(), before):
        assert torch.equal(p.detach(), q)
    out = {}
    for step in range(2):                                                # buffers are reused across steps
        gb.start_step()
        _loss(m, rank, step).backward()
        launched_in_backward = gb.launched
        gb.finish()
        for p in m.parameters():
            b = gb._bucket_of[p]
            assert p.grad.data_ptr() >= b.grad.data_ptr() and p.grad.data_ptr() < b.grad.data_ptr() + b.grad.numel() * 4
        out["g%d" % step] = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy()
        out["hooks%d" % step] = launched_in_backward
    # parameters are views of the flat buffers: an update of the buffer is an update of the model
    with torch.no_grad():
        for f in flats:
            f.add_(1.0)
    out["moved"] = float(sum((p.detach() - q).sum() for p, q in zip(m.parameters(), before)))
    out["count"] = sum(p.numel() for p in m.parameters())
    np.savez(os.path.join(out_dir, "dp%d.npz" % rank), **out)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_world2_bucketed_allreduce_equals_the_sum_of_rank_gradients(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    for attempt in range(2):
        try:
            mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt:
                raise
    for step in range(2):
        want = None
        for rank in range(world):
            m = _model()
            _loss(m, rank, step).backward()
            g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])
            want = g if want is None else want + g
        for rank in range(world):
            r = np.load(str(tmp_path / ("dp%d.npz" % rank)))
            assert np.allclose(r["g%d" % step], want.numpy(), rtol=1e-5, atol=1e-6)
            assert int(r["hooks%d" % step]) >= 2            # most buckets left from inside backward
    r = np.load(str(tmp_path / "dp0.npz"))
    assert abs(float(r["moved"]) - int(r["count"])) < 1e-3


def test_single_process_buckets_are_plain_gradient_views():
    m = _model()
    gb = GradientBuckets(list(m.parameters()), bucket_bytes=1 << 20)
    gb.start_step()
    _loss(m, 0, 0).backward()
    gb.finish()
    ref = _model()
    _loss(ref, 0, 0).backward()
    for p, q in zip(m.parameters(), ref.parameters()):
        want = q.grad if q.grad is not None else torch.zeros_like(q)
        assert torch.allclose(p.grad, want)
    assert gb.world == 1 and gb.launched == 0
    gb.close()
