"""world_size = 2 over gloo on CPU: the N > 1 host path (sharding of the slice
list and the gather of the per-slice summaries on rank 0)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from qunundrum_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.partition(n, world, rank)
    # a stand-in for the device summaries: recognisable per-slice values
    summ = np.zeros((len(idx), 8))
    summ[:, 0] = idx * 0.5
    summ[:, 4] = 1.0
    summ[:, 7] = rank
    table = shard.gather_summaries(idx, summ, n)
    dist.barrier()
    if rank == 0:
        q.put(table)
    else:
        assert table is None
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    n, world = 37, 2
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert table.shape == (n, 8)
    assert np.array_equal(table[:, 0], np.arange(n) * 0.5)
    assert np.all(table[:, 4] == 1.0)
    i = np.arange(n)
    assert np.array_equal(table[:, 7], (i + i // world) % world)  # slice i -> rank (i + i // world) mod world


def test_gather_without_process_group_is_identity():
    from qunundrum_b200 import shard
    idx = shard.partition(5, 1, 0)
    s = np.arange(40.0).reshape(5, 8)
    assert np.array_equal(shard.gather_summaries(idx, s, 5), s)


def _worker_weak(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from qunundrum_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    summ = np.full((n, 8), float(rank + 1))
    table = shard.gather_summaries(rank * n + np.arange(n), summ, n * world)   # as bench.py does
    if rank == 0:
        q.put(table)
    dist.barrier()
    dist.destroy_process_group()


def test_weak_scaling_gather_world2():
    """bench.py's layout: every rank owns a whole distribution (n slices)."""
    n, world = 11, 2
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_weak, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    table = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert table.shape == (n * world, 8)
    assert np.all(table[:n] == 1.0) and np.all(table[n:] == 2.0)
