"""world_size-2 gloo test (CPU) of the one-process-per-GPU sharding plumbing: shard bounds, the
neighbour halo exchange, and that shard + halo reproduces the unsharded result (checked with the
oracle, since no GPU is present here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, align, halo, q):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, root)
    import oracle
    from digital_signal_processsing_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n, world, align)[rank]
    shard = torch.from_numpy(oracle.fill_f32(hi - lo, 123, first_index=lo))      # each rank generates its slice
    got = sharding.exchange_halo(shard, halo, rank, world)
    if rank == 0:
        assert got is None
        ctx = shard.numpy()
        y = oracle.mavg_f64(ctx, k)
    else:
        assert got.numel() == halo
        expect = oracle.fill_f32(halo, 123, first_index=lo - halo)
        assert np.array_equal(got.numpy(), expect)
        ctx = np.concatenate([got.numpy(), shard.numpy()])
        y = oracle.mavg_f64(ctx, k)[halo:]           # halo >= k-1: warm-up falls inside the halo
    # max over ranks of a fake timing, as bench.py does it
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, lo, hi, y, float(t[0])))
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [3, 1024])
def test_two_rank_halo_exchange(oracle_mod, k):
    world, align = 2, 4096
    n = 10 * align + 777
    halo = max(align, (k + align - 1) // align * align)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, align, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert parts[0][1] == 0 and parts[0][2] == parts[1][1] and parts[1][2] == n and parts[1][1] % align == 0
    y = np.concatenate([p[3] for p in parts])
    full = oracle_mod.mavg_f64(oracle_mod.fill_f32(n, 123), k)
    np.testing.assert_allclose(y, full, rtol=1e-12, atol=1e-12)
    assert all(p[4] == 2.0 for p in parts)


def test_shard_bounds_match_library_rule():
    from digital_signal_processsing_b200.sharding import shard_bounds
    assert shard_bounds(100, 1) == [(0, 100)]
    b = shard_bounds(1 << 32, 8, 8192)
    assert b[0] == (0, 1 << 29) and b[-1][1] == 1 << 32 and all(lo % 8192 == 0 for lo, _ in b)
    b = shard_bounds(10 * 4096 + 5, 4, 4096)
    assert [lo for lo, _ in b] == [0, 12288, 24576, 32768] and b[-1][1] == 10 * 4096 + 5
    assert sum(hi - lo for lo, hi in b) == 10 * 4096 + 5


def _short_worker(rank, world, port, q):
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, root)
    from digital_signal_processsing_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = torch.zeros(100 if rank == 0 else 5000)          # rank 0 owns fewer elements than the halo
    try:
        sharding.exchange_halo(shard, 4096, rank, world)
        q.put((rank, "no error"))
    except ValueError as e:
        q.put((rank, str(e)))
    dist.destroy_process_group()


def test_halo_longer_than_left_shard_is_an_error_on_every_rank():
    """A left shard shorter than the halo used to yield a short send and a hang; now every rank
    raises the same ValueError before any send/recv is posted."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_short_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all("too short to shard" in got[r] and "rank 0" in got[r] for r in range(world)), got
