"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): one process per GPU under
torch.distributed (NCCL), contiguous shards of one synthetic signal, left context read in place
from the neighbour over CUDA IPC, output of every rank spot-checked against the fp64 oracle
recomputed from the generator -- and bit-identical to the single-GPU run of the same signal."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r'''
import ctypes, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["MAVG_ROOT"])
import digital_signal_processsing_b200 as mavg
from digital_signal_processsing_b200 import _lib, sharding
import oracle

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
lib = _lib.load()
n_per, seed = 1 << 24, 4242
for k, mode in ((3, "ipc"), (1024, "ipc"), (4096, "nccl")):
    first = rank * n_per
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(lib.mavg_device_alloc(4 * n_per, ctypes.byref(d_in)))
    _lib.check(lib.mavg_device_alloc(4 * n_per, ctypes.byref(d_out)))
    mavg.fill_synthetic_device(d_in.value, "f32", n_per, first, seed)
    torch.cuda.synchronize()
    plan = mavg.Plan(n_per, k, first_frame=first)
    halo = int(plan.info.halo_frames)
    halo_ptr, peer, staged = None, None, None
    if mode == "ipc":
        peer = sharding.PeerHalo(d_in.value, n_per, 4, halo, rank, world)
        halo_ptr = peer.halo_ptr or None
    else:
        class _Arr:
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        shard = torch.as_tensor(_Arr(d_in.value, n_per), device="cuda")
        staged = sharding.exchange_halo(shard, halo, rank, world)
        torch.cuda.synchronize()
        halo_ptr = staged.data_ptr() if staged is not None else None
    dist.barrier()
    plan.run_device_halo(d_in.value, d_out.value, halo_ptr)
    plan.synchronize()
    out = torch.empty(n_per, dtype=torch.float32, device="cuda")
    class _Out:
        __cuda_array_interface__ = {"shape": (n_per,), "typestr": "<f4", "data": (d_out.value, False), "version": 3}
    y = torch.as_tensor(_Out(), device="cuda").cpu().numpy()
    # spot checks: shard boundary +-k and random interior positions, against the generator-driven oracle
    rng = np.random.default_rng(rank)
    idx = np.unique(np.concatenate([np.arange(0, min(n_per, 2 * k + 64)), rng.integers(0, n_per, 2000), [n_per - 1]]))
    worst = 0.0
    for i in idx:
        e = oracle.point_f64(first + int(i), k, seed)
        worst = max(worst, abs(float(y[i]) - e) / abs(e))
    assert worst < 1e-5, (rank, k, worst)
    # bit-identity with an unsharded run of the same global signal on this GPU
    if rank == world - 1:
        n_all = world * n_per
        x = torch.empty(n_all, dtype=torch.float32, device="cuda")
        mavg.fill_synthetic_device(x.data_ptr(), "f32", n_all, 0, seed)
        z = torch.empty_like(x)
        torch.cuda.synchronize()
        with mavg.Plan(n_all, k) as whole:
            whole.run_device([x.data_ptr()], [z.data_ptr()])
            whole.synchronize()
        assert np.array_equal(z[first:].cpu().numpy(), y), (k, "sharded run differs from single-GPU run")
    dist.barrier()
    plan.close()
    if peer is not None:
        peer.close()
    dist.barrier()
    lib.mavg_device_free(d_in); lib.mavg_device_free(d_out)
print(f"rank {rank} ok")
dist.destroy_process_group()
'''


def test_sharded_signal_across_gpus(mavg, oracle_mod, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, MAVG_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == world
