"""One-process-per-GPU sharding of a long signal (torch.distributed is plumbing only).

A mono/interleaved signal is cut into contiguous frame ranges, one per rank; rank r > 0
needs the `halo_frames` frames that precede its range (its left context, >= k - 1 frames:
libmavg asks for whole history tiles so that the sharded result is bit-identical to the
single-GPU one).  Only the INPUT tail of the left neighbour is needed, never its output, so
all ranks filter concurrently.  Two ways to obtain the halo:

  * `exchange_halo`  -- neighbour send/recv through torch.distributed (NCCL over NVLink on
    GPUs, gloo in the CPU tests);
  * `PeerHalo`       -- CUDA IPC: the left neighbour exports its shard buffer, the kernel's
    TMA loads read the tail in place through the peer mapping (no staging copy).

The reference has no multi-GPU path at all; this is new.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(frames: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Frame ranges per rank; interior cuts rounded up to `align` (the tile size), exactly as
    libmavg's single-process multi-device plans cut them (csrc/mavg.cu, mavg_plan_create)."""
    cuts = [0]
    for i in range(1, world):
        f = frames // world * i
        f = (f + align - 1) // align * align
        cuts.append(min(f, frames))
    cuts.append(frames)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def exchange_halo(shard: torch.Tensor, halo_elems: int, rank: int, world: int,
                  group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Sends the last `halo_elems` elements of `shard` to rank+1 and returns the halo received
    from rank-1 (None on rank 0).  Works on any backend; tensors stay on their device."""
    ops = []
    recv = None
    if halo_elems <= 0:
        return None
    # every left neighbour must own at least the halo (mavg_plan_create: "signal too short to shard"): a shorter
    # shard would send fewer elements than the right neighbour waits for
    sizes = [None] * world
    dist.all_gather_object(sizes, int(shard.numel()), group=group)
    short = [r for r in range(world - 1) if sizes[r] < halo_elems]
    if short:
        raise ValueError(f"shard of rank {short[0]} holds {sizes[short[0]]} elements, fewer than the {halo_elems}-element "
                         "left context its right neighbour needs: the signal is too short to shard this way")
    # the bytes travel as uint8: NCCL has no 16-bit integer type (int16 PCM shards)
    if rank + 1 < world:
        tail = shard[-halo_elems:].contiguous()
        ops.append(dist.P2POp(dist.isend, tail.view(torch.uint8), rank + 1, group))
    if rank > 0:
        recv = torch.empty(halo_elems, dtype=shard.dtype, device=shard.device)
        ops.append(dist.P2POp(dist.irecv, recv.view(torch.uint8), rank - 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv


class PeerHalo:
    """Maps the left neighbour's shard buffer into this process (CUDA IPC) and exposes the
    address of its last `halo_elems` elements.  Buffers must come from mavg_device_alloc
    (plain cudaMalloc), not from a caching allocator."""

    def __init__(self, my_ptr: int, my_elems: int, elem_bytes: int, halo_elems: int, rank: int, world: int,
                 group: Optional[dist.ProcessGroup] = None):
        lib = _lib.load()
        self._lib = lib
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(lib.mavg_ipc_export(ctypes.c_void_p(my_ptr), handle))
        mine = (bytes(handle), int(my_elems))
        gathered: List[Optional[tuple]] = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
        self._mapped = ctypes.c_void_p()
        self.halo_ptr = 0
        short = [r for r in range(world - 1) if gathered[r][1] < halo_elems]
        if short:
            raise ValueError(f"shard of rank {short[0]} holds {gathered[short[0]][1]} elements, fewer than the "
                             f"{halo_elems}-element left context its right neighbour needs")
        if rank > 0:
            lh, lelems = gathered[rank - 1]
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(lh)
            _lib.check(lib.mavg_ipc_open(buf, ctypes.byref(self._mapped)))
            self.halo_ptr = int(self._mapped.value) + (lelems - halo_elems) * elem_bytes

    def close(self) -> None:
        if self._mapped:
            self._lib.mavg_ipc_close(self._mapped)
            self._mapped = ctypes.c_void_p()
