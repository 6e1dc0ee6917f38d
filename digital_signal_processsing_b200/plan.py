"""Python face of a libmavg plan.

Mirrors the role of the reference's `DspWorkspace` + `XxxGpuLoad` pair
(gpu_utils.h:67-160, basics/profilable_sm_vload4.cu:90-145): a plan is created once per
(signal shape, window) and reused for every timed run; `run_host` is the H2D + kernel +
D2H call, `run_device` the kernel-only call on device-resident shards.  All compute is in
libmavg's CUDA kernels -- nothing here touches sample values.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Desc, Info, Timing, check

_DTYPES = {"f32": _lib.F32, "float32": _lib.F32, "i16": _lib.I16, "int16": _lib.I16}
_LAYOUTS = {"interleaved": _lib.INTERLEAVED, "planar": _lib.PLANAR}
_PATHS = {"auto": _lib.PATH_AUTO, "stream": _lib.PATH_STREAM, "generic": _lib.PATH_GENERIC}
_NP = {_lib.F32: np.float32, _lib.I16: np.int16}
_OPS = {"mean": _lib.OP_MEAN, "average": _lib.OP_MEAN, "rms": _lib.OP_RMS}


def device_count() -> int:
    return int(_lib.load().mavg_device_count())


def version() -> int:
    return int(_lib.load().mavg_version())


class Plan:
    def __init__(self, frames: int, window: int, channels: int = 1, dtype: str = "f32",
                 layout: str = "interleaved", block_size: int = 0, path: str = "auto",
                 devices: Optional[Sequence[int]] = None, first_frame: int = 0, op: str = "mean", **tuning: int):
        self._lib = _lib.load()
        d = Desc()
        d.struct_size = ctypes.sizeof(Desc)
        d.dtype = _DTYPES[dtype]
        d.layout = _LAYOUTS[layout]
        d.channels = channels
        d.frames = frames
        d.window = window
        d.block_size = block_size
        d.path = _PATHS[path]
        if devices:
            d.num_devices = len(devices)
            for i, dev in enumerate(devices):
                d.devices[i] = dev
        d.first_frame = first_frame
        d.op = _OPS[op]
        for key, val in tuning.items():
            if not hasattr(d.tuning, key):
                raise TypeError(f"unknown tuning field {key!r}")
            setattr(d.tuning, key, int(val))
        self.desc = d
        self._h = ctypes.c_void_p()
        check(self._lib.mavg_plan_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.np_dtype = _NP[d.dtype]
        self.samples = frames * channels

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mavg_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ queries
    @property
    def info(self) -> Info:
        i = Info()
        check(self._lib.mavg_plan_info(self._h, ctypes.byref(i)))
        return i

    def timing(self) -> Timing:
        t = Timing()
        check(self._lib.mavg_get_timing(self._h, ctypes.byref(t)))
        return t

    # ------------------------------------------------------------------ runs
    def run_host(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """H2D + kernel + D2H on a host array holding the whole signal in the plan's layout."""
        x = np.ascontiguousarray(x, dtype=self.np_dtype).reshape(-1)
        if x.size != self.samples:
            raise ValueError(f"expected {self.samples} samples, got {x.size}")
        if out is None:
            out = np.empty_like(x)
        if out.dtype != self.np_dtype or out.size != x.size or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of the plan's dtype and size")
        check(self._lib.mavg_run_host(self._h, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(out.ctypes.data)))
        return out

    def run_host_with_context(self, buf: np.ndarray, context_frames: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Shard plans (first_frame > 0): `buf` holds `context_frames` (= info.halo_frames) frames of left context
        followed by the shard; mavg_run_host reads the context in front of the pointer it is given."""
        buf = np.ascontiguousarray(buf, dtype=self.np_dtype).reshape(-1)
        ctx = context_frames * int(self.desc.channels)
        if buf.size != ctx + self.samples:
            raise ValueError(f"expected {ctx} + {self.samples} samples, got {buf.size}")
        if out is None:
            out = np.empty(self.samples, dtype=self.np_dtype)
        check(self._lib.mavg_run_host(self._h, ctypes.c_void_p(buf.ctypes.data + ctx * buf.itemsize),
                                      ctypes.c_void_p(out.ctypes.data)))
        return out

    def run_host_ptr(self, in_ptr: int, out_ptr: int) -> None:
        check(self._lib.mavg_run_host(self._h, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr)))

    def run_device(self, in_ptrs: Sequence[int], out_ptrs: Sequence[int]) -> None:
        n = len(in_ptrs)
        ins = (ctypes.c_void_p * n)(*in_ptrs)
        outs = (ctypes.c_void_p * n)(*out_ptrs)
        check(self._lib.mavg_run_device(self._h, ins, outs))

    def run_cascade(self, in_ptrs: Sequence[int], out_ptrs: Sequence[int], passes: int,
                    scratch_ptrs: Optional[Sequence[int]] = None) -> None:
        """The moving average applied `passes` times in a row (box-filter cascade); the last pass lands in out."""
        n = len(in_ptrs)
        ins = (ctypes.c_void_p * n)(*in_ptrs)
        outs = (ctypes.c_void_p * n)(*out_ptrs)
        scr = (ctypes.c_void_p * n)(*scratch_ptrs) if scratch_ptrs is not None else None
        check(self._lib.mavg_run_cascade(self._h, ins, outs, scr, passes))

    def run_device_halo(self, in_ptr: int, out_ptr: int, halo_ptr: Optional[int]) -> None:
        check(self._lib.mavg_run_device_halo(self._h, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr),
                                             ctypes.c_void_p(halo_ptr or 0)))

    def run_owned(self) -> None:
        check(self._lib.mavg_run_owned(self._h))

    def synchronize(self) -> None:
        check(self._lib.mavg_synchronize(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        check(self._lib.mavg_set_stream(self._h, ctypes.c_void_p(cuda_stream)))

    def enable_timing(self, on: bool) -> None:
        check(self._lib.mavg_enable_timing(self._h, 1 if on else 0))

    def buffers(self, rank: int = 0) -> tuple[int, int]:
        a, b = ctypes.c_void_p(), ctypes.c_void_p()
        check(self._lib.mavg_plan_buffers(self._h, rank, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value or 0), int(b.value or 0)

    def fill_synthetic(self, seed: int, dist: int = _lib.DIST_U01) -> None:
        check(self._lib.mavg_fill_synthetic(self._h, seed, dist))


def fill_synthetic_device(ptr: int, dtype: str, n: int, first_index: int, seed: int, dist: int = _lib.DIST_U01,
                          stream: int = 0) -> None:
    check(_lib.load().mavg_fill_synthetic_device(ctypes.c_void_p(ptr), _DTYPES[dtype], n, first_index, seed, dist,
                                                 ctypes.c_void_p(stream)))


class PinnedArray:
    """Page-locked host buffer (mavg_host_alloc) viewed as a NumPy array: copies to and from it run at full
    PCIe speed and overlap with the kernels inside `Plan.run_host`.  Replaces the host half of the reference's
    `MemoryTraits` (gpu_utils.h:33-65).  Keep the object alive while `.array` is in use."""

    def __init__(self, count: int, dtype):
        self._lib = _lib.load()
        dt = np.dtype(dtype)
        self._ptr = ctypes.c_void_p()
        check(self._lib.mavg_host_alloc(max(1, count) * dt.itemsize, ctypes.byref(self._ptr)))
        buf = (ctypes.c_uint8 * (count * dt.itemsize)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dt, count=count)

    def close(self) -> None:
        if self._ptr:
            self.array = None
            self._lib.mavg_host_free(self._ptr)
            self._ptr = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class pinned:
    """Context manager that page-locks NumPy arrays the caller already owns (mavg_host_register) for the duration
    of the block, so that `Plan.run_host` on them runs at pinned-copy speed:

        with mavg.pinned(x, y):
            for k in windows:
                plans[k].run_host(x, out=y)
    """

    def __init__(self, *arrays: np.ndarray):
        self._lib = _lib.load()
        self._arrays = arrays
        self._done = []

    def __enter__(self):
        for a in self._arrays:
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("only contiguous arrays can be page-locked")
            if a.nbytes == 0:
                continue
            ptr = ctypes.c_void_p(a.ctypes.data)
            try:
                check(self._lib.mavg_host_register(ptr, a.nbytes))
            except Exception:
                self.__exit__()
                raise
            self._done.append(ptr)
        return self

    def __exit__(self, *exc):
        while self._done:
            self._lib.mavg_host_unregister(self._done.pop())


def prefix_sum_device(in_ptr: int, out_ptr: int, dtype: str, frames: int, channels: int = 1, stream: int = 0) -> None:
    """Per-channel inclusive prefix sum on device buffers (int16 -> int64, float32 -> float64), one pass."""
    check(_lib.load().mavg_prefix_sum(_DTYPES[dtype], ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), frames,
                                      channels, ctypes.c_void_p(stream)))


def run_host_sweep_ptr(plans: Sequence["Plan"], in_ptr: int, out_ptrs: Sequence[int]) -> None:
    """mavg_run_host_sweep on raw host addresses (pinned buffers): one upload of the input, every plan's result
    downloaded to its own buffer."""
    n = len(plans)
    if n == 0 or n != len(out_ptrs):
        raise ValueError("one output per plan")
    handles = (ctypes.c_void_p * n)(*[p._h.value for p in plans])
    outs = (ctypes.c_void_p * n)(*[int(o) for o in out_ptrs])
    check(_lib.load().mavg_run_host_sweep(handles, n, ctypes.c_void_p(in_ptr), outs))


def run_host_sweep(plans: Sequence["Plan"], x: np.ndarray, outs: Optional[Sequence[np.ndarray]] = None) -> list:
    """Several plans over the SAME host signal (a sweep over windows): the input crosses the host link once, the
    results are bit-identical to one `Plan.run_host` per plan."""
    if not plans:
        return []
    x = np.ascontiguousarray(x, dtype=plans[0].np_dtype).reshape(-1)
    if x.size != plans[0].samples:
        raise ValueError(f"expected {plans[0].samples} samples, got {x.size}")
    if outs is None:
        outs = [np.empty_like(x) for _ in plans]
    for o in outs:
        if o.dtype != x.dtype or o.size != x.size or not o.flags.c_contiguous:
            raise ValueError("outputs must be contiguous arrays of the plans' dtype and size")
    run_host_sweep_ptr(plans, x.ctypes.data, [o.ctypes.data for o in outs])
    return list(outs)


def moving_rms(x: np.ndarray, window: int, channels: int = 1, layout: str = "interleaved") -> np.ndarray:
    """One-shot convenience: moving RMS (sqrt of the windowed mean of squares) of a host array on the GPU."""
    return moving_average(x, window, channels, layout, op="rms")


def moving_average(x: np.ndarray, window: int, channels: int = 1, layout: str = "interleaved",
                   block_size: int = 0, path: str = "auto", op: str = "mean") -> np.ndarray:
    """One-shot convenience: moving average of a host array on the GPU (int16 or float32)."""
    x = np.asarray(x)
    if x.dtype == np.int16:
        dtype = "i16"
    elif x.dtype == np.float32:
        dtype = "f32"
    else:
        raise TypeError("libmavg filters int16 or float32 samples")
    flat = np.ascontiguousarray(x).reshape(-1)
    if flat.size % channels:
        raise ValueError("sample count must be a multiple of channels")
    if flat.size == 0:
        return flat.copy().reshape(x.shape)
    with Plan(flat.size // channels, window, channels, dtype, layout, block_size, path, op=op) as plan:
        return plan.run_host(flat).reshape(x.shape)
