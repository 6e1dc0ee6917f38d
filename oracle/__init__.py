"""CPU oracle bindings (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this package; the product (digital_signal_processsing_b200, libmavg)
never does.  See oracle/mavg_oracle.h for what is restated and how it is pinned.

Reference followed: basics/profilable_moving_averager.cpp:14-37.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmavg_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")

DIST_U01, DIST_USYM, DIST_I16, DIST_DC1E4 = 0, 1, 2, 3


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference exists, oracle/_ref."""
    targets = ["all"] + (["ref", "dropin"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _load():
    if not os.path.exists(_LIB_PATH):
        build(ref=False)
    lib = ctypes.CDLL(_LIB_PATH)
    u64, u32, i32, vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p
    lib.oracle_mavg_i16.argtypes = [vp, vp, u64, u32, u32]
    lib.oracle_mavg_i16.restype = None
    lib.oracle_mavg_f32_to_f64.argtypes = [vp, vp, u64, u32, u32]
    lib.oracle_mavg_f32_to_f64.restype = None
    lib.oracle_mrms_f32_to_f64.argtypes = [vp, vp, u64, u32, u32]
    lib.oracle_mrms_f32_to_f64.restype = None
    lib.oracle_mrms_i16.argtypes = [vp, vp, u64, u32, u32]
    lib.oracle_mrms_i16.restype = None
    lib.oracle_mavg_f32_running.argtypes = [vp, vp, u64, u32, u32]
    lib.oracle_mavg_f32_running.restype = None
    lib.oracle_mavg_f32_running_mt.argtypes = [vp, vp, u64, u32, u32, i32]
    lib.oracle_mavg_f32_running_mt.restype = i32
    lib.oracle_mavg_i16_mt.argtypes = [vp, vp, u64, u32, u32, i32]
    lib.oracle_mavg_i16_mt.restype = i32
    lib.oracle_mix64.argtypes = [u64, u64]
    lib.oracle_mix64.restype = u64
    lib.oracle_fill_f32.argtypes = [vp, u64, u64, u64, i32]
    lib.oracle_fill_f32.restype = None
    lib.oracle_fill_i16.argtypes = [vp, u64, u64, u64]
    lib.oracle_fill_i16.restype = None
    lib.oracle_point_f64.argtypes = [u64, u32, u64, i32]
    lib.oracle_point_f64.restype = ctypes.c_double
    lib.oracle_time_best.argtypes = [i32, vp, vp, u64, u32, u32, i32, i32]
    lib.oracle_time_best.restype = ctypes.c_double
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def _frames(x: np.ndarray, channels: int) -> int:
    if x.ndim != 1 or x.size % channels:
        raise ValueError("expected a flat interleaved array whose size divides by channels")
    return x.size // channels


def mavg_i16(x: np.ndarray, k: int, channels: int = 1) -> np.ndarray:
    """Bit-exact a1 (int16, int64 sum, truncating division)."""
    x = np.ascontiguousarray(x, dtype=np.int16)
    y = np.empty_like(x)
    lib().oracle_mavg_i16(_ptr(x), _ptr(y), _frames(x, channels), channels, k)
    return y


def mavg_f64(x: np.ndarray, k: int, channels: int = 1) -> np.ndarray:
    """fp64 evaluation of the a1 definition lifted to reals, for fp32 input."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(x.shape, dtype=np.float64)
    lib().oracle_mavg_f32_to_f64(_ptr(x), _ptr(y), _frames(x, channels), channels, k)
    return y


def mrms_f64(x: np.ndarray, k: int, channels: int = 1) -> np.ndarray:
    """Moving RMS of fp32 input evaluated in fp64: sqrt(mean of squares over the causal, zero-padded window)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty(x.shape, dtype=np.float64)
    lib().oracle_mrms_f32_to_f64(_ptr(x), _ptr(y), _frames(x, channels), channels, k)
    return y


def mrms_i16(x: np.ndarray, k: int, channels: int = 1) -> np.ndarray:
    """Moving RMS of int16 input: exact int64 sum of squares, (int16) trunc(sqrt((double) sum / k))."""
    x = np.ascontiguousarray(x, dtype=np.int16)
    y = np.empty_like(x)
    lib().oracle_mrms_i16(_ptr(x), _ptr(y), _frames(x, channels), channels, k)
    return y


def mavg_f32_running(x: np.ndarray, k: int, channels: int = 1, threads: int = 1) -> np.ndarray:
    """fp32 port of the reference running-sum loop (CPU baseline, not an accuracy oracle)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    if threads <= 1:
        lib().oracle_mavg_f32_running(_ptr(x), _ptr(y), _frames(x, channels), channels, k)
    else:
        lib().oracle_mavg_f32_running_mt(_ptr(x), _ptr(y), _frames(x, channels), channels, k, threads)
    return y


def mavg_i16_mt(x: np.ndarray, k: int, channels: int, threads: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int16)
    y = np.empty_like(x)
    lib().oracle_mavg_i16_mt(_ptr(x), _ptr(y), _frames(x, channels), channels, k, threads)
    return y


def fill_f32(n: int, seed: int, dist: int = DIST_U01, first_index: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lib().oracle_fill_f32(_ptr(out), n, first_index, seed, dist)
    return out


def fill_i16(n: int, seed: int, first_index: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.int16)
    lib().oracle_fill_i16(_ptr(out), n, first_index, seed)
    return out


def point_f64(i: int, k: int, seed: int, dist: int = DIST_U01) -> float:
    return float(lib().oracle_point_f64(i, k, seed, dist))


def time_best(which: int, x: np.ndarray, k: int, channels: int = 1, threads: int = 1, iters: int = 3) -> float:
    """Best wall-clock seconds; which: 0 i16 1T, 1 f32 1T, 2 f32 MT, 3 i16 MT."""
    y = np.empty_like(x)
    return float(lib().oracle_time_best(which, _ptr(x), _ptr(y), _frames(x, channels), channels, k, threads, iters))


# ------------------------------------------------------------- reference (_ref)

_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF_PATH)


def ref():
    """The UNMODIFIED reference function compiled from /root/reference (oracle/_ref)."""
    global _ref
    if _ref is None:
        r = ctypes.CDLL(_REF_PATH)
        r.ref_cpu_i16.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        r.ref_cpu_i16.restype = ctypes.c_int
        r.ref_cpu_i16_time.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
        r.ref_cpu_i16_time.restype = ctypes.c_double
        _ref = r
    return _ref


def ref_mavg_i16(x: np.ndarray, k: int, channels: int = 1) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int16)
    y = np.empty_like(x)
    rc = ref().ref_cpu_i16(channels, k, _ptr(x), _ptr(y), x.size)
    if rc != 0:
        raise ValueError("reference precondition frames >= k violated (its warm-up loop is unguarded)")
    return y


def ref_time_i16(x: np.ndarray, k: int, channels: int = 1, iters: int = 3) -> float:
    x = np.ascontiguousarray(x, dtype=np.int16)
    return float(ref().ref_cpu_i16_time(channels, k, _ptr(x), None, x.size, iters))
