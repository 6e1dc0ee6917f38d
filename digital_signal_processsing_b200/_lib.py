"""ctypes binding of libmavg.so (the C ABI in include/mavg.h).

There is no Python or CPU implementation behind this module: if the CUDA library is
missing or cannot be loaded, importing fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmavg.so")

MAVG_MAX_DEVICES = 16

# enums (include/mavg.h)
F32, I16 = 0, 1
INTERLEAVED, PLANAR = 0, 1
PATH_AUTO, PATH_STREAM, PATH_GENERIC = 0, 1, 2
DIST_U01, DIST_USYM, DIST_I16, DIST_DC1E4 = 0, 1, 2, 3
OP_MEAN, OP_RMS = 0, 1

OK = 0
ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC, ERR_BLOCK_SIZE, ERR_DRIVER = -1, -2, -3, -4, -5, -6, -7


class Tuning(ctypes.Structure):
    _fields_ = [
        ("threads", ctypes.c_uint32),
        ("run", ctypes.c_uint32),
        ("prefetch", ctypes.c_uint32),
        ("ctas_per_sm", ctypes.c_uint32),
        ("chunks_per_cta", ctypes.c_uint32),
        ("direct_max_k", ctypes.c_uint32),
        ("slice_bytes", ctypes.c_uint32),
        ("overlap", ctypes.c_uint32),
    ]


class Desc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("dtype", ctypes.c_uint32),
        ("layout", ctypes.c_uint32),
        ("channels", ctypes.c_uint32),
        ("frames", ctypes.c_uint64),
        ("window", ctypes.c_uint32),
        ("block_size", ctypes.c_uint32),
        ("path", ctypes.c_uint32),
        ("num_devices", ctypes.c_uint32),
        ("devices", ctypes.c_int32 * MAVG_MAX_DEVICES),
        ("first_frame", ctypes.c_uint64),
        ("tuning", Tuning),
        ("op", ctypes.c_uint32),
        ("reserved", ctypes.c_uint32),
    ]


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("compute_ms", ctypes.c_float),
                ("d2h_ms", ctypes.c_float), ("total_ms", ctypes.c_float)]


class Info(ctypes.Structure):
    _fields_ = [
        ("path", ctypes.c_uint32), ("mode", ctypes.c_uint32),
        ("threads", ctypes.c_uint32), ("run", ctypes.c_uint32),
        ("tile_samples", ctypes.c_uint32), ("history_tiles", ctypes.c_uint32),
        ("stages", ctypes.c_uint32), ("grid", ctypes.c_uint32),
        ("smem_bytes", ctypes.c_uint32), ("launches_per_run", ctypes.c_uint32),
        ("num_devices", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
        ("halo_frames", ctypes.c_uint64),
        ("shard_frames", ctypes.c_uint64 * MAVG_MAX_DEVICES),
    ]


# every symbol include/mavg.h declares: (name, restype, argtypes)
_vp, _i, _u32, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
SYMBOLS = [
    ("mavg_version", _i, []),
    ("mavg_strerror", ctypes.c_char_p, [_i]),
    ("mavg_last_error", ctypes.c_char_p, []),
    ("mavg_device_count", _i, []),
    ("mavg_plan_create", _i, [ctypes.POINTER(Desc), ctypes.POINTER(_vp)]),
    ("mavg_plan_destroy", _i, [_vp]),
    ("mavg_plan_info", _i, [_vp, ctypes.POINTER(Info)]),
    ("mavg_run_host", _i, [_vp, _vp, _vp]),
    ("mavg_run_host_sweep", _i, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    ("mavg_run_device", _i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("mavg_run_cascade", _i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.c_uint32]),
    ("mavg_run_device_halo", _i, [_vp, _vp, _vp, _vp]),
    ("mavg_synchronize", _i, [_vp]),
    ("mavg_get_timing", _i, [_vp, ctypes.POINTER(Timing)]),
    ("mavg_enable_timing", _i, [_vp, _i]),
    ("mavg_set_stream", _i, [_vp, _vp]),
    ("mavg_plan_buffers", _i, [_vp, _u32, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    ("mavg_fill_synthetic", _i, [_vp, _u64, _i]),
    ("mavg_fill_synthetic_device", _i, [_vp, _i, _u64, _u64, _u64, _i, _vp]),
    ("mavg_run_owned", _i, [_vp]),
    ("mavg_prefix_sum", _i, [_i, _vp, _vp, _u64, _u32, _vp]),
    ("mavg_ipc_export", _i, [_vp, _vp]),
    ("mavg_ipc_open", _i, [_vp, ctypes.POINTER(_vp)]),
    ("mavg_ipc_close", _i, [_vp]),
    ("mavg_device_alloc", _i, [_u64, ctypes.POINTER(_vp)]),
    ("mavg_device_free", _i, [_vp]),
    ("mavg_host_alloc", _i, [_u64, ctypes.POINTER(_vp)]),
    ("mavg_host_free", _i, [_vp]),
    ("mavg_host_register", _i, [_vp, _u64]),
    ("mavg_host_unregister", _i, [_vp]),
]


class MavgError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str):
        super().__init__(f"{what}: {detail}" if detail else what)
        self.status = status


_lib = None


def load() -> ctypes.CDLL:
    """Loads libmavg.so; raises (never falls back) when the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m digital_signal_processsing_b200.build` "
            "(nvcc, sm_100a). libmavg has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != OK:
        lib = load()
        what = lib.mavg_strerror(status).decode()
        detail = lib.mavg_last_error().decode()
        raise MavgError(status, what, detail)
