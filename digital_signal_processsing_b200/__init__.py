"""digital_signal_processsing_b200 -- host-side Python face of libmavg.

The product is the CUDA library (csrc/, C ABI in include/mavg.h); this package is the thin
ctypes layer tests, bench.py and run_benchmarks.py use.  Importing the package does not
load the library; the first Plan does, and it raises if libmavg.so is not built -- there is
no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import (DIST_DC1E4, DIST_I16, DIST_U01, DIST_USYM, MavgError)  # noqa: F401
from .plan import (PinnedArray, Plan, device_count, fill_synthetic_device, moving_average, moving_rms, pinned,  # noqa: F401
                   prefix_sum_device, run_host_sweep, run_host_sweep_ptr, version)
from . import wav  # noqa: F401
