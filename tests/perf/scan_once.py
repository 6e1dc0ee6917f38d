#!/usr/bin/env python
"""One mavg_prefix_sum call per dtype on 2^28 samples (the command the ncu captures of the primitive run)."""
import ctypes
import sys
import os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
import digital_signal_processsing_b200 as mavg

n = 1 << 28
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for dtype, tdt in (("i16", torch.int16), ("f32", torch.float32)):
    x = torch.zeros(n, dtype=tdt, device="cuda")
    mavg.fill_synthetic_device(x.data_ptr(), dtype, n, 0, 7)
    y = torch.empty(n, dtype=torch.int64 if dtype == "i16" else torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(reps):
        mavg.prefix_sum_device(x.data_ptr(), y.data_ptr(), dtype, n, 1)
    torch.cuda.synchronize()
    del x, y
print("done")
