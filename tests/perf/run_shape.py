#!/usr/bin/env python
"""Runs one plan shape a few times on device-resident synthetic data (the command ncu captures are taken from).
  python tests/perf/run_shape.py i16 256 524288 64 [reps] [key=value tuning ...]"""
import os
import sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
import digital_signal_processsing_b200 as mavg

dtype, C, frames, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
tune = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[6:]}
n = frames * C
tdt = torch.int16 if dtype == "i16" else torch.float32
x = torch.empty(n, dtype=tdt, device="cuda")
y = torch.empty(n, dtype=tdt, device="cuda")
mavg.fill_synthetic_device(x.data_ptr(), dtype, n, 0, 11)
torch.cuda.synchronize()
with mavg.Plan(frames, k, channels=C, dtype=dtype, **tune) as plan:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.run_device([x.data_ptr()], [y.data_ptr()])
    plan.synchronize()
    a.record()
    for _ in range(reps):
        plan.run_device([x.data_ptr()], [y.data_ptr()])
        plan.synchronize()
    b.record()
    b.synchronize()
    i = plan.info
    print(f"{dtype} C={C} frames={frames} k={k}: path={i.path} mode={i.mode} {a.elapsed_time(b) / reps:.4f} ms/run (host-synchronised)")
