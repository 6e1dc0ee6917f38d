#!/usr/bin/env python
"""Does the distance between the input and the output buffer matter?  One allocation, the output placed at
input + size + delta for a few deltas; kernel time per shape (CUDA events, 10 launches, best of 3).
  python tests/perf/alias_probe.py"""
import json
import os
import sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
import digital_signal_processsing_b200 as mavg

out = {}
for dtype, C, k in (("i16", 4, 1024), ("i16", 8, 64), ("i16", 2, 64), ("f32", 1, 64)):
    es = 2 if dtype == "i16" else 4
    n = 1 << 27
    frames = n // C
    nbytes = n * es
    big = torch.empty(2 * nbytes + (64 << 20), dtype=torch.uint8, device="cuda")
    base = (big.data_ptr() + 1023) // 1024 * 1024
    mavg.fill_synthetic_device(base, dtype, n, 0, 11)
    torch.cuda.synchronize()
    res = {}
    st = torch.cuda.Stream()
    with mavg.Plan(frames, k, channels=C, dtype=dtype) as plan:
        plan.set_stream(st.cuda_stream)
        plan.enable_timing(False)
        for delta in (0, 1024, 4096, 16384, 65536, 1 << 20, (1 << 20) + 4096, 2 << 20, 3 << 20, 8 << 20, 16 << 20, (32 << 20) + 8192):
            outp = base + nbytes + delta
            best = 1e9
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(3):
                    plan.run_device([base], [outp])
                st.synchronize()
                a.record(st)
                for _ in range(10):
                    plan.run_device([base], [outp])
                b.record(st)
                b.synchronize()
                best = min(best, a.elapsed_time(b) / 10)
            res[str(delta)] = round(best, 4)
    out[f"{dtype}_c{C}_k{k}"] = res
    del big
print(json.dumps(out))
