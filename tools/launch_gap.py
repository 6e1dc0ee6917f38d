#!/usr/bin/env python
"""Where do the ~10 us per launch outside the kernel go?  Times N back-to-back launches of one plan
between two events (no per-launch events), per-launch events, and a CUDA-graph replay."""
import ctypes, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import digital_signal_processsing_b200 as mavg
from digital_signal_processsing_b200 import _lib

n = 1 << 28
lib = _lib.load()
d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
_lib.check(lib.mavg_device_alloc(4 * n, ctypes.byref(d_in)))
_lib.check(lib.mavg_device_alloc(4 * n, ctypes.byref(d_out)))
stream = torch.cuda.Stream()
mavg.fill_synthetic_device(d_in.value, "f32", n, 0, 1, 0, stream.cuda_stream)
stream.synchronize()
for k in (3, 1024):
    plan = mavg.Plan(n, k)
    plan.set_stream(stream.cuda_stream)
    plan.enable_timing(False)
    run = lambda: plan.run_device_halo(d_in.value, d_out.value, None)
    for _ in range(5):
        run()
    stream.synchronize()
    N = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(N):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    a = e0.elapsed_time(e1) / N
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    evs[0].record(stream)
    for i in range(N):
        run()
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    b = sum(evs[i].elapsed_time(evs[i + 1]) for i in range(N)) / N
    c = float("nan")
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(10):
                run()
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); g.replay(); g.replay(); e1.record()
        torch.cuda.synchronize()
        c = e0.elapsed_time(e1) / 30
    except Exception as ex:
        print("graph capture failed:", ex)
    print(f"k={k}: back-to-back {a:.4f} ms/launch, per-launch events {b:.4f}, graph replay {c:.4f}; "
          f"GB/s {8*n/a/1e6:.0f} / {8*n/b/1e6:.0f} / {8*n/c/1e6:.0f}")
    plan.close()
