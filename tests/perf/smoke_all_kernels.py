#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, a quick correctness smoke of each path (compute-sanitizer is closed on this pool)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import numpy as np
import torch
import digital_signal_processsing_b200 as mavg
import oracle

def rel(y, e):
    return float(np.max(np.abs(y - e) / np.abs(e)))

n = 3 * 8192 + 100
x = oracle.fill_f32(n, 1)
for k in (3, 64, 300, 4096, 9000):
    assert rel(mavg.moving_average(x, k), oracle.mavg_f64(x, k)) < 1e-5, k
xs = oracle.fill_f32(2 * n, 2)
for k in (5, 100, 2000):
    assert rel(mavg.moving_average(xs, k, channels=2), oracle.mavg_f64(xs, k, 2)) < 1e-5, k
xi = oracle.fill_i16(2 * (2 * 16384 + 77), 3)
for k in (3, 41, 1000, 4096):
    assert np.array_equal(mavg.moving_average(xi, k, channels=2), oracle.mavg_i16(xi, k, 2)), k
xc = oracle.fill_f32(64 * (3 * 256 + 9), 4)
for k in (16, 300, 1000):
    assert rel(mavg.moving_average(xc, k, channels=64), oracle.mavg_f64(xc, k, 64)) < 1e-5, k
xg = oracle.fill_f32(3 * 20001, 5)
for k in (4, 700):
    assert rel(mavg.moving_average(xg, k, channels=3), oracle.mavg_f64(xg, k, 3)) < 1e-5, k
xp = oracle.fill_f32(4 * 8192 * 2, 6)
yp = mavg.moving_average(xp, 64, channels=4, layout="planar")
assert rel(yp[:16384], oracle.mavg_f64(xp[:16384], 64)) < 1e-5
x6 = oracle.fill_i16(6 * 9001, 8)                                 # few-channel int16: channel pairs (6), scalar (5)
for k in (3, 100, 1000):
    assert np.array_equal(mavg.moving_average(x6, k, channels=6), oracle.mavg_i16(x6, k, 6)), k
x5 = oracle.fill_i16(5 * 9001, 9)
assert np.array_equal(mavg.moving_average(x5, 48, channels=5), oracle.mavg_i16(x5, 48, 5))
x64 = oracle.fill_i16(64 * 3001, 10)                              # int16 column kernel
assert np.array_equal(mavg.moving_average(x64, 64, channels=64), oracle.mavg_i16(x64, 64, 64))
xf = oracle.fill_f32(12 * 8192 + 45, 11)                          # far-lag kernel (mono, stereo) + tail kernel
assert rel(mavg.moving_average(xf, 60001), oracle.mavg_f64(xf, 60001)) < 1e-5
xf2 = oracle.fill_f32(2 * (12 * 4096 + 21), 12)
assert rel(mavg.moving_average(xf2, 30001, channels=2), oracle.mavg_f64(xf2, 30001, 2)) < 1e-5
for ch in (3, 4, 5, 7, 8, 9, 10, 12, 16):                         # flat-stream int16 kernel, every channel count it takes
    xm = oracle.fill_i16(ch * 40001, 20 + ch)
    for k in (3, 500):
        assert np.array_equal(mavg.moving_average(xm, k, channels=ch), oracle.mavg_i16(xm, k, ch)), (ch, k)
for ch, k in ((11, 64), (14, 64), (24, 300)):                     # the older few-channel int16 kernels
    xm = oracle.fill_i16(ch * 20001, 40 + ch)
    assert np.array_equal(mavg.moving_average(xm, k, channels=ch), oracle.mavg_i16(xm, k, ch)), (ch, k)
for ch, k in ((2, 30001), (1, 46340), (8, 7000), (6, 9001), (16, 3500)):   # far-lag int16 kernel
    xm = oracle.fill_i16(ch * (900_000 // ch + 7), 60 + ch)
    assert np.array_equal(mavg.moving_average(xm, k, channels=ch), oracle.mavg_i16(xm, k, ch)), (ch, k)
x256 = oracle.fill_i16(256 * 2001, 70)                            # int16 column kernel, four column warps per row
assert np.array_equal(mavg.moving_average(x256, 8, channels=256), oracle.mavg_i16(x256, 8, 256))
d = torch.from_numpy(oracle.fill_i16(2 * 50001, 7)).cuda()
o = torch.zeros(2 * 50001, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
mavg.prefix_sum_device(d.data_ptr(), o.data_ptr(), "i16", 50001, 2)
torch.cuda.synchronize()
assert torch.equal(o, torch.cumsum(d.view(-1, 2).to(torch.int64), 0).view(-1))
print("smoke_all_kernels: all kernel families ok")
