#!/usr/bin/env python
"""How fast can this GPU WRITE?  The prefix-sum primitive's traffic is 80 % writes (2 B in, 8 B out per int16
sample), so its ceiling is the write-side bandwidth, not the copy figure of MEASURED_PEAKS.json.
Probes (2 GiB each, CUDA events, best of 5): cudaMemsetAsync (torch.zero_), torch.fill_ (vectorised stores),
torch copy (read + write), and a read-only reduction for symmetry."""
import json
import torch

n = 1 << 28  # int64 elements = 2 GiB
x = torch.empty(n, dtype=torch.int64, device="cuda")
y = torch.empty(n, dtype=torch.int64, device="cuda")


def best(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


out = {}
ms = best(lambda: x.zero_())
out["memset_zero"] = {"ms": ms, "write_gbs": 8 * n / ms / 1e6}
ms = best(lambda: x.fill_(7))
out["fill"] = {"ms": ms, "write_gbs": 8 * n / ms / 1e6}
ms = best(lambda: y.copy_(x))
out["copy"] = {"ms": ms, "read_plus_write_gbs": 16 * n / ms / 1e6}
ms = best(lambda: x.sum())
out["read_sum"] = {"ms": ms, "read_gbs": 8 * n / ms / 1e6}
# 20 % reads + 80 % writes, like the scan: int16 -> int64 conversion
s16 = torch.empty(n, dtype=torch.int16, device="cuda")
ms = best(lambda: x.copy_(s16))
out["convert_i16_to_i64"] = {"ms": ms, "total_gbs": 10 * n / ms / 1e6}
print(json.dumps(out))
