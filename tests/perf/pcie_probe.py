"""Measurement helper: what the host link gives for pinned 1 GiB buffers (H2D alone, D2H alone, both at once)
next to mavg_run_host on the same buffers -- the ceiling bench.py's `e2e` figure can reach on this box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch

import digital_signal_processsing_b200 as mavg


def main():
    n = 1 << 28
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_in.uniform_()
    d_a = torch.empty(n, dtype=torch.float32, device="cuda")
    d_b = torch.ones(n, dtype=torch.float32, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    gb = 4 * n / 1e9

    def wall(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {"bytes_each_way": 4 * n}
    out["h2d_gbs"] = round(gb / wall(h2d), 2)
    out["d2h_gbs"] = round(gb / wall(d2h), 2)
    out["both_each_way_gbs"] = round(gb / wall(both), 2)
    for k in (64,):
        for slice_mib in (2, 4, 8, 16, 32, 64):
            with mavg.Plan(n, k, slice_bytes=slice_mib << 20) as plan:
                t = wall(lambda: plan.run_host_ptr(h_in.data_ptr(), h_out.data_ptr()))
                out[f"run_host_k{k}_slice{slice_mib}MiB"] = {"gsamples_s": round(n / t / 1e9, 2),
                                                             "each_way_gbs": round(gb / t, 2)}
    with mavg.Plan(n, 64) as plan:                         # library default slice size
        t = wall(lambda: plan.run_host_ptr(h_in.data_ptr(), h_out.data_ptr()))
        out["run_host_k64_default"] = {"gsamples_s": round(n / t / 1e9, 2), "each_way_gbs": round(gb / t, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
