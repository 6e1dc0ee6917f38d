"""Measurement helper: mavg_run_host on pageable (numpy / std::vector-like) host buffers against pinned ones."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np

import digital_signal_processsing_b200 as mavg


def best(fn, reps=3):
    fn()
    t = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t = min(t, time.perf_counter() - t0)
    return t


def main():
    out = {}
    for log2 in (22, 26, 28):
        n = 1 << log2
        x = np.random.default_rng(1).random(n, dtype=np.float32)
        y = np.empty_like(x)
        with mavg.Plan(n, 64) as plan, mavg.PinnedArray(n, np.float32) as pin, mavg.PinnedArray(n, np.float32) as pout:
            pin.array[:] = x
            tp = best(lambda: plan.run_host(x, out=y))
            tq = best(lambda: plan.run_host(pin.array, out=pout.array))
            assert np.array_equal(y, pout.array)
            t0 = time.perf_counter()
            with mavg.pinned(x, y):
                t_reg = time.perf_counter() - t0
                tr = best(lambda: plan.run_host(x, out=y))
            assert np.array_equal(y, pout.array)
            out[f"2^{log2}"] = {"pageable_ms": round(tp * 1e3, 3), "pinned_ms": round(tq * 1e3, 3),
                                "registered_ms": round(tr * 1e3, 3), "register_both_ms": round(t_reg * 1e3, 3),
                                "pageable_gsamples_s": round(n / tp / 1e9, 2), "pinned_gsamples_s": round(n / tq / 1e9, 2),
                                "registered_gsamples_s": round(n / tr / 1e9, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
