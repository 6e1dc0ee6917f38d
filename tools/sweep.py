#!/usr/bin/env python
"""Tuning sweep of the streaming kernel on one GPU (kernel-only, CUDA-event timing inside libmavg).

  python tools/sweep.py [--log2 28] [--ks 3,64,1024,4096] [--iters 10] [--out gpurun_out/sweep.csv]

Each configuration: plan-owned device buffers, synthetic U[0,1) input, 3 warm-up runs, `iters`
timed runs; reports the median / min compute_ms, Gsamples/s and GB/s (8 B per sample).
"""
import argparse
import itertools
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import digital_signal_processsing_b200 as mavg  # noqa: E402

CONFIGS = [
    dict(threads=256, run=16, prefetch=2, ctas_per_sm=2),
    dict(threads=256, run=16, prefetch=3, ctas_per_sm=1),
    dict(threads=256, run=16, prefetch=6, ctas_per_sm=1),
    dict(threads=256, run=16, prefetch=1, ctas_per_sm=2),
    dict(threads=256, run=16, prefetch=1, ctas_per_sm=3),
    dict(threads=256, run=32, prefetch=2, ctas_per_sm=1),
    dict(threads=256, run=32, prefetch=3, ctas_per_sm=1),
    dict(threads=512, run=16, prefetch=2, ctas_per_sm=1),
    dict(threads=512, run=16, prefetch=3, ctas_per_sm=1),
    dict(threads=256, run=16, prefetch=2, ctas_per_sm=2, chunks_per_cta=4),
    dict(threads=256, run=16, prefetch=2, ctas_per_sm=2, chunks_per_cta=16),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=28)
    ap.add_argument("--ks", default="3,64,1024,4096")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default="")
    ap.add_argument("--configs", default="", help="JSON list of tuning dicts overriding the built-in list")
    args = ap.parse_args()
    n = 1 << args.log2
    ks = [int(v) for v in args.ks.split(",")]
    configs = json.loads(args.configs) if args.configs else CONFIGS
    rows = []
    hdr = "threads,run,prefetch,ctas_per_sm,chunks_per_cta,k,mode,stages,smem,grid,median_ms,min_ms,gsamples_s,gbs"
    print(hdr)
    for cfg, k in itertools.product(configs, ks):
        try:
            with mavg.Plan(n, k, path="stream", **cfg) as plan:
                plan.fill_synthetic(1, 0)
                ms = []
                for i in range(3 + args.iters):
                    plan.run_owned()
                    plan.synchronize()
                    if i >= 3:
                        ms.append(plan.timing().compute_ms)
                info = plan.info
                med, mn = statistics.median(ms), min(ms)
                row = (f"{cfg.get('threads')},{cfg.get('run')},{cfg.get('prefetch')},{cfg.get('ctas_per_sm')},"
                       f"{cfg.get('chunks_per_cta', 0)},{k},{info.mode},{info.stages},{info.smem_bytes},{info.grid},"
                       f"{med:.4f},{mn:.4f},{n / med / 1e6:.1f},{8 * n / med / 1e6:.1f}")
        except mavg.MavgError as e:
            row = f"{cfg},{k},ERROR {e}"
        print(row, flush=True)
        rows.append(row)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(hdr + "\n" + "\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
