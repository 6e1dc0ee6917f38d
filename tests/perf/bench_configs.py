#!/usr/bin/env python
"""Measures BASELINE.json configs 4 and 5 (and int16 variants) -- secondary to bench.py.

  config 4 : one mono float32 signal of 2^32 samples (or --log2), k = 1024, contiguous shards over N GPUs
  config 5p: 256 channels x 2^24 frames, float32, k = 64, planar [C][F], channels partitioned over N GPUs
  config 5i: same, interleaved [F][C] (the reference layout), frames partitioned over N GPUs with a halo
  i16      : stereo int16 (the reference's own input format), 2^28 samples per GPU, k sweep

Run with `python tests/perf/bench_configs.py --config 4` (1 GPU) or under torch.distributed.run for N > 1.
Every run spot-checks its output against the fp64 oracle recomputed from the generator.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import digital_signal_processsing_b200 as mavg  # noqa: E402
from digital_signal_processsing_b200 import _lib, sharding  # noqa: E402
import oracle  # noqa: E402

SEED = 0x5EED0004


class _Arr:
    def __init__(self, ptr, count, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


def alloc(nbytes):
    p = ctypes.c_void_p()
    _lib.check(_lib.load().mavg_device_alloc(nbytes, ctypes.byref(p)))
    return p


def timed(fn, stream, iters, warmup, world):
    for _ in range(warmup):
        fn()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["4", "5p", "5i", "i16", "g3", "g6i", "mci", "ci", "gen", "s2", "scan"])
    ap.add_argument("--log2", type=int, default=32, help="total samples (log2) for configs 4/5")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tune", default="", help="comma list key=value forwarded to mavg_tuning")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    lib = _lib.load()
    tune = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune.split(",") if kv}
    out = {"config": args.config, "n_gpus": world}

    if args.config == "4":
        total = 1 << args.log2
        n = total // world
        first, k = rank * n, 1024
        d_in, d_out = alloc(4 * n), alloc(4 * n)
        mavg.fill_synthetic_device(d_in.value, "f32", n, first, SEED, 0, stream.cuda_stream)
        stream.synchronize()
        plan = mavg.Plan(n, k, first_frame=first)
        plan.set_stream(stream.cuda_stream)
        plan.enable_timing(False)
        halo = int(plan.info.halo_frames)
        peer = sharding.PeerHalo(d_in.value, n, 4, halo, rank, world) if world > 1 else None
        hp = (peer.halo_ptr if peer else 0) or None
        ms = timed(lambda: plan.run_device_halo(d_in.value, d_out.value, hp), stream, args.iters, args.warmup, world)
        y = torch.as_tensor(_Arr(d_out.value, n), device="cuda")
        rng = np.random.default_rng(rank)
        idx = np.unique(np.concatenate([np.arange(0, 3 * k), rng.integers(0, n, 3000), [n - 1]]))
        got = y[torch.from_numpy(idx).cuda()].cpu().numpy()
        worst = max(abs(float(g) - oracle.point_f64(first + int(i), k, SEED)) / oracle.point_f64(first + int(i), k, SEED)
                    for g, i in zip(got, idx))
        out.update(workload=f"mono f32 2^{args.log2} samples sharded over {world} GPU(s), k=1024, halo {halo} samples via "
                            f"{'CUDA IPC peer reads' if peer else 'none'}",
                   ms=ms, gsamples_s=total / ms / 1e6, hbm_gbs_per_gpu=8 * n / ms / 1e6, max_rel_err_spot=worst,
                   spot_checks=int(idx.size))
        plan.close()
        if peer:
            peer.close()

    elif args.config in ("5p", "5i"):
        C, k = 256, 64
        F = (1 << args.log2) // C
        if args.config == "5p":
            c_per = C // world
            n = c_per * F
            first = rank * n            # planar: channel c occupies [c*F, (c+1)*F)
            d_in, d_out = alloc(4 * n), alloc(4 * n)
            mavg.fill_synthetic_device(d_in.value, "f32", n, first, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            plan = mavg.Plan(F, k, channels=c_per, layout="planar")
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, args.iters, args.warmup, world)
            y = torch.as_tensor(_Arr(d_out.value, n), device="cuda")
            rng = np.random.default_rng(rank)
            worst = 0.0
            for _ in range(300):
                c, f = int(rng.integers(0, c_per)), int(rng.integers(0, F))
                if rng.random() < 0.2:
                    f = int(rng.integers(0, 2 * k))
                lo = max(0, f - k + 1)
                x = oracle.fill_f32(f - lo + 1, SEED, 0, first_index=first + c * F + lo).astype(np.float64)
                e = x.sum() / k
                worst = max(worst, abs(float(y[c * F + f]) - e) / e)
            layout = "planar [C][F], channels partitioned"
            info = plan.info
        else:
            f_per = F // world
            n = f_per * C
            first_frame = rank * f_per
            d_in, d_out = alloc(4 * n), alloc(4 * n)
            mavg.fill_synthetic_device(d_in.value, "f32", n, first_frame * C, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            plan = mavg.Plan(f_per, k, channels=C, layout="interleaved", first_frame=first_frame, **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            halo = int(plan.info.halo_frames)
            peer = sharding.PeerHalo(d_in.value, n, 4, halo * C, rank, world) if world > 1 else None
            hp = (peer.halo_ptr if peer else 0) or None
            ms = timed(lambda: plan.run_device_halo(d_in.value, d_out.value, hp), stream, args.iters, args.warmup, world)
            y = torch.as_tensor(_Arr(d_out.value, n), device="cuda")
            rng = np.random.default_rng(rank)
            worst = 0.0
            for _ in range(300):
                c, f = int(rng.integers(0, C)), int(rng.integers(0, f_per))
                if rng.random() < 0.2:
                    f = int(rng.integers(0, 2 * k))
                gf = first_frame + f
                lo = max(0, gf - k + 1)
                x = oracle.fill_f32((gf - lo + 1) * C, SEED, 0, first_index=lo * C)[c::C].astype(np.float64)
                e = x.sum() / k
                worst = max(worst, abs(float(y[f * C + c]) - e) / e)
            layout = "interleaved [F][C], frames partitioned with halo"
            info = plan.info
            if peer:
                peer.close()
        total = C * F
        out.update(workload=f"256 ch x 2^{args.log2 - 8} frames f32, k=64, {layout}, {world} GPU(s)", ms=ms, tune=tune,
                   gsamples_s=total / ms / 1e6, hbm_gbs_per_gpu=8 * n / ms / 1e6, max_rel_err_spot=worst,
                   path="stream" if info.path == 1 else "generic")
        plan.close()

    elif args.config == "s2":  # stereo float32 through the flat streaming kernel (C = 2)
        n_frames, C = 1 << 27, 2
        n = n_frames * C
        d_in, d_out = alloc(4 * n), alloc(4 * n)
        mavg.fill_synthetic_device(d_in.value, "f32", n, 0, SEED, 0, stream.cuda_stream)
        stream.synchronize()
        res = {}
        for k in (3, 16, 64, 256, 1024, 4096):
            plan = mavg.Plan(n_frames, k, channels=C, **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 5, 2, world)
            y = torch.as_tensor(_Arr(d_out.value, n), device="cuda")
            m = 1 << 16
            x = oracle.fill_f32(m, SEED)
            e = oracle.mavg_f64(x, k, C)
            err = float(np.max(np.abs(y[:m].cpu().numpy() - e) / np.abs(e)))
            res[str(k)] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(8 * n / ms / 1e6, 1),
                           "max_rel_err_head": err, "path": "stream" if plan.info.path == 1 else "generic"}
            plan.close()
        out.update(workload="stereo float32, 2^28 samples (2^27 frames), k sweep, device resident", per_k=res)

    elif args.config == "scan":  # mavg_prefix_sum, the look-back scan primitive
        res = {}
        for dtype, C, esz, tstr in (("i16", 1, 2, "<i2"), ("i16", 2, 2, "<i2"), ("f32", 1, 4, "<f4")):
            n_frames = (1 << 28) // C
            n = n_frames * C
            d_in, d_out = alloc(esz * n), alloc(8 * n)
            mavg.fill_synthetic_device(d_in.value, dtype, n, 0, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            ms = timed(lambda: mavg.prefix_sum_device(d_in.value, d_out.value, dtype, n_frames, C, stream.cuda_stream),
                       stream, 5, 2, world)
            res[f"{dtype}_c{C}"] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1),
                                    "hbm_gbs": round((esz + 8) * n / ms / 1e6, 1)}
            lib.mavg_device_free(d_in)
            lib.mavg_device_free(d_out)
        out.update(workload="mavg_prefix_sum on 2^28 samples (bytes = input + 8-byte output per sample)", per_k=res)

    elif args.config == "gen":  # shapes that still take the generic kernel
        res = {}
        for name, dtype, C, frames, k in (("i16_c64_k64", "i16", 64, 1 << 21, 64), ("i16_c256_k64", "i16", 256, 1 << 19, 64), ("i16_c256_k8", "i16", 256, 1 << 19, 8), ("i16_c64_k1000", "i16", 64, 1 << 21, 1000), ("f32_c40_k20000", "f32", 40, 1 << 21, 20000), ("i16_c256_k8", "i16", 256, 1 << 19, 8), ("i16_c64_k1000", "i16", 64, 1 << 21, 1000),
                                          ("f32_c34_k64", "f32", 34, 1 << 22, 64), ("f32_c64_k2048", "f32", 64, 1 << 21, 2048),
                                          ("f32_c1_k60000", "f32", 1, 1 << 27, 60000), ("f32_c1_k300000", "f32", 1, 1 << 27, 300000), ("f32_c2_k30000", "f32", 2, 1 << 26, 30000), ("f32_c1_k4096_tail", "f32", 1, (1 << 27) + 5, 4096), ("f32_c1_k4096", "f32", 1, 1 << 27, 4096), ("i16_c2_k40000", "i16", 2, 1 << 26, 40000), ("i16_c1_k46000", "i16", 1, 1 << 27, 46000),
                                          ("i16_c8_k19200", "i16", 8, 1 << 24, 19200), ("i16_c6_k19200", "i16", 6, (1 << 27) // 6, 19200),
                                          ("i16_c4_k19200", "i16", 4, 1 << 25, 19200), ("i16_c2_k50000", "i16", 2, 1 << 26, 50000)):
            es = 4 if dtype == "f32" else 2
            n = frames * C
            d_in, d_out = alloc(es * n), alloc(es * n)
            mavg.fill_synthetic_device(d_in.value, dtype, n, 0, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            plan = mavg.Plan(frames, k, channels=C, dtype=dtype, **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 3, 1, world)
            res[name] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(2 * es * n / ms / 1e6, 1),
                         "path": "stream" if plan.info.path == 1 else "generic", "mode": int(plan.info.mode),
                         "launches": int(plan.info.launches_per_run)}
            plan.close()
            lib.mavg_device_free(d_in)
            lib.mavg_device_free(d_out)
        out.update(workload="shapes served by the generic kernel (2^27 samples each)", per_k=res)

    elif args.config == "g3":  # shapes only the generic kernel takes: 3-channel interleaved float32
        n_frames, C = 1 << 25, 3
        n = n_frames * C
        d_in, d_out = alloc(4 * n), alloc(4 * n)
        mavg.fill_synthetic_device(d_in.value, "f32", n, 0, SEED, 0, stream.cuda_stream)
        stream.synchronize()
        res = {}
        for k in (3, 32, 64, 128, 256, 1024, 4096):
            plan = mavg.Plan(n_frames, k, channels=C, **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 3, 1, world)
            res[str(k)] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(8 * n / ms / 1e6, 1),
                           "path": "stream" if plan.info.path == 1 else "generic", "launches": int(plan.info.launches_per_run)}
            plan.close()
        out.update(workload="3-channel interleaved float32, 3 x 2^25 samples, k sweep (few-channel kernel up to k=256, "
                            "generic kernel above)", per_k=res)

    elif args.config == "g6i":  # 5.1 PCM audio: six interleaved int16 channels (few-channel int16 kernel)
        n_frames, C = 1 << 25, 6
        n = n_frames * C
        d_in, d_out = alloc(2 * n), alloc(2 * n)
        mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, stream.cuda_stream)
        stream.synchronize()
        res = {}
        for k in (3, 32, 64, 128, 256, 1024):
            plan = mavg.Plan(n_frames, k, channels=C, dtype="i16", **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 3, 1, world)
            y = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda")
            m = 6 << 14
            ok = bool(np.array_equal(y[:m].cpu().numpy(), oracle.mavg_i16(oracle.fill_i16(m, SEED), k, C)))
            res[str(k)] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(4 * n / ms / 1e6, 1),
                           "bit_exact_head": ok, "path": "stream" if plan.info.path == 1 else "generic",
                           "launches": int(plan.info.launches_per_run)}
            plan.close()
        out.update(workload="6-channel interleaved int16, 6 x 2^25 samples, k sweep", per_k=res)

    elif args.config == "ci":  # many-channel interleaved int16 (column kernel), 2^27 samples
        res = {}
        for C, k in ((64, 8), (64, 64), (64, 1000), (128, 64), (256, 8), (256, 64), (1024, 64), (72, 64), (32, 64)):
            n_frames = (1 << 27) // C
            n = n_frames * C
            d_in, d_out = alloc(2 * n), alloc(2 * n)
            mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            plan = mavg.Plan(n_frames, k, channels=C, dtype="i16", **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 5, 2, world)
            y = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda")
            m = C * 3000
            ok = bool(np.array_equal(y[:m].cpu().numpy(), oracle.mavg_i16(oracle.fill_i16(m, SEED), k, C)))
            res[f"c{C}_k{k}"] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(4 * n / ms / 1e6, 1),
                                 "bit_exact_head": ok, "mode": int(plan.info.mode), "threads": int(plan.info.threads),
                                 "path": "stream" if plan.info.path == 1 else "generic"}
            plan.close()
            lib.mavg_device_free(d_in)
            lib.mavg_device_free(d_out)
        out.update(workload="C-channel interleaved int16 (C >= 32), 2^27 samples", per_k=res)

    elif args.config == "mci":  # multichannel PCM: 3 / 4 / 6 / 8 (flat-stream kernel) and 5 / 12 (few-channel kernels) int16 channels
        res = {}
        for C in (3, 4, 6, 8, 5, 7, 12, 16, 10, 9):
            n_frames = (1 << 27) // C
            n = n_frames * C
            d_in, d_out = alloc(2 * n), alloc(2 * n)
            mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, stream.cuda_stream)
            stream.synchronize()
            for k in (3, 64, 1024, 4096):
                plan = mavg.Plan(n_frames, k, channels=C, dtype="i16", **tune)
                plan.set_stream(stream.cuda_stream)
                plan.enable_timing(False)
                ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 5, 2, world)
                y = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda")
                m = C << 15
                ok = bool(np.array_equal(y[:m].cpu().numpy(), oracle.mavg_i16(oracle.fill_i16(m, SEED), k, C)))
                res[f"c{C}_k{k}"] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(4 * n / ms / 1e6, 1),
                                     "bit_exact_head": ok, "mode": int(plan.info.mode),
                                     "path": "stream" if plan.info.path == 1 else "generic"}
                plan.close()
            lib.mavg_device_free(d_in)
            lib.mavg_device_free(d_out)
        out.update(workload="C-channel interleaved int16, 2^27 samples, k sweep", per_k=res)

    else:  # i16: the reference's own input format
        n_frames, C = 1 << 27, 2
        n = n_frames * C
        d_in, d_out = alloc(2 * n), alloc(2 * n)
        mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, stream.cuda_stream)
        stream.synchronize()
        res = {}
        for k in (3, 16, 64, 256, 1024, 4096):
            plan = mavg.Plan(n_frames, k, channels=C, dtype="i16", **tune)
            plan.set_stream(stream.cuda_stream)
            plan.enable_timing(False)
            ms = timed(lambda: plan.run_device([d_in.value], [d_out.value]), stream, 3, 1, world)
            y = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda")
            m = 1 << 16
            x = oracle.fill_i16(m, SEED)
            ok = bool(np.array_equal(y[:m].cpu().numpy(), oracle.mavg_i16(x, k, C)))
            res[str(k)] = {"ms": round(ms, 4), "gsamples_s": round(n / ms / 1e6, 1), "hbm_gbs": round(4 * n / ms / 1e6, 1),
                           "bit_exact_head": ok, "path": "stream" if plan.info.path == 1 else "generic"}
            plan.close()
        out.update(workload="stereo int16, 2^28 samples, k sweep, device resident", per_k=res, tune=tune)

    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
