#!/usr/bin/env python
"""bench.py -- headline benchmark of libmavg (BASELINE.json metric).

Metric: Gsamples/s of the moving-average hot path on a 2^28-sample mono float32 synthetic
signal per GPU, window sweep k = 3,16,64,256,1024,4096 (BASELINE.json configs[1]+[2]),
plus the fraction of the measured HBM roofline.  One "step" = one pass of the sweep (six
kernel launches) over the device-resident signal.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [--gpus N] [--steps K] ...   # the reference CPU path

N > 1 is launched by the driver with torch.distributed.run, one rank per GPU.  The signal is
then N * 2^28 samples sharded contiguously (weak scaling); each rank reads its left context
(whole history tiles, >= k-1 samples) in place from the left neighbour's buffer over NVLink
(CUDA IPC peer mapping; NCCL send/recv fallback).  No data-path collective.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks.  Inputs (1 GiB) and outputs (1 GiB) per launch are far larger than the 126 MB L2, so
no explicit flush is needed between iterations.

After the timed region the same process adds the evidence blocks of the line (none of them inside
the headline timing): `parity` (every rank's output of every k of the sweep against the fp64
oracle recomputed from the generator: shard head, random positions, last sample), `dense_k`
(>= 48 windows: every lag misalignment, the mode seams 8/9 and 256/257, primes, 4095),
`i16` (the reference's own sample type, stereo, full-size bit-exact check against the CPU oracle),
`configs` (BASELINE.json configs[3] and [4] at the current world size) and `e2e` (host buffers).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KS_DEFAULT = [3, 16, 64, 256, 1024, 4096]
# dense pass: seams of the arithmetic modes (8|9 additions-only -> direct sums, 256|257 direct -> tile-rebased scan),
# every lag misalignment (k mod 4), primes, tile-size neighbours
KS_DENSE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 23, 31, 32, 33, 37, 61, 63, 64, 65, 100, 127, 128, 129, 200,
            251, 255, 256, 257, 258, 259, 260, 500, 509, 511, 512, 513, 1000, 1021, 1023, 1024, 1025, 1500, 2039, 2047,
            2048, 2049, 3000, 4093, 4094, 4095, 4096]
SEED = 0x5EED0001
BYTES_PER_SAMPLE = 8  # 4 B read + 4 B written


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mavg", choices=["mavg", "reference"])
    ap.add_argument("--samples-log2", type=int, default=28, help="samples per GPU (log2)")
    ap.add_argument("--ks", default=",".join(str(k) for k in KS_DEFAULT))
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="ipc", choices=["ipc", "nccl"])
    ap.add_argument("--tune", default="", help="comma list key=value forwarded to mavg_tuning")
    ap.add_argument("--overlap", type=int, default=2,
                    help="mavg_tuning.overlap of the bench plans: 2 = programmatic dependent launch, tile loads may start "
                         "under the previous kernel's tail (the input buffer is never written), 1 = wait before the first "
                         "load, 3 = off")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of replaying a CUDA graph of one step")
    ap.add_argument("--skip", default="", help="comma list of evidence blocks to skip: parity,dense_k,i16,configs,shapes,e2e,steps")
    ap.add_argument("--configs-log2", type=int, default=32, help="total samples (log2) of BASELINE configs 4 and 5")
    ap.add_argument("--ref-samples-log2", type=int, default=0,
                    help="reference arm: samples per k and step (log2); 0 = the GPU arm's --samples-log2")
    return ap.parse_args(argv)


def peaks():
    """(GB/s, source): MAVG_HBM_PEAK_GBS in the environment, else MEASURED_PEAKS.json (driver-written), else the
    fallback B200_PROFILING.md states."""
    env = os.environ.get("MAVG_HBM_PEAK_GBS")
    if env:
        try:
            return float(env), "environment (MAVG_HBM_PEAK_GBS)"
        except ValueError:
            pass
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
    committed `ncu --set full` capture of this same workload (three launches: k = 3, 64, 4096); None when the
    summary is missing.  Never measured under the profiler at bench time."""
    import csv
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "ncu_stream_full.csv")
        try:
            rows = {r[0]: r for r in csv.reader(open(p))}
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            total = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                r = rows[key]
                vals = [float(v) for v in r[2:]]
                total += scale[r[1]] * sum(vals) / len(vals)
            return total, "profiles/%s/ncu_stream_full.csv" % rnd
        except Exception:
            continue
    return None, None


def metric_name(samples_log2, ks):
    """BASELINE.json's metric on its named configuration; both arms print the same string."""
    return "Gsamples/s, 2^%d-sample mono float32 moving average per GPU, k sweep %s" % (samples_log2, ks)


def workload_name(samples_log2, world, ks):
    n = 1 << samples_log2
    return ("mono float32 synthetic U[0,1) signal, 2^%d samples per GPU (contiguous shards of one %d-sample signal), "
            "window sweep k=%s, device resident" % (samples_log2, world * n, ks))


def config_block(samples_log2, world, ks):
    """The `config` object; identical in both arms (arm-specific detail lives in other keys of the line)."""
    return {"workload": workload_name(samples_log2, world, ks), "samples_per_gpu": 1 << samples_log2, "ks": list(ks)}


# ---------------------------------------------------------------------------- reference arm
def cpu_reference_line(args, ks, world):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference) timed on the
    host cores.  It is single threaded by construction (basics/profilable_moving_averager.cpp:14-37),
    so cores = 1; samples are int16 because wav_header.h:34 admits nothing else.  One step = the k sweep over
    one GPU's share of the workload (2^28 samples per k by default: the same count the GPU arm filters per k)."""
    import oracle
    log2 = args.ref_samples_log2 or args.samples_log2
    n = 1 << log2
    kind = "reference" if oracle.ref_available() else "port"
    x = oracle.fill_i16(n, SEED)

    def one_step():
        t = 0.0
        for k in ks:
            if kind == "reference":
                t += oracle.ref_time_i16(x, k, 1, iters=1)
            else:
                t += oracle.time_best(0, x, k, 1, 1, iters=1)
        return t

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    ms = 1e3 * total / max(1, args.steps)
    value = len(ks) * n * args.steps / total / 1e9
    sample = (f"mono int16 (the reference's only sample type), 2^{log2} samples x k in {ks} per step = one GPU's share of "
              "the workload, single thread (the reference CPU path has no threads)")
    ts = sorted(times)
    line = {
        "impl": "reference",
        "metric": metric_name(args.samples_log2, ks),
        "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i16", "data": "synthetic",
        "config": config_block(args.samples_log2, world, ks),
        "reference_arm": "profilable_cpu_computations (basics/profilable_moving_averager.cpp:14-37) compiled unmodified "
                         "into oracle/_ref, rank 0 only; its throughput does not depend on k or on the signal length",
        "step_ms": {"min": 1e3 * ts[0], "median": 1e3 * ts[len(ts) // 2], "mean": ms},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host_cores": os.cpu_count(),
    }
    return line


def cpu_baseline_block(ks):
    """cpu_baseline for the main line: the reference on one thread (it has no threads), plus our
    float32 port on all cores for context.  Bounded: 2^26 int16 samples per k."""
    import oracle
    out = {}
    n = 1 << 26
    x = oracle.fill_i16(n, SEED)
    kind = "reference" if oracle.ref_available() else "port"
    t = 0.0
    for k in ks:
        t += oracle.ref_time_i16(x, k, 1, iters=1) if kind == "reference" else oracle.time_best(0, x, k, 1, 1, 1)
    out["cpu_baseline"] = {"value": len(ks) * n / t / 1e9, "unit": "Gsamples/s", "cores": 1, "kind": kind,
                           "sample": f"mono int16 2^26 samples x k in {ks}, once each, single thread "
                                     "(the reference CPU path has no threads)"}
    cores = os.cpu_count() or 1
    xf = oracle.fill_f32(n, SEED)
    t = 0.0
    for k in ks:
        t += oracle.time_best(2, xf, k, 1, cores, 1)
    out["cpu_port_all_cores"] = {"value": len(ks) * n / t / 1e9, "unit": "Gsamples/s", "cores": cores, "kind": "port",
                                 "sample": f"float32 running-sum port, 2^26 samples x k in {ks}, {cores} pthreads"}
    return out


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def sample(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            time.sleep(0.005)

    def start(self):
        self._t.start()

    def stop(self):
        self._stop.set()
        self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------- device helpers
class _Arr:
    """__cuda_array_interface__ view of a raw device pointer (torch.as_tensor wraps it without copying)."""

    def __init__(self, ptr, count, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


class Ctx:
    """Everything the evidence blocks share: rank info, stream, library handles."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import digital_signal_processsing_b200 as mavg
        from digital_signal_processsing_b200 import _lib, sharding
        self.torch, self.dist, self.mavg, self._lib, self.sharding = torch, dist, mavg, _lib, sharding
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.lib = None
        self.stream = None
        self.tune = {}
        self.halo_mode = "none"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def allmax(self, vals):
        """element-wise max over ranks of a list of floats"""
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def allmin(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return [float(v) for v in t]

    def alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._lib.check(self.lib.mavg_device_alloc(nbytes, ctypes.byref(p)))
        return p

    def free(self, p):
        if p is not None and p.value:
            self.lib.mavg_device_free(p)
            p.value = None

    def plan(self, frames, k, **kw):
        kw = dict(kw)
        for key, val in self.tune.items():
            kw.setdefault(key, val)
        p = self.mavg.Plan(frames, k, **kw)
        p.set_stream(self.stream.cuda_stream)
        p.enable_timing(False)   # bench.py times the stream itself; skip the plan's four event records per run
        return p

    def time_launches(self, fn, reps, warm=2):
        """mean ms per call of `fn` over `reps` back-to-back calls on the stream, max over ranks"""
        torch = self.torch
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.stream.synchronize()
        self.barrier()
        a.record(self.stream)
        for _ in range(reps):
            fn()
        b.record(self.stream)
        b.synchronize()
        return self.allmax([a.elapsed_time(b) / reps])[0]


class LeftContext:
    """Left context of a rank's shard: the last `max_elems` elements of the left neighbour's INPUT buffer, read in place
    through a CUDA IPC peer mapping (the kernels' TMA loads go over NVLink) or staged once over NCCL send/recv.
    ptr(h) = address of the last h elements (None on rank 0: zero padding, the start of the signal)."""

    def __init__(self, cx, d_in_ptr, n_elems, elem_bytes, max_elems, torch_dtype, typestr, mode="ipc"):
        self.cx, self.es, self.max = cx, elem_bytes, int(max_elems)
        self.peer, self.staged, self.base, self.mode = None, None, 0, "none"
        if cx.world == 1 or self.max == 0:
            return
        if mode == "ipc":
            try:
                self.peer = cx.sharding.PeerHalo(d_in_ptr, n_elems, elem_bytes, self.max, cx.rank, cx.world)
                self.base = self.peer.halo_ptr
                self.mode = "ipc-peer (TMA reads the neighbour's tail in place over NVLink)"
            except Exception as e:  # pragma: no cover - depends on the box
                if cx.rank == 0:
                    print(f"[bench] CUDA IPC unavailable ({e}); using NCCL send/recv", file=sys.stderr)
                self.peer = None
        if self.peer is None:
            shard = cx.torch.as_tensor(_Arr(d_in_ptr, n_elems, typestr), device="cuda")
            self.staged = cx.sharding.exchange_halo(shard, self.max, cx.rank, cx.world)
            cx.torch.cuda.synchronize()
            self.base = self.staged.data_ptr() if self.staged is not None else 0
            self.mode = "nccl send/recv into a staging buffer"

    def ptr(self, h_elems):
        if self.cx.rank == 0 or not self.base:
            return None
        h_elems = int(h_elems)
        if h_elems > self.max:
            raise ValueError("left context asked for %d elements, %d mapped" % (h_elems, self.max))
        return self.base + self.es * (self.max - h_elems)

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        self.staged = None


def spot_check_f32(cx, d_out_ptr, n, first, k, rng, npoints=2000, seed=SEED):
    """max relative error of this rank's output against oracle.point_f64 (fp64 window sum recomputed from the
    generator at the GLOBAL index): the shard's first 2k + 64 samples (they depend on the left context), `npoints`
    random positions and the last sample.  Returns (max_rel_err, points)."""
    import numpy as np
    import oracle
    torch = cx.torch
    idx = np.unique(np.concatenate([np.arange(0, min(n, 2 * k + 64)), rng.integers(0, n, npoints), [n - 1]])).astype(np.int64)
    y = torch.as_tensor(_Arr(d_out_ptr, n), device="cuda")
    got = y[torch.from_numpy(idx).cuda()].cpu().numpy().astype(np.float64)
    worst = 0.0
    for g, i in zip(got, idx):
        e = oracle.point_f64(first + int(i), k, seed)
        err = abs(g - e) / abs(e) if e != 0.0 else abs(g)
        if not (err <= worst):      # also catches NaN
            worst = err if err == err else float("inf")
    return worst, int(idx.size)


# ---------------------------------------------------------------------------- evidence blocks
def block_parity(cx, plans, ks, d_in, d_out, n, first, left, halo_elems):
    """Every k of the sweep, every rank: run once, spot-check the output, all-reduce MAX."""
    import numpy as np
    rng = np.random.default_rng(1234 + cx.rank)
    per_k, checked, worst_all = {}, 0, 0.0
    for k in ks:
        plans[k].run_device_halo(d_in.value, d_out.value, left.ptr(halo_elems[k]))
        cx.stream.synchronize()
        worst, cnt = spot_check_f32(cx, d_out.value, n, first, k, rng)
        worst = cx.allmax([worst])[0]
        per_k[str(k)] = worst
        worst_all = max(worst_all, worst)
        checked += cnt
    total_checked = int(_allsum(cx, checked))
    return {"max_rel_err": worst_all, "tolerance": 1e-5, "ok": bool(worst_all <= 1e-5), "checked": total_checked,
            "ranks": cx.world, "per_k_max_rel_err": per_k,
            "what": "each rank's d_out of each k vs oracle.point_f64 at the global index: first 2k+64 samples of the shard "
                    "(left-context dependent), 2000 random positions, last sample; MAX over ranks"}


def _allsum(cx, v):
    if cx.world == 1:
        return float(v)
    t = cx.torch.tensor([float(v)], device="cuda", dtype=cx.torch.float64)
    cx.dist.all_reduce(t, op=cx.dist.ReduceOp.SUM)
    return float(t[0])


def block_dense_k(cx, d_in, d_out, n, first, left_full, peak, reps=10):
    """>= 48 windows through the same device-resident signal: launch time, roofline fraction, spot-checked error."""
    import numpy as np
    rng = np.random.default_rng(99 + cx.rank)
    res, worst_err = {}, 0.0
    for k in KS_DENSE:
        p = cx.plan(n, k, first_frame=first, overlap=cx.args.overlap)
        h = int(p.info.halo_frames)
        hp = left_full.ptr(h)
        run = lambda: p.run_device_halo(d_in.value, d_out.value, hp)
        ms = min(cx.time_launches(run, reps), cx.time_launches(run, reps, warm=0))   # best of two passes of `reps` launches
        cx.stream.synchronize()
        err, _ = spot_check_f32(cx, d_out.value, n, first, k, rng, npoints=64)
        err = cx.allmax([err])[0]
        worst_err = max(worst_err, err)
        info = p.info
        res[k] = (ms, BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9 / peak, int(info.mode), int(info.path))
        p.close()
    kmin = min(res, key=lambda q: res[q][1])
    return {"windows": len(res), "min_frac_over_k": round(res[kmin][1], 4), "k_at_min": kmin,
            "max_frac_over_k": round(max(v[1] for v in res.values()), 4),
            "max_rel_err": worst_err, "parity_ok": bool(worst_err <= 1e-5),
            "all_stream_path": all(v[3] == 1 for v in res.values()),
            "per_k_ms": {str(k): round(v[0], 4) for k, v in res.items()},
            "per_k_frac": {str(k): round(v[1], 4) for k, v in res.items()},
            "note": "best of two passes of %d back-to-back launches per k on the bench stream (launch gaps included), max over ranks; "
                    "frac = 8 B x samples / ms / measured peak; error spot-checked at the shard head + 64 random positions" % reps}


def block_i16(cx, ks, samples_log2, peak, reps=10):
    """The reference's own sample type (wav_header.h:34): interleaved stereo int16, 2^samples_log2 samples per GPU
    (contiguous shards of one signal), same window sweep; every output sample of every k compared with the CPU oracle
    (oracle.mavg_i16 = the restatement pinned to the reference), bit for bit."""
    import numpy as np
    import oracle
    torch = cx.torch
    C = 2
    n = 1 << samples_log2                 # samples per rank
    frames = n // C
    first_frame = cx.rank * frames
    d_in, d_out = cx.alloc(2 * n), cx.alloc(2 * n)
    cx.mavg.fill_synthetic_device(d_in.value, "i16", n, first_frame * C, SEED, 0, cx.stream.cuda_stream)
    cx.stream.synchronize()
    plans = {k: cx.plan(frames, k, channels=C, dtype="i16", first_frame=first_frame, overlap=cx.args.overlap) for k in ks}
    halo = {k: int(plans[k].info.halo_frames) * C for k in ks}
    left = LeftContext(cx, d_in.value, n, 2, max(halo.values()), torch.int16, "<i2", cx.args.halo)
    cx.barrier()
    # host copy of the input with max(k) frames of left context (the oracle warms up inside it)
    ctx_frames = max(ks) if cx.rank > 0 else 0
    x_host = oracle.fill_i16(n + ctx_frames * C, SEED, first_index=(first_frame - ctx_frames) * C)
    threads = max(1, (os.cpu_count() or 1) // max(1, cx.world))
    per_k, exact_all, mism_total = {}, True, 0
    y_dev = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda")
    for k in ks:
        p = plans[k]
        hp = left.ptr(halo[k])
        ms = cx.time_launches(lambda: p.run_device_halo(d_in.value, d_out.value, hp), reps)
        cx.stream.synchronize()
        got = y_dev.cpu().numpy()
        exp = (oracle.mavg_i16_mt(x_host, k, C, threads) if threads > 1 else oracle.mavg_i16(x_host, k, C))[ctx_frames * C:]
        mism = int(np.count_nonzero(got != exp))
        mism = int(_allsum(cx, mism))
        exact_all = exact_all and mism == 0
        mism_total += mism
        info = p.info
        per_k[str(k)] = {"ms": round(ms, 4), "gsamples_s": round(n / (ms * 1e-3) / 1e9, 1),
                         "hbm_gbs": round(4 * n / (ms * 1e-3) / 1e9, 1),
                         "frac_measured": round(4 * n / (ms * 1e-3) / 1e9 / peak, 4), "mismatches": mism,
                         "mode": int(info.mode), "path": "stream" if info.path == 1 else "generic"}
    for p in plans.values():
        p.close()
    left.close()
    cx.barrier()
    cx.free(d_in)
    cx.free(d_out)
    fr = [v["frac_measured"] for v in per_k.values()]
    return {"workload": "interleaved stereo int16, 2^%d samples per GPU (shards of one %d-sample signal), k sweep %s, "
                        "device resident" % (samples_log2, cx.world * n, ks),
            "bit_exact": bool(exact_all), "mismatches": mism_total, "compared_samples_per_k": cx.world * n,
            "algorithmic_bytes_per_sample": 4, "min_frac": min(fr), "per_k": per_k, "halo": left.mode,
            "note": "%d back-to-back launches per k; every output sample of every rank compared with the CPU oracle "
                    "(oracle_mavg_i16: int64 sums, truncating division)" % reps}


def block_configs(cx, total_log2, peak, reps=10):
    """BASELINE.json configs[3] (one 2^32-sample signal sharded contiguously, k = 1024, halo over NVLink) and
    configs[4] (256 channels x 2^24 frames, k = 64; planar = channels partitioned, interleaved = frames partitioned
    with a halo) at the current world size, through the same plan API, spot-checked against the fp64 oracle."""
    import numpy as np
    import oracle
    torch = cx.torch
    out = {}
    total = 1 << total_log2
    n = total // cx.world                 # samples per rank in all three layouts
    d_in, d_out = cx.alloc(4 * n), cx.alloc(4 * n)
    seed = 0x5EED0004
    y = torch.as_tensor(_Arr(d_out.value, n), device="cuda")

    def finish(name, ms, worst, extra):
        worst = cx.allmax([worst])[0]
        out[name] = dict(ms=round(ms, 4), gsamples_s=round(total / (ms * 1e-3) / 1e9, 1),
                         hbm_gbs_per_gpu=round(8 * n / (ms * 1e-3) / 1e9, 1),
                         frac_measured=round(8 * n / (ms * 1e-3) / 1e9 / peak, 4), max_rel_err_spot=worst,
                         ok=bool(worst <= 1e-5), **extra)

    # ---- config 4: mono, k = 1024
    k = 1024
    first = cx.rank * n
    cx.mavg.fill_synthetic_device(d_in.value, "f32", n, first, seed, 0, cx.stream.cuda_stream)
    cx.stream.synchronize()
    p = cx.plan(n, k, first_frame=first, overlap=cx.args.overlap)
    h = int(p.info.halo_frames)
    left = LeftContext(cx, d_in.value, n, 4, h, torch.float32, "<f4", cx.args.halo)
    cx.barrier()
    ms = cx.time_launches(lambda: p.run_device_halo(d_in.value, d_out.value, left.ptr(h)), reps)
    cx.stream.synchronize()
    worst, cnt = spot_check_f32(cx, d_out.value, n, first, k, np.random.default_rng(7 + cx.rank), npoints=3000, seed=seed)
    finish("config4", ms, worst, {"workload": "mono f32, one 2^%d-sample signal sharded over %d GPU(s), k=1024, halo %d samples (%s)"
                                  % (total_log2, cx.world, h, left.mode), "spot_checks_per_rank": cnt})
    p.close()
    left.close()
    cx.barrier()

    # ---- config 5, planar [C][F]: channels partitioned, no communication
    C, k = 256, 64
    F = total // C
    if C % cx.world == 0:
        c_per = C // cx.world
        first = cx.rank * n                # channel c of the batch occupies [c*F, (c+1)*F) of the generator's index space
        cx.mavg.fill_synthetic_device(d_in.value, "f32", n, first, seed, 0, cx.stream.cuda_stream)
        cx.stream.synchronize()
        p = cx.plan(F, k, channels=c_per, layout="planar", overlap=cx.args.overlap)
        ms = cx.time_launches(lambda: p.run_device([d_in.value], [d_out.value]), reps)
        cx.stream.synchronize()
        rng = np.random.default_rng(11 + cx.rank)
        cs = rng.integers(0, c_per, 300)
        fs = np.where(rng.random(300) < 0.2, rng.integers(0, 2 * k, 300), rng.integers(0, F, 300))
        got = y[torch.from_numpy(cs * F + fs).cuda()].cpu().numpy().astype(np.float64)
        worst = 0.0
        for g, c, f in zip(got, cs, fs):
            lo = max(0, int(f) - k + 1)
            e = oracle.fill_f32(int(f) - lo + 1, seed, 0, first_index=first + int(c) * F + lo).astype(np.float64).sum() / k
            worst = max(worst, abs(g - e) / e)
        finish("config5_planar", ms, worst, {"workload": "256 ch x 2^%d frames f32 planar [C][F], k=64, %d channels per GPU"
                                             % (total_log2 - 8, c_per), "path": "stream" if p.info.path == 1 else "generic"})
        p.close()

    # ---- config 5, interleaved [F][C] (the reference layout): frames partitioned, halo frames over NVLink
    f_per = F // cx.world
    first_frame = cx.rank * f_per
    cx.mavg.fill_synthetic_device(d_in.value, "f32", n, first_frame * C, seed, 0, cx.stream.cuda_stream)
    cx.stream.synchronize()
    p = cx.plan(f_per, k, channels=C, layout="interleaved", first_frame=first_frame)
    h = int(p.info.halo_frames)
    left = LeftContext(cx, d_in.value, n, 4, h * C, torch.float32, "<f4", cx.args.halo)
    cx.barrier()
    ms = cx.time_launches(lambda: p.run_device_halo(d_in.value, d_out.value, left.ptr(h * C)), reps)
    cx.stream.synchronize()
    rng = np.random.default_rng(13 + cx.rank)
    cs = rng.integers(0, C, 300)
    fs = np.where(rng.random(300) < 0.2, rng.integers(0, 2 * k, 300), rng.integers(0, f_per, 300))
    got = y[torch.from_numpy(fs * C + cs).cuda()].cpu().numpy().astype(np.float64)
    worst = 0.0
    for g, c, f in zip(got, cs, fs):
        gf = first_frame + int(f)
        lo = max(0, gf - k + 1)
        e = oracle.fill_f32((gf - lo + 1) * C, seed, 0, first_index=lo * C)[int(c)::C].astype(np.float64).sum() / k
        worst = max(worst, abs(g - e) / e)
    finish("config5_interleaved", ms, worst, {"workload": "256 ch x 2^%d frames f32 interleaved [F][C], k=64, frames partitioned, "
                                              "halo %d frames (%s)" % (total_log2 - 8, h, left.mode),
                                              "path": "stream" if p.info.path == 1 else "generic"})
    p.close()
    left.close()
    cx.barrier()
    cx.free(d_in)
    cx.free(d_out)
    out["note"] = "%d back-to-back launches each, max over ranks; gsamples_s is the whole job" % reps
    return out


def block_shapes(cx, peak, reps=10):
    """The other kernels of the library on the shapes they exist for, 2^27 samples each, every rank the same signal:
    multichannel int16 (flat-stream kernel: 4 / 6 / 8 channels; column kernel: 64 / 256 channels), every output sample
    compared with the CPU oracle; a far window (float32, k = 60 000, far-lag kernel), spot-checked against the fp64
    oracle; the prefix-sum primitive (int16 -> int64, 2^28 samples), head compared with an int64 cumsum."""
    import numpy as np
    import oracle
    torch = cx.torch
    threads = max(1, (os.cpu_count() or 1) // max(1, cx.world))
    res = {}
    for name, C, k in (("i16_c4_k1024", 4, 1024), ("i16_c6_k64", 6, 64), ("i16_c8_k64", 8, 64), ("i16_c64_k64", 64, 64),
                       ("i16_c256_k64", 256, 64)):
        frames = (1 << 27) // C
        n = frames * C
        d_in, d_out = cx.alloc(2 * n), cx.alloc(2 * n)
        cx.mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, cx.stream.cuda_stream)
        cx.stream.synchronize()
        p = cx.plan(frames, k, channels=C, dtype="i16")
        # Best of two passes, the second one into a second output allocation: two of eight runs of this block on the pool's
        # boxes measured ONE shape 30 % slow for as long as its buffers lived (0.118 instead of 0.090 ms; the same shape is
        # insensitive to the VIRTUAL distance of its buffers, tests/perf/alias_probe.py) -- a property of where the
        # allocation landed physically, not of the kernel.
        ms = cx.time_launches(lambda: p.run_device([d_in.value], [d_out.value]), reps, warm=5)
        d_out2 = cx.alloc(2 * n)
        ms = min(ms, cx.time_launches(lambda: p.run_device([d_in.value], [d_out2.value]), reps, warm=5))
        cx.free(d_out2)
        p.run_device([d_in.value], [d_out.value])
        cx.stream.synchronize()
        got = torch.as_tensor(_Arr(d_out.value, n, "<i2"), device="cuda").cpu().numpy()
        x = oracle.fill_i16(n, SEED)
        exp = oracle.mavg_i16_mt(x, k, C, threads) if threads > 1 else oracle.mavg_i16(x, k, C)
        mism = int(_allsum(cx, int(np.count_nonzero(got != exp))))
        info = p.info
        res[name] = {"ms": round(ms, 4), "hbm_gbs": round(4 * n / (ms * 1e-3) / 1e9, 1),
                     "frac_measured": round(4 * n / (ms * 1e-3) / 1e9 / peak, 4), "mismatches": mism,
                     "compared_samples": n, "mode": int(info.mode), "threads": int(info.threads), "run": int(info.run)}
        p.close()
        cx.free(d_in)
        cx.free(d_out)
    # far window, float32 mono
    n, k = 1 << 27, 60_000
    d_in, d_out = cx.alloc(4 * n), cx.alloc(4 * n)
    cx.mavg.fill_synthetic_device(d_in.value, "f32", n, 0, SEED, cx.mavg.DIST_U01, cx.stream.cuda_stream)
    cx.stream.synchronize()
    p = cx.plan(n, k)
    ms = min(cx.time_launches(lambda: p.run_device([d_in.value], [d_out.value]), reps, warm=5) for _ in range(2))
    cx.stream.synchronize()
    worst, cnt = spot_check_f32(cx, d_out.value, n, 0, k, np.random.default_rng(99), npoints=150)
    worst = cx.allmax([worst])[0]
    info = p.info
    res["f32_c1_k60000"] = {"ms": round(ms, 4), "hbm_gbs": round(8 * n / (ms * 1e-3) / 1e9, 1),
                            "frac_measured": round(8 * n / (ms * 1e-3) / 1e9 / peak, 4), "max_rel_err_spot": worst,
                            "ok": bool(worst <= 1e-5), "mode": int(info.mode), "threads": int(info.threads)}
    p.close()
    cx.free(d_in)
    cx.free(d_out)
    # prefix-sum primitive
    n = 1 << 28
    d_in, d_out = cx.alloc(2 * n), cx.alloc(8 * n)
    cx.mavg.fill_synthetic_device(d_in.value, "i16", n, 0, SEED, 0, cx.stream.cuda_stream)
    cx.stream.synchronize()
    ms = min(cx.time_launches(lambda: cx.mavg.prefix_sum_device(d_in.value, d_out.value, "i16", n, 1, cx.stream.cuda_stream), reps,
                              warm=5) for _ in range(2))
    cx.stream.synchronize()
    m = 1 << 22
    got = torch.as_tensor(_Arr(d_out.value, n, "<i8"), device="cuda")
    head_ok = bool(np.array_equal(got[:m].cpu().numpy(), np.cumsum(oracle.fill_i16(m, SEED).astype(np.int64))))
    # the last prefix equals the sum of all samples: a checksum over the whole look-back chain
    total = int(torch.as_tensor(_Arr(d_in.value, n, "<i2"), device="cuda").to(torch.int64).sum().item())
    last_ok = int(got[n - 1].item()) == total
    res["prefix_sum_i16_2^28"] = {"ms": round(ms, 4), "hbm_gbs": round(10 * n / (ms * 1e-3) / 1e9, 1),
                                  "frac_measured": round(10 * n / (ms * 1e-3) / 1e9 / peak, 4),
                                  "ok": bool(_allsum(cx, int(not (head_ok and last_ok))) == 0),
                                  "what": "int16 -> int64, 2 + 8 bytes per sample; head of 2^22 prefixes == cumsum, last prefix == sum of all samples"}
    cx.free(d_in)
    cx.free(d_out)
    ok = all(v.get("mismatches", 0) == 0 and v.get("ok", True) for v in res.values())
    return {"ok": bool(ok), "per_shape": res,
            "note": "best of two passes of %d back-to-back launches per shape on the bench stream (launch gaps included), max over ranks; int16 "
                    "shapes: 4 bytes per sample, every output sample compared with the CPU oracle" % reps}


def host_copy_ceiling(cx, nbytes, reps=3):
    """What the host side can deliver with NO kernel in the way: pinned H2D and D2H copies of `nbytes` each running
    at the same time on two streams, all ranks at once (so that shared PCIe uplinks / one NUMA node show up)."""
    torch = cx.torch
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = None
    for i in range(reps + 1):
        cx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = cx.allmax([time.perf_counter() - t0])[0]
        if i > 0:
            best = dt if best is None else min(best, dt)
    gbs = nbytes / best / 1e9
    # the download alone (what bounds the sweep call: one upload, one download per window)
    best_d = None
    for i in range(reps + 1):
        cx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        dt = cx.allmax([time.perf_counter() - t0])[0]
        if i > 0:
            best_d = dt if best_d is None else min(best_d, dt)
    gbs_d = nbytes / best_d / 1e9
    return {"gbs_each_way_per_gpu": round(gbs, 2), "ranks": cx.world,
            "gsamples_s_f32": round(cx.world * gbs / 4, 2), "gsamples_s_i16": round(cx.world * gbs / 2, 2),
            "d2h_alone_gbs_per_gpu": round(gbs_d, 2),
            "sweep_gsamples_s_f32": round(cx.world * gbs_d / 4, 2), "sweep_gsamples_s_i16": round(cx.world * gbs_d / 2, 2),
            "what": "pinned cudaMemcpyAsync H2D + D2H of %d MiB each, concurrently, all ranks at once, best of %d (wall "
                    "clock, max over ranks); gsamples_s_* = what this host side allows the per-window calls (as many bytes "
                    "up as down); sweep_gsamples_s_* = what the download alone allows the sweep call (one output sample "
                    "per sample and window, the upload of the one input hidden under it)" % (nbytes >> 20, reps)}


def block_e2e(cx, ks, samples_log2, steps):
    """The same metric end to end through mavg_run_host (the C-ABI call XxxGpuLoad maps to): pinned HOST buffers,
    H2D + kernel + D2H inside the timing.  float32 = the headline; the int16 leg is the reference's own sample type."""
    torch, mavg = cx.torch, cx.mavg
    n = 1 << samples_log2
    first = cx.rank * n
    out = {}

    def leg(dtype, tdtype, es, channels):
        frames = n // channels
        first_frame = cx.rank * frames
        hplans = {k: mavg.Plan(frames, k, channels=channels, dtype=dtype, first_frame=first_frame, **cx.tune) for k in ks}
        halo_elems = max(int(p.info.halo_frames) for p in hplans.values()) * channels if cx.rank > 0 else 0
        # mavg_run_host wants each plan's own halo directly in front of the shard: lay out [max halo | shard]
        h_in = torch.empty(halo_elems + n, dtype=tdtype, pin_memory=True)
        h_out = torch.empty(n, dtype=tdtype, pin_memory=True)
        tmp = torch.empty(halo_elems + n, dtype=tdtype, device="cuda")
        mavg.fill_synthetic_device(tmp.data_ptr(), dtype, halo_elems + n, first_frame * channels - halo_elems, SEED,
                                   mavg.DIST_U01, torch.cuda.current_stream().cuda_stream)
        h_in.copy_(tmp)
        torch.cuda.synchronize()
        del tmp
        in_ptr = h_in.data_ptr() + es * halo_elems

        def host_step():
            for k in ks:
                hplans[k].run_host_ptr(in_ptr, h_out.data_ptr())   # H2D + kernel + D2H, blocking

        host_step()  # warm-up: allocates the plan-owned device buffers
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            host_step()
        dt = cx.allmax([time.perf_counter() - t0])[0]
        tm = hplans[ks[-1]].timing()
        # the last k's output is still in h_out: check it against the oracle at a few points
        res = {"value": cx.world * n * len(ks) * steps / dt / 1e9, "unit": "Gsamples/s",
               "h2d_bytes_per_step": es * (n + halo_elems) * len(ks), "d2h_bytes_per_step": es * n * len(ks),
               "steps": steps, "api": "mavg_run_host (pinned host buffers, H2D + kernel + D2H per k)",
               "gbs_each_way_per_gpu": round(es * n * len(ks) * steps / dt / 1e9, 2),
               "last_call_phases_ms": {"h2d": round(tm.h2d_ms, 3), "compute": round(tm.compute_ms, 3),
                                       "d2h": round(tm.d2h_ms, 3)}}
        res["check"] = e2e_check(cx, h_out, dtype, channels, first_frame, ks[-1])
        # the same sweep through mavg_run_host_sweep: the step's input crosses the host link ONCE, every window's result
        # comes back (one pinned output buffer per window)
        try:
            outs = [torch.empty(n, dtype=tdtype, pin_memory=True) for _ in ks]
            plist = [hplans[k] for k in ks]
            optrs = [o.data_ptr() for o in outs]
            mavg.run_host_sweep_ptr(plist, in_ptr, optrs)      # warm-up: allocates the sweep's device input buffer
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                mavg.run_host_sweep_ptr(plist, in_ptr, optrs)
            dts = cx.allmax([time.perf_counter() - t0])[0]
            tms = hplans[ks[0]].timing()
            res["sweep"] = {"value": cx.world * n * len(ks) * steps / dts / 1e9, "unit": "Gsamples/s",
                            "h2d_bytes_per_step": es * (n + halo_elems), "d2h_bytes_per_step": es * n * len(ks),
                            "steps": steps, "api": "mavg_run_host_sweep (one upload of the signal per step, one pinned "
                                                   "output per window; results bit-identical to the per-window calls)",
                            "phases_ms": {"h2d": round(tms.h2d_ms, 3), "compute": round(tms.compute_ms, 3),
                                          "d2h": round(tms.d2h_ms, 3)},
                            "check_first_k": e2e_check(cx, outs[0], dtype, channels, first_frame, ks[0]),
                            "check_last_k": e2e_check(cx, outs[-1], dtype, channels, first_frame, ks[-1])}
            del outs
        except Exception as e:  # pragma: no cover
            res["sweep"] = {"error": "%s: %s" % (type(e).__name__, e)}
        for p in hplans.values():
            p.close()
        # Headline of the block = the call a user makes for this workload (one signal, a sweep of windows):
        # mavg_run_host_sweep.  The per-window calls (one mavg_run_host per k, the call every XxxGpuLoad maps to; the
        # input is uploaded again for every window) stay in the line next to it.
        sw = res.pop("sweep", None)
        if sw and "value" in sw:
            per_call = {key: res[key] for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps", "api",
                                                  "gbs_each_way_per_gpu", "last_call_phases_ms", "check")}
            res = {"value": sw["value"], "unit": sw["unit"], "h2d_bytes_per_step": sw["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": sw["d2h_bytes_per_step"], "steps": sw["steps"], "api": sw["api"],
                   "phases_ms": sw["phases_ms"], "check": sw["check_last_k"], "check_first_k": sw["check_first_k"],
                   "per_window_calls": per_call}
        elif sw:
            res["sweep_error"] = sw.get("error")
        return res

    # the copy probe runs before and after the legs and the better pass is reported: PCIe throughput of a shared box
    # drifts by tens of percent within seconds, and a probe taken in a slow moment would sit below the pipeline it bounds
    try:
        ceil0 = host_copy_ceiling(cx, 4 * n)
    except Exception:  # pragma: no cover
        ceil0 = None
    out = leg("f32", torch.float32, 4, 1)
    try:
        out["i16"] = leg("i16", torch.int16, 2, 2)
        out["i16"]["workload"] = "interleaved stereo int16, 2^%d samples per GPU, same k sweep" % samples_log2
    except Exception as e:  # pragma: no cover
        out["i16"] = {"error": str(e)}
    try:
        ceil1 = host_copy_ceiling(cx, 4 * n)
        if ceil0 and ceil0["gbs_each_way_per_gpu"] > ceil1["gbs_each_way_per_gpu"]:
            ceil0, ceil1 = ceil1, ceil0
        if ceil0:
            ceil1["other_pass_gbs_each_way_per_gpu"] = ceil0["gbs_each_way_per_gpu"]
        out["host_ceiling"] = ceil1
    except Exception as e:  # pragma: no cover
        out["host_ceiling"] = {"error": str(e)}
    return out


def e2e_check(cx, h_out, dtype, channels, first_frame, k):
    """A few points of the host output of the last run_host call against the oracle."""
    import numpy as np
    import oracle
    y = h_out.numpy()
    n = y.size
    rng = np.random.default_rng(5 + cx.rank)
    if dtype == "f32":
        idx = np.unique(np.concatenate([np.arange(0, min(n, 64)), rng.integers(0, n, 200), [n - 1]]))
        worst = 0.0
        for i in idx:
            e = oracle.point_f64(first_frame + int(i), k, SEED)
            worst = max(worst, abs(float(y[i]) - e) / abs(e) if e else abs(float(y[i])))
        worst = cx.allmax([worst])[0]
        return {"max_rel_err": worst, "ok": bool(worst <= 1e-5), "points": int(idx.size)}
    # int16: the first 2^16 frames of the shard and its last 2^16 frames, bit for bit
    m = min(n // channels, 1 << 16)
    bad = 0
    for f0 in (0, n // channels - m):
        gf0 = first_frame + f0
        lo = max(0, gf0 - k)
        x = oracle.fill_i16((gf0 + m - lo) * channels, SEED, first_index=lo * channels)
        exp = oracle.mavg_i16(x, k, channels)[(gf0 - lo) * channels:]
        bad += int(np.count_nonzero(exp != y[f0 * channels:(f0 + m) * channels]))
    bad = int(_allsum(cx, bad))
    return {"mismatches": bad, "ok": bad == 0, "frames_compared": 2 * m}


# ---------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    ks = [int(v) for v in args.ks.split(",") if v]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    skip = set(s for s in args.skip.split(",") if s)

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(cpu_reference_line(args, ks, max(world, args.gpus))), flush=True)
        return 0

    # Everything except the final JSON line goes to stderr (NCCL prints a version banner on stdout).
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    cx = Ctx(args)
    torch, dist, mavg, _lib = cx.torch, cx.dist, cx.mavg, cx._lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmavg has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    barrier = cx.barrier

    n = 1 << args.samples_log2            # samples per rank
    first = rank * n                      # global index of this rank's first sample
    tune = dict(kv.split("=") for kv in args.tune.split(",") if kv)
    cx.tune = {k: int(v) for k, v in tune.items()}
    cx.lib = lib = _lib.load()

    # device buffers from plain cudaMalloc (exportable over CUDA IPC)
    d_in, d_out = cx.alloc(4 * n), cx.alloc(4 * n)
    cx.stream = stream = torch.cuda.Stream()
    mavg.fill_synthetic_device(d_in.value, "f32", n, first, SEED, mavg.DIST_U01, stream.cuda_stream)
    stream.synchronize()

    plans = {k: cx.plan(n, k, first_frame=first, overlap=args.overlap) for k in ks}
    halo_elems = {k: int(plans[k].info.halo_frames) for k in ks}
    # left context: read in place from the left neighbour (IPC) or staged once over NCCL; sized for the dense-k pass too
    max_halo = max(max(halo_elems.values()), 8192)
    left = LeftContext(cx, d_in.value, n, 4, max_halo, torch.float32, "<f4", args.halo)
    halo_ptr = {k: left.ptr(halo_elems[k]) for k in ks}
    barrier()

    launches_per_step = 0

    def step():
        nonlocal launches_per_step
        cnt = 0
        for k in ks:
            plans[k].run_device_halo(d_in.value, d_out.value, halo_ptr[k])
            cnt += int(plans[k].info.launches_per_run)
        launches_per_step = cnt

    with torch.cuda.stream(stream):
        for _ in range(max(3, args.warmup)):
            step()
    stream.synchronize()

    # One step (the six launches of the k sweep) captured into a CUDA graph: libmavg's launches are capture-safe,
    # and replaying the graph removes most of the host-side launch latency between the 0.34 ms kernels.
    graph = None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                step()
            with torch.cuda.stream(stream):
                graph.replay()
            stream.synchronize()
        except Exception as e:  # pragma: no cover - depends on the box
            print(f"[bench] CUDA graph capture unavailable ({e}); timing plain stream launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            step()

    # ---------------- timed region: EXACTLY K steps, bracketed by events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    with torch.cuda.stream(stream):
        e0.record()
        for s in range(args.steps):
            one_step()
        e1.record()
    sampler.sample()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    elapsed_ms = cx.allmax([elapsed_ms])[0]

    peak, peak_src = peaks()
    line_extra = {}

    # ---------------- per-step distribution: a separate pass, one event pair per step
    if "steps" not in skip:
        nst = max(5, min(args.steps, 30))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nst + 1)]
        barrier()
        with torch.cuda.stream(stream):
            evs[0].record()
            for i in range(nst):
                one_step()
                evs[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(cx.allmax([evs[i].elapsed_time(evs[i + 1]) for i in range(nst)]))
        line_extra["step_ms"] = {"min": round(ts[0], 4), "median": round(ts[len(ts) // 2], 4),
                                 "mean": round(sum(ts) / len(ts), 4), "max": round(ts[-1], 4), "n": nst,
                                 "note": "separate pass after the timed region, one event pair per step (max over ranks per step)"}

    # ---------------- per-window detail: a separate pass of back-to-back launches of each k
    reps = max(3, min(args.steps, 20))
    per_k_ms = [cx.time_launches(lambda k=k: plans[k].run_device_halo(d_in.value, d_out.value, halo_ptr[k]), reps, warm=1)
                for k in ks]

    total_samples = world * n * len(ks) * args.steps
    value = total_samples / (elapsed_ms * 1e-3) / 1e9
    kernel_ms = elapsed_ms / (args.steps * len(ks))   # mean launch duration over the timed region (events e0..e1)
    achieved = BYTES_PER_SAMPLE * n / (kernel_ms * 1e-3) / 1e9
    per_k = {str(k): {"ms": round(ms, 4), "gsamples_s": round(n / (ms * 1e-3) / 1e9, 1),
                      "hbm_gbs": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9, 1),
                      "frac_measured": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9 / peak, 4),
                      "frac_nominal_8tbs": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9 / 8000.0, 4)}
             for k, ms in zip(ks, per_k_ms)}

    def guarded(name, fn):
        """an evidence block must never take the headline number down with it"""
        if name in skip:
            return
        try:
            t0 = time.perf_counter()
            line_extra[name] = fn()
            if isinstance(line_extra[name], dict):
                line_extra[name]["block_seconds"] = round(time.perf_counter() - t0, 1)
        except Exception as e:  # pragma: no cover
            import traceback
            traceback.print_exc(file=sys.stderr)
            line_extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.synchronize()

    guarded("parity", lambda: block_parity(cx, plans, ks, d_in, d_out, n, first, left, halo_elems))
    guarded("dense_k", lambda: block_dense_k(cx, d_in, d_out, n, first, left, peak))
    info = plans[ks[-1]].info
    kernel_cfg = {"threads": info.threads, "run": info.run, "tile_samples": info.tile_samples,
                  "stages": info.stages, "grid": info.grid, "smem_bytes": info.smem_bytes}
    for p in plans.values():
        p.close()
    halo_mode = left.mode
    left.close()
    barrier()
    cx.free(d_in)
    cx.free(d_out)

    guarded("i16", lambda: block_i16(cx, ks, args.samples_log2, peak))
    guarded("configs", lambda: block_configs(cx, args.configs_log2, peak))
    guarded("shapes", lambda: block_shapes(cx, peak))

    # ---------------- end to end through the C ABI with HOST buffers (pinned), copies inside the timing
    e2e = None
    if args.e2e_steps > 0 and "e2e" not in skip:
        try:
            e2e = block_e2e(cx, ks, args.samples_log2, args.e2e_steps)
        except Exception as e:  # pragma: no cover
            import traceback
            traceback.print_exc(file=sys.stderr)
            e2e = None
            line_extra["e2e_error"] = "%s: %s" % (type(e).__name__, e)

    traffic, traffic_src = ncu_traffic_bytes()
    line = {
        "metric": metric_name(args.samples_log2, ks),
        "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_block(args.samples_log2, world, ks),
        "run": {"halo": halo_mode,
                "launch_mode": ("CUDA graph replay, one graph = one step of %d kernel nodes" % len(ks)) if graph is not None
                else "stream launches",
                "overlap": {1: "programmatic dependent launch, wait before the first load",
                            2: "programmatic dependent launch, tile loads start under the previous kernel's tail "
                               "(input never written), nothing is stored before the previous kernel has completed",
                            3: "off"}.get(args.overlap, str(args.overlap)),
                "l2": "inputs (%.0f MiB) + outputs per launch exceed the 126 MB L2; no flush needed" % (4 * n / 2**20),
                "kernel": kernel_cfg},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "DRAM bytes per launch, ncu --set full, 2^28 samples (%s)" % traffic_src,
                     "peak_source": peak_src, "frac_of_nominal_8tbs": achieved / 8000.0,
                     "kernel": "mavg::stream_f32_kernel (mean launch duration over the timed region: elapsed / launches)",
                     "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n},
        "per_k": per_k,
        "per_k_note": "separate pass after the timed region: %d back-to-back stream launches per k, one event pair per k" % reps,
    }
    line.update(line_extra)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line.update(cpu_baseline_block(ks))
        except Exception as e:  # the CPU yardstick must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "Gsamples/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    os.dup2(2, 1)

    barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
