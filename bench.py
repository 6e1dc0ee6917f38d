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
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KS_DEFAULT = [3, 16, 64, 256, 1024, 4096]
SEED = 0x5EED0001
BYTES_PER_SAMPLE = 8  # 4 B read + 4 B written


def parse_args():
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
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of replaying a CUDA graph of one step")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the
    committed `ncu --set full` capture of this same workload (profiles/r01/ncu_stream_full.csv, three launches:
    k = 3, 64, 4096); None when the summary is missing.  Never measured under the profiler at bench time."""
    p = os.path.join(ROOT, "profiles", "r01", "ncu_stream_full.csv")
    try:
        import csv
        rows = {r[0]: r for r in csv.reader(open(p))}
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        total = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            r = rows[key]
            vals = [float(v) for v in r[2:]]
            total += scale[r[1]] * sum(vals) / len(vals)
        return total
    except Exception:
        return None


# ---------------------------------------------------------------------------- reference arm
def metric_name(samples_log2, ks):
    """BASELINE.json's metric on its named configuration; both arms print the same string."""
    return "Gsamples/s, 2^%d-sample mono float32 moving average per GPU, k sweep %s" % (samples_log2, ks)


def workload_name(samples_log2, world, ks):
    n = 1 << samples_log2
    return ("mono float32 synthetic U[0,1) signal, 2^%d samples per GPU (contiguous shards of one %d-sample signal), "
            "window sweep k=%s, device resident" % (samples_log2, world * n, ks))


def cpu_reference_line(args, ks, world):
    """The reference's own CPU implementation (oracle/_ref, built from /root/reference) timed on the
    host cores.  It is single threaded by construction (basics/profilable_moving_averager.cpp:14-37),
    so cores = 1; samples are int16 because wav_header.h:34 admits nothing else."""
    import numpy as np
    import oracle
    n = 1 << 24  # bounded sample per k: 2^24 int16 samples (~75 ms each at 0.22 Gsamples/s)
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
    total = 0.0
    for _ in range(args.steps):
        total += one_step()
    ms = 1e3 * total / max(1, args.steps)
    value = len(ks) * n * args.steps / total / 1e9
    sample = f"mono int16, 2^24 samples x k in {ks} per step, single thread"
    line = {
        "impl": "reference",
        "metric": metric_name(args.samples_log2, ks),
        "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i16", "data": "synthetic",
        "config": {"workload": workload_name(args.samples_log2, world, ks),
                   "reference_arm": "the reference's CPU path profilable_cpu_computations (int16, its only sample type) on "
                                    "a bounded sample of that workload, rank 0 only",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host_cores": os.cpu_count(),
    }
    return line


def cpu_baseline_block(ks):
    """cpu_baseline for the main line: the reference on one thread (it has no threads), plus our
    float32 port on all cores for context.  Bounded: 2^26 int16 samples per k."""
    import numpy as np
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


# ---------------------------------------------------------------------------- our arm
def main():
    args = parse_args()
    ks = [int(v) for v in args.ks.split(",") if v]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(cpu_reference_line(args, ks, max(world, args.gpus))), flush=True)
        return 0

    # Everything except the final JSON line goes to stderr (NCCL prints a version banner on stdout).
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import digital_signal_processsing_b200 as mavg
    from digital_signal_processsing_b200 import _lib, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmavg has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    n = 1 << args.samples_log2            # samples per rank
    first = rank * n                      # global index of this rank's first sample
    tune = dict(kv.split("=") for kv in args.tune.split(",") if kv)
    tune = {k: int(v) for k, v in tune.items()}
    lib = _lib.load()
    import ctypes

    # device buffers from plain cudaMalloc (exportable over CUDA IPC)
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(lib.mavg_device_alloc(4 * n, ctypes.byref(d_in)))
    _lib.check(lib.mavg_device_alloc(4 * n, ctypes.byref(d_out)))
    stream = torch.cuda.Stream()
    mavg.fill_synthetic_device(d_in.value, "f32", n, first, SEED, mavg.DIST_U01, stream.cuda_stream)
    stream.synchronize()

    plans = {}
    for k in ks:
        p = mavg.Plan(n, k, first_frame=first, **tune)
        p.set_stream(stream.cuda_stream)
        p.enable_timing(False)   # bench.py times the stream itself; skip the plan's four event records per run
        plans[k] = p
    max_halo = max(int(p.info.halo_frames) for p in plans.values())

    # left context: read in place from the left neighbour (IPC) or staged once over NCCL
    halo_ptr = {k: 0 for k in ks}
    peer = None
    staged = None
    halo_mode = "none"
    if world > 1:
        if args.halo == "ipc":
            try:
                peer = sharding.PeerHalo(d_in.value, n, 4, max_halo, rank, world)
                halo_mode = "ipc-peer (TMA reads the neighbour's tail in place over NVLink)"
                if rank > 0:
                    for k in ks:
                        h = int(plans[k].info.halo_frames)
                        halo_ptr[k] = peer.halo_ptr + 4 * (max_halo - h)
            except Exception as e:  # pragma: no cover - depends on the box
                if rank == 0:
                    print(f"[bench] CUDA IPC unavailable ({e}); using NCCL send/recv", file=sys.stderr)
                peer = None
        if peer is None:
            shard = torch.empty(0)  # placeholder to keep names defined
            import numpy as np
            # wrap our cudaMalloc'ed buffer as a torch tensor without copying
            class _Arr:  # __cuda_array_interface__ provider
                def __init__(self, ptr, count):
                    self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 3}
            shard = torch.as_tensor(_Arr(d_in.value, n), device="cuda")
            staged = sharding.exchange_halo(shard, max_halo, rank, world)
            torch.cuda.synchronize()
            halo_mode = "nccl send/recv into a staging buffer"
            if rank > 0:
                for k in ks:
                    h = int(plans[k].info.halo_frames)
                    halo_ptr[k] = staged.data_ptr() + 4 * (max_halo - h)
    barrier()

    launches_per_step = 0

    def step(record=None):
        nonlocal launches_per_step
        cnt = 0
        if record is not None:
            record[0].record(stream)
        for i, k in enumerate(ks):
            plans[k].run_device_halo(d_in.value, d_out.value, halo_ptr[k] or None)
            if record is not None:
                record[i + 1].record(stream)   # one event per boundary: launch i is timed from event i to i+1
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

    # ---------------- timed region: EXACTLY K steps, bracketed by events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    with torch.cuda.stream(stream):
        e0.record()
        for s in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                step()
        e1.record()
    sampler.sample()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)

    # per-window detail: a separate, untimed-for-the-headline pass of back-to-back launches of each k
    reps = max(3, min(args.steps, 20))
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ks]
    for i, k in enumerate(ks):
        evk[i][0].record(stream)
        for _ in range(reps):
            plans[k].run_device_halo(d_in.value, d_out.value, halo_ptr[k] or None)
        evk[i][1].record(stream)
    torch.cuda.synchronize()
    per_k_ms = [a.elapsed_time(b) / reps for a, b in evk]
    if world > 1:
        t = torch.tensor([elapsed_ms] + per_k_ms, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, per_k_ms = float(t[0]), [float(v) for v in t[1:]]

    total_samples = world * n * len(ks) * args.steps
    value = total_samples / (elapsed_ms * 1e-3) / 1e9
    peak, peak_src = peaks()
    kernel_ms = elapsed_ms / (args.steps * len(ks))   # mean launch duration over the timed region (events e0..e1)
    achieved = BYTES_PER_SAMPLE * n / (kernel_ms * 1e-3) / 1e9
    per_k = {str(k): {"ms": round(ms, 4), "gsamples_s": round(n / (ms * 1e-3) / 1e9, 1),
                      "hbm_gbs": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9, 1),
                      "frac_measured": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9 / peak, 4),
                      "frac_nominal_8tbs": round(BYTES_PER_SAMPLE * n / (ms * 1e-3) / 1e9 / 8000.0, 4)}
             for k, ms in zip(ks, per_k_ms)}

    # ---------------- end to end through the C ABI with HOST buffers (pinned), copies inside the timing
    e2e = None
    if args.e2e_steps > 0:
        halo_elems = max_halo if rank > 0 else 0
        h_in = torch.empty(halo_elems + n, dtype=torch.float32, pin_memory=True)
        h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
        tmp = torch.empty(halo_elems + n, dtype=torch.float32, device="cuda")
        mavg.fill_synthetic_device(tmp.data_ptr(), "f32", halo_elems + n, first - halo_elems, SEED, mavg.DIST_U01,
                                   torch.cuda.current_stream().cuda_stream)
        h_in.copy_(tmp)
        torch.cuda.synchronize()
        del tmp
        hplans = {k: mavg.Plan(n, k, first_frame=first, **tune) for k in ks}
        in_ptr = h_in.data_ptr() + 4 * halo_elems

        def host_step():
            for k in ks:
                hplans[k].run_host_ptr(in_ptr, h_out.data_ptr())   # H2D + kernel + D2H, blocking

        host_step()  # warm-up: allocates the plan-owned device buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        tm = hplans[ks[-1]].timing()
        e2e = {"value": world * n * len(ks) * args.e2e_steps / dt / 1e9, "unit": "Gsamples/s",
               "h2d_bytes_per_step": 4 * (n + halo_elems) * len(ks), "d2h_bytes_per_step": 4 * n * len(ks),
               "steps": args.e2e_steps, "api": "mavg_run_host (pinned host buffers, H2D + kernel + D2H per k)",
               "last_call_phases_ms": {"h2d": round(tm.h2d_ms, 3), "compute": round(tm.compute_ms, 3),
                                       "d2h": round(tm.d2h_ms, 3)}}
        for p in hplans.values():
            p.close()

    info = plans[ks[-1]].info
    line = {
        "metric": metric_name(args.samples_log2, ks),
        "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.samples_log2, world, ks),
                   "samples_per_gpu": n, "ks": ks, "halo": halo_mode,
                   "launch_mode": ("CUDA graph replay, one graph = one step of %d kernel nodes" % len(ks)) if graph is not None
                   else "stream launches",
                   "l2": "inputs (%.0f MiB) + outputs per launch exceed the 126 MB L2; no flush needed" % (4 * n / 2**20),
                   "kernel": {"threads": info.threads, "run": info.run, "tile_samples": info.tile_samples,
                              "stages": info.stages, "grid": info.grid, "smem_bytes": info.smem_bytes}},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic_bytes(), "traffic_unit": "DRAM bytes per launch, ncu --set full, 2^28 samples "
                     "(profiles/r01/ncu_stream_full.csv)", "peak_source": peak_src, "frac_of_nominal_8tbs": achieved / 8000.0,
                     "kernel": "mavg::stream_f32_kernel (mean launch duration over the timed region: elapsed / launches)",
                     "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n},
        "per_k": per_k,
        "per_k_note": "separate pass after the timed region: %d back-to-back stream launches per k, one event pair per k" % reps,
    }
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

    for p in plans.values():
        p.close()
    if peer is not None:
        peer.close()
    barrier()
    lib.mavg_device_free(d_in)
    lib.mavg_device_free(d_out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
