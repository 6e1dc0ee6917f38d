#!/usr/bin/env python
"""Sweep driver for the drop-in `averager` binaries -- successor of the reference's
basics/run_benchmarks.py.

What is kept: one child process per (binary, input, grade, block size) invoked as
`<bin> <wav_path> <grade> <block_size>` (run_benchmarks.py:86-91), failures counted by exit code
(:93-97), results appended by the children to a CSV.  What changes (BASELINE.json north_star):

  * the script parses (the reference has an unclosed print at :111) and its inputs are SEEDED
    (the reference draws unseeded np.random samples, :38);
  * inputs are canonical 44-byte-header WAVs written by hand, int16 stereo like the reference or
    float32 mono (scipy writes float32 with a 58-byte header that wav_header.h cannot parse);
  * the sweep covers GPU counts 1/2/4/8 (`--gpus`), forwarded to the binaries as `--gpus N`;
  * one CSV name everywhere (benchmark_data.csv; the reference writes benchmark_data.csv but
    announces benchmark_results.csv, :115), with the extra columns GPUs, Dtype, Layout, Gsamples_s,
    HBM_GBs, Pct_HBM_nominal, Pct_HBM_measured filled in by the binaries' CsvLogger;
  * an optional CPU binary (`--cpu-bin`, e.g. the reference's own bin_cpu) is timed beside the GPU
    runs; its core count is recorded in the summary.

Usage:
  python -m digital_signal_processsing_b200.run_benchmarks --sizes 1048576 --grades 3,16,64 --gpus 1
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time

import numpy as np

from . import wav

PKG = os.path.dirname(os.path.abspath(__file__))
BIN_DIR = os.path.join(os.path.dirname(PKG), "host", "bin")

# same table as the reference (run_benchmarks.py:8-18); every GPU name is the one libmavg program
EXECUTABLES = [
    {"name": "Parallel_Avg", "path": "bin_parallel", "needs_block": True},
    {"name": "SharedMem", "path": "bin_shared", "needs_block": True},
    {"name": "Vectorized_int2", "path": "bin_vec2", "needs_block": True},
    {"name": "Vectorized_int4", "path": "bin_vec4", "needs_block": True},
    {"name": "HillisSteele", "path": "bin_hillis", "needs_block": True},
    {"name": "V_HillisSteele", "path": "bin_vhillis", "needs_block": True},
    {"name": "Blelloch", "path": "bin_blelloch", "needs_block": True},
    {"name": "V_Blelloch", "path": "bin_vblelloch", "needs_block": True},
]
BLOCK_SIZES = [32, 64, 128, 256, 512, 1024]
GRADES = list(range(1, 11)) + list(range(11, 51, 5)) + list(range(50, 1001, 50))  # run_benchmarks.py:23
HEADLINE_GRADES = [3, 16, 64, 256, 1024, 4096]                                    # BASELINE.json configs
TEMP_WAV = "temp_bench.wav"
CSV_NAME = "benchmark_data.csv"
SEED = 0x5EED0001


def generate_wav(path: str, num_samples: int, channels: int = 2, dtype: str = "int16", seed: int = SEED) -> None:
    """Seeded random WAV with a canonical header (the reference: stereo int16, run_benchmarks.py:31-49)."""
    rng = np.random.default_rng(seed)
    frames = num_samples // channels
    if dtype == "int16":
        data = rng.integers(-32768, 32767, size=frames * channels, dtype=np.int16, endpoint=True)
    else:
        data = rng.random(frames * channels, dtype=np.float32)
    wav.write_samples(path, wav.make_header(data.size, channels, data.dtype), data)


def hbm_peak_args():
    """--hbm-peak for the binaries' Pct_HBM_measured column: MAVG_HBM_PEAK_GBS, else the driver-written
    MEASURED_PEAKS.json at the repository root, else nothing (the binaries then quote the B200_PROFILING.md figure)."""
    if os.environ.get("MAVG_HBM_PEAK_GBS"):
        return []          # the binaries read the variable themselves
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        import json
        return ["--hbm-peak", str(float(json.load(open(p))["hbm_gbs"]))]
    except Exception:
        return []


def run_suite(sizes, grades, blocks, gpus_list, executables, dtype, channels, cpu_bin, csv_name=CSV_NAME,
              bin_dir=BIN_DIR, verbose=True):
    start, counter, failures = time.time(), 0, 0
    rows = []
    for n_samples in sizes:
        generate_wav(TEMP_WAV, int(n_samples), channels, dtype)
        if cpu_bin:
            for grade in grades:
                if grade >= n_samples // channels:
                    continue
                counter += 1
                r = subprocess.run([cpu_bin, TEMP_WAV, str(grade), "256"], capture_output=True, text=True)
                failures += r.returncode != 0
        for gpus in gpus_list:
            for exe in executables:
                path = os.path.join(bin_dir, exe["path"])
                if not os.path.exists(path):
                    print(f"Binary not found: {path}")
                    continue
                for grade in grades:
                    if grade >= n_samples // channels:     # run_benchmarks.py:78
                        continue
                    for b in (blocks if exe["needs_block"] else [256]):
                        counter += 1
                        cmd = [path, TEMP_WAV, str(grade), str(b), "--gpus", str(gpus), "--csv", csv_name] + hbm_peak_args()
                        r = subprocess.run(cmd, capture_output=True, text=True)
                        rows.append((exe["name"], n_samples, grade, b, gpus, r.returncode))
                        if r.returncode != 0:
                            failures += 1
                            print(f"Failure: {exe['name']} (N={n_samples}, G={grade}, B={b}, GPUs={gpus}) rc={r.returncode}")
                            print(f"Error: {r.stderr.strip()}")
                        if verbose and counter % 50 == 0:
                            print(f"{counter} runs. Elapsed: {time.time() - start:.1f}s")
    if os.path.exists(TEMP_WAV):
        os.remove(TEMP_WAV)
    print("\n__________________________________")
    print("BENCHMARK COMPLETE")
    print(f"Total Runs: {counter}")
    print(f"Total Failures/Crashes: {failures}")
    print(f"Host cores: {os.cpu_count()}")
    print(f"Results saved to: {csv_name}")
    print("__________________________________\n")
    return counter, failures, rows


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--sizes", default="1048576", help="comma list of total sample counts")
    ap.add_argument("--grades", default=",".join(map(str, HEADLINE_GRADES)), help="comma list, or 'reference'")
    ap.add_argument("--blocks", default="256", help="comma list, or 'reference'")
    ap.add_argument("--gpus", default="1", help="comma list of GPU counts, e.g. 1,2,4,8")
    ap.add_argument("--bins", default="bin_vec4", help="comma list of binary names, or 'all'")
    ap.add_argument("--dtype", default="int16", choices=["int16", "float32"])
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--cpu-bin", default="", help="optional CPU averager binary to run beside the GPU ones")
    ap.add_argument("--csv", default=CSV_NAME)
    a = ap.parse_args(argv)
    grades = GRADES if a.grades == "reference" else [int(v) for v in a.grades.split(",")]
    blocks = BLOCK_SIZES if a.blocks == "reference" else [int(v) for v in a.blocks.split(",")]
    exes = EXECUTABLES if a.bins == "all" else [e for e in EXECUTABLES if e["path"] in a.bins.split(",")]
    _, failures, _ = run_suite([int(v) for v in a.sizes.split(",")], grades, blocks,
                               [int(v) for v in a.gpus.split(",")], exes, a.dtype, a.channels, a.cpu_bin, a.csv)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
