#!/usr/bin/env python
"""Secondary yardstick: the reference's own GPU binaries, recompiled unchanged for sm_100a
(oracle/Makefile refgpu -> oracle/_ref/gpu/), run on the B200 next to libmavg's drop-in program on the
same stereo int16 WAV.  Reads each program's benchmark_data.csv row (Compute_ms = kernels only).

  python tests/perf/ref_gpu_yardstick.py [--log2 27] [--grades 3,64,1024] [--out gpurun_out/ref_gpu_yardstick.csv]
"""
import argparse
import csv
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from digital_signal_processsing_b200 import build, run_benchmarks  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref", "gpu")
NAMES = ["bin_parallel", "bin_shared", "bin_vec2", "bin_vec4", "bin_hillis", "bin_vhillis", "bin_blelloch", "bin_vblelloch"]


def run_one(exe, wav, grade, block, cwd, extra=()):
    csv_path = os.path.join(cwd, "benchmark_data.csv")
    if os.path.exists(csv_path):
        os.remove(csv_path)
    try:
        r = subprocess.run([exe, wav, str(grade), str(block), *extra], cwd=cwd, capture_output=True, text=True, timeout=300)
    except subprocess.TimeoutExpired:
        return None, "timeout"
    if r.returncode != 0 or not os.path.exists(csv_path):
        return None, f"rc={r.returncode} {r.stderr.strip()[:120]}"
    rows = list(csv.DictReader(open(csv_path)))
    std = [x for x in rows if x.get("MemoryMode", "Standard") == "Standard"]
    return (std or rows)[0], ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=27)
    ap.add_argument("--grades", default="3,64,1024")
    ap.add_argument("--out", default="gpurun_out/ref_gpu_yardstick.csv")
    a = ap.parse_args()
    n = 1 << a.log2
    ours = os.path.join(os.path.dirname(build.build_host()[0]), "averager")
    tmp = tempfile.mkdtemp(prefix="yard_")
    wav = os.path.join(tmp, "in.wav")
    run_benchmarks.generate_wav(wav, n, channels=2, dtype="int16")
    rows = []
    for grade in [int(g) for g in a.grades.split(",")]:
        r, err = run_one(ours, wav, grade, 256, tmp, ("--rounds", "5", "--warmup", "2"))
        rows.append(("libmavg", grade, r, err))
        for name in NAMES:
            exe = os.path.join(REF_DIR, name)
            if not os.path.exists(exe):
                continue
            r, err = run_one(exe, wav, grade, 256, tmp)
            rows.append((f"reference {name}", grade, r, err))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write("program,grade,n_samples,h2d_ms,compute_ms,d2h_ms,total_ms,kernel_gsamples_s,note\n")
        for name, grade, r, err in rows:
            if r is None:
                f.write(f"{name},{grade},{n},,,,,,{err}\n")
                continue
            c = float(r["Compute_ms"])
            f.write(f"{name},{grade},{r['N_Samples']},{r['H2D_ms']},{r['Compute_ms']},{r['D2H_ms']},{r['Total_ms']},"
                    f"{(n / c / 1e6) if c > 0 else 0:.2f},\n")
    print(open(a.out).read())
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
