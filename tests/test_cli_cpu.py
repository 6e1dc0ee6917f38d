"""The drop-in `averager` program's argument handling and exit codes (no GPU needed: all of these end before any
device work), plus: on a box without a GPU a valid WAV fails loudly instead of falling back to anything."""
import os
import subprocess

import numpy as np
import pytest


@pytest.fixture(scope="module")
def bin_dir(mavg):
    from digital_signal_processsing_b200 import build
    bins = build.build_host()
    assert bins, "host/averager_main.cpp missing"
    return os.path.dirname(bins[0])


def _run(bin_dir, name, *args, cwd):
    return subprocess.run([os.path.join(bin_dir, name)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)


def test_usage_block_size_and_unreadable_input(bin_dir, tmp_path):
    r = _run(bin_dir, "bin_vec4", cwd=tmp_path)
    assert r.returncode == 1 and "Usage" in r.stderr                                       # basics/profilable_sm_vload4.cu:221-226
    for bad in (100, 16, 2048):
        r = _run(bin_dir, "bin_hillis", "x.wav", 5, bad, cwd=tmp_path)
        assert r.returncode == 1 and "Block size must be multiple of 32" in r.stderr       # :231-234
    assert _run(bin_dir, "bin_vec4", "x.wav", 0, 256, cwd=tmp_path).returncode == 1         # grade must be >= 1
    assert _run(bin_dir, "bin_vec4", "missing.wav", 5, 256, cwd=tmp_path).returncode == 2   # fixed: the reference exits 0
    junk = tmp_path / "junk.wav"
    junk.write_bytes(b"not a wav file at all, but longer than forty-four bytes for sure......")
    r = _run(bin_dir, "bin_vec4", junk, 5, 256, cwd=tmp_path)
    assert r.returncode == 2 and "RIFF" in r.stdout
    assert not (tmp_path / "benchmark_data.csv").exists()                                   # nothing was measured
    assert _run(bin_dir, "averager", "x.wav", 5, 256, "--gpus", 99, cwd=tmp_path).returncode == 1


def test_all_reference_binary_names_resolve(bin_dir):
    for name in ("bin_parallel", "bin_shared", "bin_vec2", "bin_vec4", "bin_hillis", "bin_vhillis", "bin_blelloch",
                 "bin_vblelloch"):                                                          # basics/run_benchmarks.py:8-18
        assert os.path.exists(os.path.join(bin_dir, name)), name


def test_fails_loudly_without_a_gpu(bin_dir, mavg, tmp_path):
    if mavg.device_count() > 0:
        pytest.skip("a GPU is present")
    from digital_signal_processsing_b200 import wav
    x = np.arange(-2000, 2000, dtype=np.int16)
    wav.write_samples(str(tmp_path / "ok.wav"), wav.make_header(x.size, 2, np.int16), x)
    r = _run(bin_dir, "bin_vec4", "ok.wav", 5, 256, cwd=tmp_path)
    assert r.returncode != 0 and "libmavg" in (r.stderr + r.stdout)
    assert not (tmp_path / "benchmark_data.csv").exists()


def test_csv_logger_keeps_the_reference_schema(mavg, tmp_path):
    """CsvLogger (host/mavg_workspace.h, reached through the compat name gpu_utils.h): the first 14 columns are the
    reference's (gpu_utils.h:196-199), in order, with the same formulas (:202-227); the new columns follow."""
    from digital_signal_processsing_b200 import _lib
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    src = tmp_path / "csv.cpp"
    src.write_text('''
#include "gpu_utils.h"
#include "benchmark.h"
int main() {
    ProfileResult r{};
    r.transfer_h2d_ms = 1.0f; r.compute_ms = 0.5f; r.transfer_d2h_ms = 1.5f; r.total_ms = 3.0f; r.initialization_ms = 7.0f;
    CsvLogger("out.csv").log("Vectorized_int4", "Standard", 1000000, 64, 256, r, 2, 2, 1, "int16", "interleaved");
    CsvLogger("out.csv").log("Vectorized_int4", "Standard", 1000000, 64, 256, r, 2, 2, 1, "int16", "interleaved");
    return 0;
}''')
    exe = tmp_path / "csv"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "host"), str(src),
                    "-o", str(exe), "-L", libdir, "-lmavg", f"-Wl,-rpath,{libdir}", "-lpthread"], check=True)
    subprocess.run([str(exe)], cwd=tmp_path, check=True, capture_output=True)
    lines = (tmp_path / "out.csv").read_text().strip().split("\n")
    assert len(lines) == 3                                             # one header, two rows (appended)
    cols = lines[0].split(",")
    assert cols[:14] == ["Algorithm", "MemoryMode", "N_Samples", "Grade", "BlockSize", "H2D_ms", "Compute_ms", "D2H_ms",
                         "Total_ms", "Init_ms", "ColdStart_Total_ms", "Bandwidth_GBs", "Throughput_MSs", "ColdStart_MSs"]
    assert cols[14:] == ["GPUs", "Dtype", "Layout", "Gsamples_s", "HBM_GBs", "Pct_HBM_nominal", "Pct_HBM_measured"]
    row = dict(zip(cols, lines[1].split(",")))
    assert row["Algorithm"] == "Vectorized_int4" and row["N_Samples"] == "1000000" and row["Grade"] == "64"
    assert float(row["ColdStart_Total_ms"]) == 10.0
    # six significant digits, like the reference's default ostream precision
    assert abs(float(row["Bandwidth_GBs"]) - 1e6 * 4 / 1e9 / 3e-3) < 1e-4          # N * (in + out bytes) / total time
    assert abs(float(row["Throughput_MSs"]) - 1.0 / 3e-3) < 1e-3                   # Msamples per second, steady state
    assert abs(float(row["ColdStart_MSs"]) - 1.0 / 10e-3) < 1e-3
    assert abs(float(row["Gsamples_s"]) - 1e6 / 1e9 / 0.5e-3) < 1e-4               # kernel-only


def test_sweep_driver_counts_runs_and_failures(bin_dir, mavg, monkeypatch, tmp_path, capsys):
    """The run_benchmarks.py successor on a box without a GPU: inputs are seeded (two generations are identical),
    grades >= frames are skipped like the reference does (run_benchmarks.py:78), every GPU child fails loudly and
    is counted as a failure, the exit status says so.  With the reference's own CPU binary beside it (when
    oracle/_ref was built here) those runs succeed and write the reference's CSV."""
    if mavg.device_count() > 0:
        pytest.skip("a GPU is present: the GPU children would succeed (covered by tests/test_gpu_dropin.py)")
    from digital_signal_processsing_b200 import run_benchmarks as rb
    monkeypatch.chdir(tmp_path)
    rb.generate_wav("a.wav", 4000, 2, "int16")
    rb.generate_wav("b.wav", 4000, 2, "int16")
    assert (tmp_path / "a.wav").read_bytes() == (tmp_path / "b.wav").read_bytes()
    ref_cpu = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "bin_cpu")
    cpu_bin = os.path.abspath(ref_cpu) if os.path.exists(ref_cpu) else ""
    exes = [e for e in rb.EXECUTABLES if e["path"] in ("bin_vec4", "bin_hillis")]
    counter, failures, rows = rb.run_suite([4000], [3, 64, 5000], [128, 256], [1], exes, "int16", 2, cpu_bin, verbose=False)
    gpu_runs = 2 * 2 * 2                                   # binaries x grades below 2000 frames x block sizes
    assert len(rows) == gpu_runs and all(r[-1] != 0 for r in rows)
    assert counter == gpu_runs + (2 if cpu_bin else 0) and failures == gpu_runs
    assert not os.path.exists(rb.TEMP_WAV)
    assert "Total Failures/Crashes: %d" % gpu_runs in capsys.readouterr().out
    assert rb.main(["--sizes", "4000", "--grades", "3", "--bins", "bin_vec4"]) == 1
