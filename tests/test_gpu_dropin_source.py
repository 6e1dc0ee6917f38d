"""Source-level drop-in proof (`make -C oracle dropin`).

oracle/_ref/dropin/bin_vec4_mavg is the reference's OWN basics/profilable_sm_vload4.cu -- profiler, CSV logger, main(),
WAV reader and writer untouched -- with its kernel deleted and the body of vload4AveragerGpuLoad replaced by one call
into libmavg's C ABI (host/mavg_dropin.h).  oracle/_ref/dropin/bin_cpu_write is the reference CPU program
(basics/profilable_moving_averager.cpp) with its missing brace and a writeSamples call.  Both read the same WAV with the
reference's reader and write with the reference's writer; the files must be identical byte for byte."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
GPU_BIN = os.path.join(DROPIN, "bin_vec4_mavg")
CPU_BIN = os.path.join(DROPIN, "bin_cpu_write")


def _build():
    """(re)build where the reference tree exists; on the GPU box the prebuilt pair travels with the snapshot"""
    if os.path.isdir("/root/reference"):
        from digital_signal_processsing_b200 import build
        build.build_lib()
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"], check=True)
    if not (os.path.exists(GPU_BIN) and os.path.exists(CPU_BIN)):
        pytest.skip("oracle/_ref/dropin not built (no reference tree here)")


def _wav(tmp_path, oracle_mod, frames, channels, seed):
    from digital_signal_processsing_b200 import wav
    x = oracle_mod.fill_i16(frames * channels, seed)
    src = tmp_path / "in.wav"
    wav.write_samples(str(src), wav.make_header(x.size, channels, np.int16), x)
    return x, src


def test_reference_cpu_program_writes_the_oracle_output(oracle_mod, tmp_path):
    """CPU half of the pair (runs without a GPU): the patched reference CPU binary's file equals header + oracle."""
    _build()
    from digital_signal_processsing_b200 import wav
    x, src = _wav(tmp_path, oracle_mod, 30_000, 2, 7)
    r = subprocess.run([CPU_BIN, str(src), "41", "256"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    h, y = wav.extract_samples(str(tmp_path / "profile_cpu_averager.wav"))
    assert h.pack() == wav.make_header(x.size, 2, np.int16).pack()
    assert np.array_equal(y, oracle_mod.mavg_i16(x, 41, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("channels,frames,k,block", [(2, 400_000, 5, 256), (2, 400_000, 41, 128), (1, 300_001, 1000, 1024),
                                                     (2, 1 << 20, 4096, 256), (6, 50_000, 64, 32)])
def test_reference_vload4_binary_on_libmavg_matches_reference_cpu_binary(oracle_mod, tmp_path, channels, frames, k, block):
    _build()
    x, src = _wav(tmp_path, oracle_mod, frames, channels, 1000 + k)
    g = subprocess.run([GPU_BIN, str(src), str(k), str(block)], cwd=tmp_path, capture_output=True, text=True)
    assert g.returncode == 0, g.stdout + g.stderr
    # the reference's own profiler ran: both memory modes reported, CSV row written by its CsvLogger
    assert "MEM MODE: STANDARD" in g.stdout and "Kernel Compute" in g.stdout
    assert (tmp_path / "benchmark_data.csv").exists()
    c = subprocess.run([CPU_BIN, str(src), str(k), "256"], cwd=tmp_path, capture_output=True, text=True)
    assert c.returncode == 0, c.stdout + c.stderr
    a = (tmp_path / "profile_sm_averager.wav").read_bytes()
    b = (tmp_path / "profile_cpu_averager.wav").read_bytes()
    assert len(a) == 44 + 2 * x.size
    assert a == b, "reference GPU binary on libmavg and reference CPU binary disagree"
