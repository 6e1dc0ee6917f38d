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
