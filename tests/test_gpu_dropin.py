"""GPU tests of the drop-in `averager` program (host/averager_main.cpp over the C ABI): same argv,
same WAV in, same output layout as the reference binaries; outputs diffed against the oracle."""
import csv
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def averager(mavg):
    from digital_signal_processsing_b200 import build
    bins = build.build_host()
    assert bins, "host/averager_main.cpp missing"
    return os.path.dirname(bins[0])


def _run(bin_dir, name, *args, cwd):
    return subprocess.run([os.path.join(bin_dir, name)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)


def test_int16_stereo_wav_bit_exact(averager, mavg, oracle_mod, tmp_path):
    from digital_signal_processsing_b200 import wav
    x = oracle_mod.fill_i16(2 * 50_000, 42)
    src = tmp_path / "in.wav"
    wav.write_samples(str(src), wav.make_header(x.size, 2, np.int16), x)
    for name, k in (("bin_vec4", 5), ("bin_hillis", 1000), ("averager", 41)):
        out = tmp_path / f"out_{k}.wav"
        r = _run(averager, name, src, k, 256, "--out", out, "--rounds", 2, "--warmup", 1, cwd=tmp_path)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "total samples: 100000" in r.stdout and f"point: {k}" in r.stdout
        h, y = wav.extract_samples(str(out))
        assert h.pack() == wav.make_header(x.size, 2, np.int16).pack()        # header verbatim
        assert np.array_equal(y, oracle_mod.mavg_i16(x, k, 2))
    rows = list(csv.DictReader(open(tmp_path / "benchmark_data.csv")))
    assert [r["Algorithm"] for r in rows] == ["Vectorized_int4", "HillisSteele", "libmavg"]
    first = rows[0]
    for col in ("MemoryMode", "N_Samples", "Grade", "BlockSize", "H2D_ms", "Compute_ms", "D2H_ms", "Total_ms", "Init_ms",
                "ColdStart_Total_ms", "Bandwidth_GBs", "Throughput_MSs", "ColdStart_MSs", "GPUs", "Dtype", "Layout",
                "Gsamples_s", "HBM_GBs", "Pct_HBM_nominal", "Pct_HBM_measured"):
        assert col in first
    assert first["N_Samples"] == "100000" and first["Grade"] == "5" and first["GPUs"] == "1" and first["Dtype"] == "int16"
    assert float(first["Compute_ms"]) > 0


def test_float32_mono_wav(averager, mavg, oracle_mod, tmp_path):
    from digital_signal_processsing_b200 import wav
    n, k = 1 << 20, 5                                  # BASELINE.json configs[0]
    x = oracle_mod.fill_f32(n, 0x5EED0001)
    src, out = tmp_path / "in.wav", tmp_path / "out.wav"
    wav.write_samples(str(src), wav.make_header(n, 1, np.float32), x)
    r = _run(averager, "bin_shared", src, k, 128, "--out", out, "--rounds", 2, "--warmup", 1, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "TMA stream" in r.stdout
    _, y = wav.extract_samples(str(out))
    e = oracle_mod.mavg_f64(x, k)
    assert np.max(np.abs(y - e) / np.abs(e)) < 1e-5


def test_scipy_float32_stereo_wav(averager, mavg, oracle_mod, tmp_path):
    """A float32 WAV as scipy writes it (18-byte fmt + fact chunk: 58-byte header) goes straight through the
    drop-in program; the output file carries a canonical 44-byte header."""
    sciwav = pytest.importorskip("scipy.io.wavfile")
    from digital_signal_processsing_b200 import wav
    frames, k = 70_001, 64
    x = oracle_mod.fill_f32(2 * frames, 4242)
    src, out = tmp_path / "sci.wav", tmp_path / "out.wav"
    sciwav.write(str(src), 44100, x.reshape(-1, 2))
    r = _run(averager, "bin_vec2", src, k, 64, "--out", out, "--rounds", 2, "--warmup", 1, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "channels: 2" in r.stdout and "TMA stream" in r.stdout
    h, y = wav.extract_samples(str(out))
    assert (h.fmtSize, h.audioFormat, h.numChannels, h.bitsPerSample) == (16, 3, 2, 32)
    e = oracle_mod.mavg_f64(x, k, 2)
    assert np.max(np.abs(y - e) / np.abs(e)) < 1e-5


def test_six_channel_float32_wav(averager, mavg, oracle_mod, tmp_path):
    from digital_signal_processsing_b200 import wav
    frames, k = 40_000, 33
    x = oracle_mod.fill_f32(6 * frames, 777)
    src, out = tmp_path / "six.wav", tmp_path / "six_out.wav"
    wav.write_samples(str(src), wav.make_header(x.size, 6, np.float32), x)
    r = _run(averager, "averager", src, k, 256, "--out", out, "--rounds", 2, "--warmup", 1, cwd=tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "channels: 6" in r.stdout and "TMA stream" in r.stdout
    _, y = wav.extract_samples(str(out))
    e = oracle_mod.mavg_f64(x, k, 6)
    assert np.max(np.abs(y - e) / np.abs(e)) < 1e-5


def test_exit_codes(averager, tmp_path):
    assert _run(averager, "bin_vec4", cwd=tmp_path).returncode == 1                       # usage
    r = _run(averager, "bin_vec4", "x.wav", 5, 100, cwd=tmp_path)
    assert r.returncode == 1 and "Block size must be multiple of 32" in r.stderr           # reference rule
    assert _run(averager, "bin_vec4", "missing.wav", 5, 256, cwd=tmp_path).returncode == 2  # fixed: reference exits 0
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wav file at all, but longer than forty-four bytes for sure......")
    assert _run(averager, "bin_vec4", bad, 5, 256, cwd=tmp_path).returncode == 2


def test_sweep_driver(averager, mavg, tmp_path, monkeypatch):
    from digital_signal_processsing_b200 import run_benchmarks as rb
    monkeypatch.chdir(tmp_path)
    total, failures, rows = rb.run_suite([40_000], [3, 64], [64, 256], [1], [e for e in rb.EXECUTABLES if e["path"] in ("bin_vec4", "bin_blelloch")],
                                         "int16", 2, "", verbose=False)
    assert total == 8 and failures == 0
    got = list(csv.DictReader(open(tmp_path / "benchmark_data.csv")))
    assert len(got) == 8 and {r["Algorithm"] for r in got} == {"Vectorized_int4", "Blelloch"}


def test_single_process_multi_device_plan(mavg, oracle_mod):
    """One process driving several GPUs (the --gpus N mode of the binaries); with one GPU visible the
    same device is listed twice, which exercises sharding + in-place peer halo on real hardware."""
    ndev = mavg.device_count()
    devs = [0, 1] if ndev >= 2 else [0, 0]
    n, T = 40 * 8192 + 96, 8192
    for k in (3, 300, 4096):
        x = oracle_mod.fill_f32(n, 77 + k)
        with mavg.Plan(n, k, devices=devs) as plan:
            i = plan.info
            assert i.num_devices == 2 and i.shard_frames[0] % T == 0 and i.shard_frames[0] + i.shard_frames[1] == n
            y2 = plan.run_host(x)
        with mavg.Plan(n, k) as plan:
            y1 = plan.run_host(x)
        assert np.array_equal(y1, y2), k                     # sharding does not change a single bit
        e = oracle_mod.mavg_f64(x, k)
        assert np.max(np.abs(y1 - e) / np.abs(e)) < 1e-5
    xi = oracle_mod.fill_i16(3 * 30_000, 5)
    with mavg.Plan(30_000, 700, channels=3, dtype="i16", devices=devs) as plan:
        assert np.array_equal(plan.run_host(xi), oracle_mod.mavg_i16(xi, 700, 3))
    # every frame-sharded fast path: stereo float32, stereo int16 (stream), 64-channel interleaved (column kernel)
    xs = oracle_mod.fill_f32(2 * (20 * 4096 + 77), 8)
    with mavg.Plan(20 * 4096 + 77, 300, channels=2, devices=devs) as plan:
        ys2 = plan.run_host(xs)
    with mavg.Plan(20 * 4096 + 77, 300, channels=2) as plan:
        assert np.array_equal(plan.run_host(xs), ys2)
    xi2 = oracle_mod.fill_i16(2 * (9 * 8192 + 33), 9)
    with mavg.Plan(9 * 8192 + 33, 1000, channels=2, dtype="i16", devices=devs) as plan:
        assert plan.info.path == 1
        assert np.array_equal(plan.run_host(xi2), oracle_mod.mavg_i16(xi2, 1000, 2))
    xc = oracle_mod.fill_f32(64 * (30 * 128 + 5), 10)
    with mavg.Plan(30 * 128 + 5, 64, channels=64, devices=devs) as plan:
        assert plan.info.mode == 3
        yc2 = plan.run_host(xc)
    with mavg.Plan(30 * 128 + 5, 64, channels=64) as plan:
        assert np.array_equal(plan.run_host(xc), yc2)
    e = oracle_mod.mavg_f64(xc, 64, 64)
    assert np.max(np.abs(yc2 - e) / np.abs(e)) < 1e-5
    x6 = oracle_mod.fill_f32(6 * (50 * 1280 + 17), 12)        # 5.1 audio: six interleaved channels (few-channel kernel)
    with mavg.Plan(50 * 1280 + 17, 100, channels=6, devices=devs) as plan:
        assert plan.info.mode == 4
        y6 = plan.run_host(x6)
    e6 = oracle_mod.mavg_f64(x6, 100, 6)
    assert np.max(np.abs(y6 - e6) / np.abs(e6)) < 1e-5
    # few-channel kernels with windows that reach several tiles back (prefix mode), float32 and int16 pairs
    with mavg.Plan(50 * 1280 + 17, 2000, channels=6, devices=devs) as plan:
        assert plan.info.mode == 4 and plan.info.history_tiles >= 2
        y6l = plan.run_host(x6)
    with mavg.Plan(50 * 1280 + 17, 2000, channels=6) as plan:
        assert np.array_equal(plan.run_host(x6), y6l)
    e6 = oracle_mod.mavg_f64(x6, 2000, 6)
    assert np.max(np.abs(y6l - e6) / np.abs(e6)) < 1e-5
    xi6 = oracle_mod.fill_i16(6 * (50 * 1280 + 17), 13)
    for k in (100, 1500):
        with mavg.Plan(50 * 1280 + 17, k, channels=6, dtype="i16", devices=devs) as plan:
            assert plan.info.mode == 6          # the flat-stream int16 kernel (3 / 4 / 6 / 8 channels)
            assert np.array_equal(plan.run_host(xi6), oracle_mod.mavg_i16(xi6, k, 6)), k
    xp = oracle_mod.fill_f32(6 * 8192 * 3, 6)
    with mavg.Plan(8192 * 3, 64, channels=6, layout="planar", devices=devs) as plan:
        yp = plan.run_host(xp)
    for c in range(6):
        seg = slice(c * 8192 * 3, (c + 1) * 8192 * 3)
        e = oracle_mod.mavg_f64(xp[seg], 64)
        assert np.max(np.abs(yp[seg] - e) / np.abs(e)) < 1e-5


def test_run_host_pipeline_slices_bit_identical(mavg, oracle_mod):
    """Large host buffers go through the sliced H2D/kernel/D2H pipeline; slices are whole tiles with
    their left context read from the already-uploaded previous slice, so bits do not change."""
    n = (1 << 25) + 8192 * 3 + 40           # > 2 slices of 32 MiB, ragged tail
    for k in (5, 1024):
        x = oracle_mod.fill_f32(n, 31 + k)
        with mavg.Plan(n, k) as plan:
            y = plan.run_host(x)
            assert plan.info.launches_per_run >= 3
            t = plan.timing()
            assert t.total_ms > 0
        import torch
        dx = torch.from_numpy(x).cuda()
        dy = torch.empty_like(dx)
        torch.cuda.synchronize()
        with mavg.Plan(n, k) as plan:
            plan.run_device([dx.data_ptr()], [dy.data_ptr()])
            plan.synchronize()
        assert np.array_equal(y, dy.cpu().numpy())
    xi = oracle_mod.fill_i16(2 * (1 << 24) + 2, 9)
    assert np.array_equal(mavg.moving_average(xi, 77, channels=2), oracle_mod.mavg_i16(xi, 77, 2))


@pytest.mark.gpu
def test_box_filter_cascade(mavg, oracle_mod):
    """mavg_run_cascade: the filter applied p times on the device == the oracle applied p times (int16 truncates
    after every pass, exactly like running the reference on its own output); sharded plans order the passes across
    devices and stay bit-identical to the single-device run."""
    import torch
    ndev = mavg.device_count()
    devs = [0, 1] if ndev >= 2 else [0, 0]
    frames = 24 * 8192 + 100
    for dtype, ch, k in (("f32", 1, 33), ("f32", 2, 300), ("i16", 2, 7), ("i16", 1, 1000), ("f32", 6, 50)):
        n = frames * ch
        tdt = torch.float32 if dtype == "f32" else torch.int16
        x = oracle_mod.fill_f32(n, 41 + k) if dtype == "f32" else oracle_mod.fill_i16(n, 41 + k)
        dx = torch.from_numpy(x).cuda()
        for passes in (1, 2, 3, 4):
            e = x
            for _ in range(passes):
                e = oracle_mod.mavg_f64(e, k, ch).astype(np.float32) if dtype == "f32" else oracle_mod.mavg_i16(e, k, ch)
            dy = torch.zeros(n, dtype=tdt, device="cuda")
            with mavg.Plan(frames, k, channels=ch, dtype=dtype) as plan:
                plan.run_cascade([dx.data_ptr()], [dy.data_ptr()], passes)
                plan.synchronize()
                assert plan.timing().compute_ms > 0
            y1 = dy.cpu().numpy()
            if dtype == "f32":
                assert np.max(np.abs(y1 - e) / np.abs(e)) < 1e-5, (dtype, ch, k, passes)
            else:
                assert np.array_equal(y1, e), (dtype, ch, k, passes)
            # two shards (two devices, or the same device twice): passes are ordered across devices
            with mavg.Plan(frames, k, channels=ch, dtype=dtype, devices=devs) as plan:
                f0 = int(plan.info.shard_frames[0])
                bufs = []
                for r, dev in enumerate(devs):
                    lo, hi = (0, f0 * ch) if r == 0 else (f0 * ch, n)
                    with torch.cuda.device(dev):
                        bufs.append((torch.from_numpy(x[lo:hi]).cuda(dev), torch.zeros(hi - lo, dtype=tdt, device=f"cuda:{dev}")))
                torch.cuda.synchronize()
                plan.run_cascade([b[0].data_ptr() for b in bufs], [b[1].data_ptr() for b in bufs], passes)
                plan.synchronize()
                y2 = np.concatenate([b[1].cpu().numpy() for b in bufs])
            assert np.array_equal(y1, y2), (dtype, ch, k, passes)
    with mavg.Plan(1000, 5) as plan:
        with pytest.raises(Exception):
            plan.run_cascade([dx.data_ptr()], [dx.data_ptr() + 4096], 0)


# ------------------------------------------------------------------ mavg_run_host_sweep: one upload, many windows
@pytest.mark.parametrize("case", [("f32", 1, (1 << 22) + 37, [3, 16, 64, 256, 1024, 4096]),
                                  ("i16", 2, (1 << 21) + 5, [3, 64, 4096, 30_000]),          # 30 000: far-lag int16 kernel
                                  ("i16", 6, 300_000, [2, 64, 1000, 9000]),
                                  ("f32", 3, 200_000, [5, 300, 20_000]),                      # 20 000: prefix sum + difference
                                  ("f32", 1, 1 << 21, [64, 60_000])])                         # 60 000: far-lag float32 kernel
def test_run_host_sweep_matches_one_call_per_window(mavg, oracle_mod, case):
    """The sweep call uploads the input once and runs every plan on each slice; its outputs have to be bit-identical
    to one mavg_run_host per plan (the slices start on tile boundaries of every plan) -- except for the float32 far-lag
    kernel, whose carried window sum starts from a different warm-up in every slice (1e-6)."""
    dtype, ch, frames, ks = case
    n = frames * ch
    x = oracle_mod.fill_f32(n, 51_000 + frames % 1000) if dtype == "f32" else oracle_mod.fill_i16(n, 51_000 + frames % 1000)
    plans = [mavg.Plan(frames, k, channels=ch, dtype=dtype, slice_bytes=1 << 20) for k in ks]
    try:
        single = [p.run_host(x).copy() for p in plans]
        xin = mavg.PinnedArray(n, x.dtype)
        xin.array[:] = x
        outs = [mavg.PinnedArray(n, x.dtype) for _ in ks]
        mavg.run_host_sweep_ptr(plans, xin.array.ctypes.data, [o.array.ctypes.data for o in outs])
        for k, p, s, o in zip(ks, plans, single, outs):
            if dtype == "f32" and int(p.info.mode) == 5:
                e = oracle_mod.mavg_f64(x, k, ch)
                assert float(np.max(np.abs(o.array - s) / np.abs(e))) < 1e-6, k
            else:
                assert np.array_equal(o.array, s), k
        # pageable outputs take the per-plan route and give the same result
        got = mavg.run_host_sweep(plans, x)
        for k, p, s, g in zip(ks, plans, single, got):
            if not (dtype == "f32" and int(p.info.mode) == 5):
                assert np.array_equal(g, s), k
        t = plans[0].timing()
        assert t.total_ms > 0
    finally:
        for p in plans:
            p.close()


def test_run_host_sweep_shard_plans_and_argument_checks(mavg, oracle_mod):
    """Shard plans (first_frame > 0): the longest halo of the sweep sits in front of h_in; every plan finds its own
    shorter halo right in front of the shard.  Bad arguments are refused with a status code."""
    frames, cut, ch = (1 << 20) + 11, 300_000, 2
    ks = [5, 300, 4096]
    x = oracle_mod.fill_i16(frames * ch, 52_000)
    plans = [mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut, slice_bytes=1 << 19) for k in ks]
    try:
        halo = max(int(p.info.halo_frames) for p in plans)
        assert halo <= cut
        buf = mavg.PinnedArray((frames - cut + halo) * ch, np.int16)
        buf.array[:] = x[(cut - halo) * ch:]
        outs = [mavg.PinnedArray((frames - cut) * ch, np.int16) for _ in ks]
        mavg.run_host_sweep_ptr(plans, buf.array.ctypes.data + 2 * halo * ch, [o.array.ctypes.data for o in outs])
        for k, o in zip(ks, outs):
            assert np.array_equal(o.array, oracle_mod.mavg_i16(x, k, ch)[cut * ch:]), k
        with pytest.raises(mavg.MavgError):
            mavg.run_host_sweep_ptr([plans[0], plans[0]], buf.array.ctypes.data, [outs[0].array.ctypes.data, outs[1].array.ctypes.data])
        with pytest.raises(mavg.MavgError):
            mavg.run_host_sweep_ptr(plans[:2], buf.array.ctypes.data, [outs[0].array.ctypes.data, outs[0].array.ctypes.data])
    finally:
        for p in plans:
            p.close()
