"""GPU parity tests: libmavg's CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): int16 bit-exact with the reference CPU function
(basics/profilable_moving_averager.cpp:14-37); float32 max relative error <= 1e-5 against
the fp64 oracle on U[0,1) input; zero-mean input is judged by the forward error relative to
the mean absolute window content (elementwise relative error is ill-posed there).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases(npz):
    names = sorted({k.rsplit("__", 1)[0] for k in npz.files})
    for n in names:
        frames, ch, k = (int(v) for v in npz[n + "__meta"])
        yield n, npz[n + "__x"], npz[n + "__y"], frames, ch, k


def _rel(y, e):
    m = e != 0
    return float(np.max(np.abs(y[m].astype(np.float64) - e[m]) / np.abs(e[m]))) if m.any() else 0.0


def _fwd_err(y, e, x, k, oracle_mod):
    """max |y - e| / ((1/k) * sum |x_j| over the window)"""
    scale = oracle_mod.mavg_f64(np.abs(x), k)
    m = scale > 0
    return float(np.max(np.abs(y[m].astype(np.float64) - e[m]) / scale[m]))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


# ------------------------------------------------------------------ golden fixtures
def test_golden_i16_bit_exact(mavg):
    g = np.load(os.path.join(GOLD, "golden_i16.npz"))
    for name, x, y, frames, ch, k in _cases(g):
        got = mavg.moving_average(x, k, channels=ch)
        assert got.dtype == np.int16
        assert np.array_equal(got, y), name


def test_golden_f32(mavg, oracle_mod):
    g = np.load(os.path.join(GOLD, "golden_f32.npz"))
    for name, x, y, frames, ch, k in _cases(g):
        got = mavg.moving_average(x, k, channels=ch)
        if "usym" in name:
            assert _fwd_err(got, y, x, k, oracle_mod) < TOL, name
        else:
            assert _rel(got, y) < TOL, name


# ------------------------------------------------------------------ streaming kernel, all modes
K_SWEEP = [1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 17, 31, 33, 64, 100, 255, 256, 257, 511, 1000, 1024, 2047, 4095, 4096]


@pytest.mark.parametrize("k", K_SWEEP)
def test_stream_path_vs_oracle(mavg, oracle_mod, k):
    n = 5 * 4096 + 32 * 7                      # whole rows, several tiles, ragged last tile
    x = oracle_mod.fill_f32(n, 1000 + k)
    with mavg.Plan(n, k, path="stream") as plan:
        assert plan.info.path == 1
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mavg_f64(x, k)) < TOL


@pytest.mark.parametrize("k", [3, 16, 64, 256, 1024, 4096])
def test_stream_path_matches_arithmetic_model_bitwise(mavg, oracle_mod, k):
    """The CUDA kernel performs exactly the fp32 operations of tests/algo_model.py."""
    from algo_model import stream_model
    n = 3 * 8192
    x = oracle_mod.fill_f32(n, 2000 + k)
    for threads, run in ((256, 16), (512, 16), (256, 32)):
        with mavg.Plan(n, k, path="stream", threads=threads, run=run) as plan:
            y = plan.run_host(x)
        assert np.array_equal(y, stream_model(x, k, threads, run)), (threads, run)


@pytest.mark.parametrize("shape", [(256, 16), (256, 32), (512, 16)])
@pytest.mark.parametrize("k", [5, 64, 300, 4096, 9000])
def test_stream_shapes(mavg, oracle_mod, shape, k):
    threads, run = shape
    n = 4 * threads * run + 64
    x = oracle_mod.fill_f32(n, 3000 + k)
    with mavg.Plan(n, k, path="stream", threads=threads, run=run) as plan:
        y = plan.run_host(x)
        assert plan.info.tile_samples == threads * run
    assert _rel(y, oracle_mod.mavg_f64(x, k)) < TOL


@pytest.mark.parametrize("k", [3, 100, 2000])
def test_chunking_and_prefetch_do_not_change_bits(mavg, oracle_mod, k):
    n = 300 * 4096
    x = oracle_mod.fill_f32(n, 4000 + k)
    with mavg.Plan(n, k, path="stream") as plan:
        base = plan.run_host(x)
    for tune in (dict(chunks_per_cta=3), dict(prefetch=1), dict(prefetch=3, ctas_per_sm=1), dict(ctas_per_sm=1, chunks_per_cta=7)):
        with mavg.Plan(n, k, path="stream", **tune) as plan:
            assert np.array_equal(plan.run_host(x), base), tune
    assert _rel(base, oracle_mod.mavg_f64(x, k)) < TOL


@pytest.mark.parametrize("dist", ["USYM", "I16", "DC1E4"])
@pytest.mark.parametrize("k", [7, 64, 1024])
def test_other_distributions(mavg, oracle_mod, dist, k):
    d = getattr(oracle_mod, "DIST_" + dist)
    n = 6 * 4096
    x = oracle_mod.fill_f32(n, 5000 + k, d)
    y = mavg.moving_average(x, k)
    e = oracle_mod.mavg_f64(x, k)
    assert _fwd_err(y, e, x, k, oracle_mod) < TOL
    if dist == "DC1E4":
        assert _rel(y, e) < TOL   # tile-local rebasing keeps the DC offset from eating the mantissa


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 4095, 4096, 4097, 8191, 12345])
@pytest.mark.parametrize("k", [1, 3, 50, 5000])
def test_ragged_lengths_f32(mavg, oracle_mod, n, k):
    x = oracle_mod.fill_f32(n, 6000 + n + k)
    y = mavg.moving_average(x, k)
    assert y.shape == x.shape
    assert _rel(y, oracle_mod.mavg_f64(x, k)) < TOL


def test_empty_input(mavg):
    assert mavg.moving_average(np.zeros(0, dtype=np.float32), 5).size == 0
    assert mavg.moving_average(np.zeros(0, dtype=np.int16), 5).size == 0


@pytest.mark.parametrize("ch", [1, 2, 3, 8])
@pytest.mark.parametrize("k", [1, 2, 5, 41, 64, 850, 1000, 70000])
def test_i16_bit_exact_vs_oracle(mavg, oracle_mod, ch, k):
    frames = 3000 if k < 70000 else 800
    x = oracle_mod.fill_i16(frames * ch, 7000 + k + ch)
    y = mavg.moving_average(x, k, channels=ch)
    assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))


def test_i16_extremes_bit_exact(mavg, oracle_mod):
    for val in (-32768, 32767):
        x = np.full(5000, val, dtype=np.int16)
        for k in (1, 3, 4096):
            assert np.array_equal(mavg.moving_average(x, k), oracle_mod.mavg_i16(x, k))


@pytest.mark.parametrize("ch", [2, 3, 5])
@pytest.mark.parametrize("k", [4, 64, 700])
def test_interleaved_f32(mavg, oracle_mod, ch, k):
    frames = 5000
    x = oracle_mod.fill_f32(frames * ch, 8000 + k + ch)
    y = mavg.moving_average(x, k, channels=ch)
    assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL


@pytest.mark.parametrize("frames", [4096 * 2, 5000, 4099])
@pytest.mark.parametrize("k", [3, 64, 1500])
def test_planar_batch(mavg, oracle_mod, frames, k):
    ch = 5
    x = oracle_mod.fill_f32(frames * ch, 9000 + k)
    y = mavg.moving_average(x, k, channels=ch, layout="planar")
    for c in range(ch):
        seg = slice(c * frames, (c + 1) * frames)
        assert _rel(y[seg], oracle_mod.mavg_f64(x[seg], k)) < TOL, c


def test_unaligned_device_pointer_takes_generic_path(mavg, oracle_mod, torch_cuda):
    torch = torch_cuda
    n, k = 20000, 37
    x = oracle_mod.fill_f32(n + 1, 11)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(n + 1, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()   # the plan runs on its own non-blocking stream
    with mavg.Plan(n, k) as plan:
        plan.run_device([dx.data_ptr() + 4], [dy.data_ptr() + 4])
        plan.synchronize()
        assert plan.info.launches_per_run >= 1
    y = dy.cpu().numpy()[1:]
    assert _rel(y, oracle_mod.mavg_f64(x[1:], k)) < TOL


def test_block_size_rule_matches_reference(mavg):
    with pytest.raises(mavg.MavgError) as e:
        mavg.Plan(1000, 5, block_size=100)
    assert e.value.status == -6
    mavg.Plan(1000, 5, block_size=256).close()


# ------------------------------------------------------------------ device-resident runs, generator, halo
def test_device_generator_matches_oracle(mavg, oracle_mod, torch_cuda):
    torch = torch_cuda
    n = 100_000
    for dist in (0, 1, 2, 3):
        d = torch.empty(n, dtype=torch.float32, device="cuda")
        mavg.fill_synthetic_device(d.data_ptr(), "f32", n, 12345, 99, dist)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), oracle_mod.fill_f32(n, 99, dist, first_index=12345)), dist
    d = torch.empty(n, dtype=torch.int16, device="cuda")
    mavg.fill_synthetic_device(d.data_ptr(), "i16", n, 777, 5)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle_mod.fill_i16(n, 5, first_index=777))


@pytest.mark.parametrize("k", [3, 64, 1024, 4096])
def test_shard_with_halo_is_bit_identical(mavg, oracle_mod, torch_cuda, k):
    """Second half of a signal filtered as its own shard, left context read in place from the
    first half's tail: the same bits as the unsharded run."""
    torch = torch_cuda
    T = 8192
    n = 64 * T
    cut = 24 * T
    x = oracle_mod.fill_f32(n, 13)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(n, dtype=torch.float32, device="cuda")
    dz = torch.zeros(n - cut, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()   # the plan runs on its own non-blocking stream
    with mavg.Plan(n, k, path="stream") as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    whole = dy.cpu().numpy()
    with mavg.Plan(n - cut, k, path="stream", first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        assert halo >= k and halo % T == 0
        plan.run_device_halo(dx.data_ptr() + 4 * cut, dz.data_ptr(), dx.data_ptr() + 4 * (cut - halo))
        plan.synchronize()
    assert np.array_equal(dz.cpu().numpy(), whole[cut:])
    assert _rel(whole, oracle_mod.mavg_f64(x, k)) < TOL


def test_generic_path_with_halo(mavg, oracle_mod, torch_cuda):
    torch = torch_cuda
    ch, k, frames, cut = 2, 300, 9000, 4000
    x = oracle_mod.fill_i16(frames * ch, 17)
    dx = torch.from_numpy(x).cuda()
    dz = torch.zeros((frames - cut) * ch, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut, path="generic") as plan:
        halo = int(plan.info.halo_frames)
        assert halo == k and plan.info.path == 2      # the generic kernel asks for exactly k frames of context
        plan.run_device_halo(dx.data_ptr() + 2 * cut * ch, dz.data_ptr(), dx.data_ptr() + 2 * (cut - halo) * ch)
        plan.synchronize()
    assert np.array_equal(dz.cpu().numpy(), oracle_mod.mavg_i16(x, k, ch)[cut * ch:])
    # 40-channel int16 takes the generic kernel by itself (no streaming kernel for that shape); float32 forced
    with mavg.Plan(frames, k, channels=40, dtype="i16") as plan:
        assert plan.info.path == 2
    xf = oracle_mod.fill_f32(frames * 3, 18)
    dxf = torch.from_numpy(xf).cuda()
    dzf = torch.zeros((frames - cut) * 3, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=3, first_frame=cut, path="generic") as plan:
        halo = int(plan.info.halo_frames)
        assert halo == k and plan.info.path == 2
        plan.run_device_halo(dxf.data_ptr() + 4 * cut * 3, dzf.data_ptr(), dxf.data_ptr() + 4 * (cut - halo) * 3)
        plan.synchronize()
    assert _rel(dzf.cpu().numpy(), oracle_mod.mavg_f64(xf, k, 3)[cut * 3:]) < TOL


def test_timing_and_info(mavg, oracle_mod):
    n = 1 << 22
    x = oracle_mod.fill_f32(n, 21)
    with mavg.Plan(n, 64) as plan:
        plan.run_host(x)
        t = plan.timing()
        assert t.compute_ms > 0 and t.h2d_ms > 0 and t.d2h_ms > 0
        assert abs(t.total_ms - (t.h2d_ms + t.compute_ms + t.d2h_ms)) < 1e-3
        i = plan.info
        assert i.path == 1 and i.launches_per_run == 4 and i.grid > 0     # 16 MiB of samples = four 4 MiB slices
    with mavg.Plan(n, 64, slice_bytes=64 << 20) as plan:
        assert np.array_equal(plan.run_host(x), mavg.moving_average(x, 64))   # slicing never changes a bit
        assert plan.info.launches_per_run == 1


def test_owned_buffers_synthetic_run(mavg, oracle_mod, torch_cuda):
    n, k = 1 << 20, 5          # BASELINE.json configs[0]: 2^20 mono, k = 5
    with mavg.Plan(n, k) as plan:
        plan.fill_synthetic(0x5EED0001, 0)
        plan.run_owned()
        plan.synchronize()
        # read the owned output back through a host run on the same input
        x = oracle_mod.fill_f32(n, 0x5EED0001)
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mavg_f64(x, k)) < TOL


# ------------------------------------------------------------------ BASELINE.json full size
@pytest.mark.parametrize("k", [3, 9, 16, 64, 256, 257, 1024, 4095, 4096])
def test_full_size_2p28_vs_oracle(mavg, oracle_mod, k):
    """Every window of the headline sweep (BASELINE.json configs[1], [2]) at the full 2^28 samples, plus the first
    window of each arithmetic mode past a seam (9, 257) and an odd window with a misaligned lag run (4095)."""
    n = 1 << 28
    x = oracle_mod.fill_f32(n, 0x5EED0000 + k)
    with mavg.Plan(n, k) as plan:
        y = plan.run_host(x)
    e = oracle_mod.mavg_f64(x, k)
    assert _rel(y, e) < TOL
    del e
    # size-independent property: constant after warm-up for a constant signal, exactly c for c = 1
    x.fill(1.0)
    with mavg.Plan(n, k) as plan:
        y = plan.run_host(x)
    assert np.all(np.abs(y[k:] - 1.0) < 1e-6)


# ------------------------------------------------------------------ int16 streaming kernel (the reference's format)
@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 16, 31, 41, 64, 255, 256, 272, 273, 544, 545, 1000, 1024, 4095, 4096, 9000,
                               20000, 32768, 32769])
def test_i16_stream_kernel_bit_exact(mavg, oracle_mod, ch, k):
    """Multi-tile int16 signals through the TMA streaming kernel: bit-identical to the reference CPU path
    (int64-exact sums, truncating division) for every window, both arithmetic modes, mono and stereo."""
    frames = (5 * 16384 + 64 * 3) // ch + 5          # several tiles, ragged rows, ragged tail
    x = oracle_mod.fill_i16(frames * ch, 12000 + k + ch)
    with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
        if 2 <= k <= 4096:      # k == 1 is the identity and is left to the generic kernel
            assert plan.info.path == 1, "headline windows must take the streaming kernel"
        y = plan.run_host(x)
    assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))


def test_i16_stream_extremes_and_negative_truncation(mavg, oracle_mod):
    n = 3 * 16384
    for val in (-32768, 32767, -1, 1):
        x = np.full(n, val, dtype=np.int16)
        for k in (1, 3, 7, 4096, 30000):
            assert np.array_equal(mavg.moving_average(x, k), oracle_mod.mavg_i16(x, k)), (val, k)
    # alternating signs exercise truncation toward zero on both sides
    x = (np.arange(n) % 7 - 3).astype(np.int16) * 1111
    for k in (2, 5, 100):
        assert np.array_equal(mavg.moving_average(x, k, channels=2), oracle_mod.mavg_i16(x, k, 2))


@pytest.mark.parametrize("ch", [1, 2])
@pytest.mark.parametrize("k", [5, 700, 4096])
def test_i16_shard_with_halo_bit_exact(mavg, oracle_mod, torch_cuda, ch, k):
    torch = torch_cuda
    tile_frames = 16384 // ch
    frames, cut = 20 * tile_frames + 333, 7 * tile_frames
    x = oracle_mod.fill_i16(frames * ch, 13000 + k)
    dx = torch.from_numpy(x).cuda()
    dz = torch.zeros((frames - cut) * ch, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        assert plan.info.path == 1 and halo % tile_frames == 0 and halo >= k
        plan.run_device_halo(dx.data_ptr() + 2 * cut * ch, dz.data_ptr(), dx.data_ptr() + 2 * (cut - halo) * ch)
        plan.synchronize()
    assert np.array_equal(dz.cpu().numpy(), oracle_mod.mavg_i16(x, k, ch)[cut * ch:])


def test_i16_planar_batch(mavg, oracle_mod):
    ch, frames, k = 3, 2 * 16384 + 8 * 5, 100
    x = oracle_mod.fill_i16(ch * frames, 14000)
    y = mavg.moving_average(x, k, channels=ch, layout="planar")
    for c in range(ch):
        seg = slice(c * frames, (c + 1) * frames)
        assert np.array_equal(y[seg], oracle_mod.mavg_i16(x[seg], k))


# ------------------------------------------------------------------ stereo float32 through the streaming kernel
@pytest.mark.parametrize("k", [1, 2, 3, 5, 7, 8, 9, 16, 17, 64, 128, 129, 255, 256, 1000, 1024, 2048, 4096, 5000])
def test_stereo_f32_stream_kernel(mavg, oracle_mod, k):
    frames = 3 * 4096 + 37 * 16 + 3            # several stereo tiles (4096 frames each), ragged rows and tail
    x = oracle_mod.fill_f32(2 * frames, 15000 + k)
    with mavg.Plan(frames, k, channels=2) as plan:
        assert plan.info.path == 1
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mavg_f64(x, k, 2)) < TOL
    # channel independence: the left channel equals the mono filter of the left samples
    if k in (3, 64, 1024):
        left = np.ascontiguousarray(x[0::2])
        np.testing.assert_allclose(y[0::2], mavg.moving_average(left, k), rtol=2e-6, atol=0)


@pytest.mark.parametrize("k", [3, 300, 4096])
def test_stereo_f32_shard_with_halo_bit_identical(mavg, oracle_mod, torch_cuda, k):
    torch = torch_cuda
    tf = 4096                                   # frames per stereo tile
    frames, cut = 30 * tf + 100, 11 * tf
    x = oracle_mod.fill_f32(2 * frames, 16000 + k)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(2 * frames, dtype=torch.float32, device="cuda")
    dz = torch.zeros(2 * (frames - cut), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=2) as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    with mavg.Plan(frames - cut, k, channels=2, first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        assert halo % tf == 0 and halo >= k
        plan.run_device_halo(dx.data_ptr() + 8 * cut, dz.data_ptr(), dx.data_ptr() + 8 * (cut - halo))
        plan.synchronize()
    whole = dy.cpu().numpy()
    assert np.array_equal(dz.cpu().numpy(), whole[2 * cut:])
    assert _rel(whole, oracle_mod.mavg_f64(x, k, 2)) < TOL


# ------------------------------------------------------------------ many-channel interleaved float32 (column kernel)
@pytest.mark.parametrize("ch", [32, 36, 64, 100, 256])
@pytest.mark.parametrize("k", [9, 16, 17, 64, 255, 256, 700, 1024, 1500, 5, 1, 2, 3, 8])
def test_many_channel_interleaved_f32(mavg, oracle_mod, ch, k):
    frames = 3 * 256 + 77 if ch >= 100 else 9 * 256 + 13
    x = oracle_mod.fill_f32(frames * ch, 17000 + k + ch)
    with mavg.Plan(frames, k, channels=ch) as plan:
        y = plan.run_host(x)
        i = plan.info
        if k <= 1024:
            assert i.path == 1 and i.mode == 3, "expected the column kernel"
        else:
            assert i.path == 2
    assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL


@pytest.mark.parametrize("k", [64, 1000])
def test_many_channel_shard_with_halo_bit_identical(mavg, oracle_mod, torch_cuda, k):
    torch = torch_cuda
    ch, tf = 64, 256
    frames, cut = 40 * tf + 50, 13 * tf
    x = oracle_mod.fill_f32(frames * ch, 18000 + k)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(frames * ch, dtype=torch.float32, device="cuda")
    dz = torch.zeros((frames - cut) * ch, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch) as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        assert halo % 128 == 0 and halo >= k and cut % halo == 0 or halo <= cut
        plan.run_device_halo(dx.data_ptr() + 4 * cut * ch, dz.data_ptr(), dx.data_ptr() + 4 * (cut - halo) * ch)
        plan.synchronize()
    whole = dy.cpu().numpy()
    assert np.array_equal(dz.cpu().numpy(), whole[cut * ch:])
    assert _rel(whole, oracle_mod.mavg_f64(x, k, ch)) < TOL


# ------------------------------------------------------------------ generic kernel, long windows (block-sum start)
@pytest.mark.parametrize("case", [("f32", 3, 4096), ("f32", 6, 1000), ("i16", 5, 3000), ("i16", 3, 257), ("f32", 1, 60000)])
def test_generic_long_windows(mavg, oracle_mod, case):
    dtype, ch, k = case
    frames = 150_001
    if dtype == "f32":
        x = oracle_mod.fill_f32(frames * ch, 19000 + k)
        y = mavg.moving_average(x, k, channels=ch, path="generic")
        assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL
    else:
        x = oracle_mod.fill_i16(frames * ch, 19000 + k)
        y = mavg.moving_average(x, k, channels=ch, path="generic")
        assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))


# ------------------------------------------------------------------ full size, the other fast paths
def test_full_size_i16_stereo_2p28_bit_exact(mavg, oracle_mod):
    """2^28 stereo int16 samples (the reference's format at the benchmark size): every output bit-identical
    to the CPU oracle, for a direct-mode and a scan-mode window."""
    n = 1 << 28
    x = oracle_mod.fill_i16(n, 0x5EED0100)
    for k in (5, 1000):
        with mavg.Plan(n // 2, k, channels=2, dtype="i16") as plan:
            assert plan.info.path == 1
            y = plan.run_host(x)
        e = oracle_mod.mavg_i16_mt(x, k, 2, 8)
        assert np.array_equal(y, e), k
        del y, e


def test_full_size_f32_stereo_2p27_frames(mavg, oracle_mod):
    n = 1 << 28                                   # 2^27 stereo frames
    x = oracle_mod.fill_f32(n, 0x5EED0200)
    with mavg.Plan(n // 2, 64, channels=2) as plan:
        assert plan.info.path == 1
        y = plan.run_host(x)
    e = oracle_mod.mavg_f64(x, 64, 2)
    assert _rel(y, e) < TOL


def test_pinned_host_buffers(mavg, oracle_mod):
    """run_host on page-locked buffers from the library (the drop-in binaries use the same allocator)."""
    n, k = 3 * (1 << 22) + 7, 100
    with mavg.PinnedArray(n, np.float32) as hin, mavg.PinnedArray(n, np.float32) as hout:
        hin.array[:] = oracle_mod.fill_f32(n, 23000)
        with mavg.Plan(n, k) as plan:
            y = plan.run_host(hin.array, out=hout.array)
            assert y is hout.array
            assert plan.timing().total_ms > 0
        assert _rel(hout.array, oracle_mod.mavg_f64(hin.array, k)) < TOL


def test_caller_owned_buffers_page_locked_in_place(mavg, oracle_mod):
    """mavg_host_register on NumPy-owned memory (what a std::vector is to the reference's GpuLoad functions): same
    result, and the arrays stay ordinary arrays afterwards."""
    n, k = (1 << 22) + 5, 77
    x = oracle_mod.fill_f32(n, 23100)
    y = np.empty_like(x)
    e = oracle_mod.mavg_f64(x, k)
    with mavg.Plan(n, k) as plan:
        with mavg.pinned(x, y):
            with mavg.pinned(x):                   # registering twice is harmless
                assert plan.run_host(x, out=y) is y
            assert _rel(y, e) < TOL
        y[:] = 0
        plan.run_host(x, out=y)                    # pageable again
        assert _rel(y, e) < TOL
    with pytest.raises(ValueError):
        with mavg.pinned(x[::2]):
            pass


# ------------------------------------------------------------------ 3..31 interleaved float32 channels (few-channel kernel)
@pytest.mark.parametrize("ch", [3, 5, 6, 7, 8, 12, 24, 31])
@pytest.mark.parametrize("k", [1, 2, 3, 8, 9, 16, 17, 64, 100, 255, 256, 300])
def test_few_channel_interleaved_f32(mavg, oracle_mod, ch, k):
    frames = 3 * (512 // ch) * 16 + 41           # several tiles, ragged tail (flat length not a multiple of 32)
    x = oracle_mod.fill_f32(frames * ch, 24000 + k + ch)
    runs = 512 // ch
    while runs * ch % 16:
        runs -= 1
    with mavg.Plan(frames, k, channels=ch) as plan:
        y = plan.run_host(x)
        i = plan.info
        if k <= 256 and (k + 15) // 16 <= runs:
            assert i.path == 1 and i.mode == 4, "expected the few-channel kernel"
    assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL


@pytest.mark.parametrize("case", [(6, 64), (3, 5), (7, 200)])
def test_few_channel_shard_with_halo_bit_identical(mavg, oracle_mod, torch_cuda, case):
    torch = torch_cuda
    ch, k = case
    with mavg.Plan(100_000, k, channels=ch) as probe:
        tf = int(probe.info.halo_frames)            # one history tile
        assert probe.info.mode == 4 and tf >= k
    frames, cut = 37 * tf + 123, 9 * tf
    x = oracle_mod.fill_f32(frames * ch, 25000 + k)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(frames * ch, dtype=torch.float32, device="cuda")
    dz = torch.zeros((frames - cut) * ch, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch) as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        plan.run_device_halo(dx.data_ptr() + 4 * cut * ch, dz.data_ptr(), dx.data_ptr() + 4 * (cut - halo) * ch)
        plan.synchronize()
    whole = dy.cpu().numpy()
    tail_frames = ((frames - cut) * ch // 32) * 32 // ch      # frames produced by the streaming kernel in the shard run
    assert np.array_equal(dz.cpu().numpy()[:tail_frames * ch], whole[cut * ch:(cut + tail_frames) * ch])
    assert _rel(dz.cpu().numpy(), oracle_mod.mavg_f64(x, k, ch)[cut * ch:]) < TOL
    assert _rel(whole, oracle_mod.mavg_f64(x, k, ch)) < TOL


# ------------------------------------------------------------------ 3..31 interleaved int16 channels (multichannel PCM)
def _fewc_runs(ch):
    runs = 512 // ch
    while runs * ch % 16:
        runs -= 1
    return runs


# channel counts the flat-stream int16 kernel takes (info.mode 6) with runs of whole frames: (threads, run) of the shape
# with one CTA per SM; 3 / 4 / 6 / 8 channels also have a 512-thread shape (tuning.threads = 512)
FLAT_I16_SHAPE = {3: (224, 72), 6: (224, 72), 9: (224, 72), 12: (224, 72), 4: (256, 64), 8: (256, 64), 16: (256, 64),
                  5: (384, 40), 10: (384, 40), 7: (256, 56)}
FLAT_I16_CH = tuple(sorted(FLAT_I16_SHAPE))


@pytest.mark.parametrize("ch", [3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 24, 31])
@pytest.mark.parametrize("k", [1, 2, 3, 7, 8, 31, 32, 33, 64, 100, 255, 256, 300])
def test_few_channel_interleaved_i16_bit_exact(mavg, oracle_mod, ch, k):
    """3 ... 8, 12, 16 channels: the stereo kernel's delta scan over the flat stream (runs of whole frames).  Other odd
    channel counts: 2-byte accesses, 32-frame runs.  Other even counts: channel pairs as 32-bit words, 16-frame runs."""
    pair = ch % 2 == 0
    runs = _fewc_runs(ch // 2 if pair else ch)
    rf = 16 if pair else 32
    frames = 3 * runs * rf + 41                    # several tiles, ragged tail (flat length not a multiple of 64)
    if ch in FLAT_I16_CH:
        frames = 3 * FLAT_I16_SHAPE[ch][0] * FLAT_I16_SHAPE[ch][1] // ch + 41
    x = oracle_mod.fill_i16(frames * ch, 26000 + k + ch)
    with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
        y = plan.run_host(x)
        i = plan.info
        if ch in FLAT_I16_CH:
            if k >= 2:
                assert i.path == 1 and i.mode == 6 and i.run == FLAT_I16_SHAPE[ch][1], "expected the flat-stream int16 kernel"
        elif 2 <= k <= 256 and (k + rf - 1) // rf <= runs:
            assert i.path == 1 and i.mode == 4 and i.run == rf, "expected the few-channel int16 kernel"
    assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))
    if ch in (3, 4, 6, 8) and k >= 2:                 # the 512-thread shape (runs of 24 / 32 samples)
        with mavg.Plan(frames, k, channels=ch, dtype="i16", threads=512) as plan:
            assert plan.info.mode == 6 and plan.info.run == (24 if ch % 3 == 0 else 32)
            assert np.array_equal(plan.run_host(x), oracle_mod.mavg_i16(x, k, ch))


@pytest.mark.parametrize("ch", FLAT_I16_CH)
def test_flat_multichannel_i16_every_lag_alignment_and_ring_depth(mavg, oracle_mod, ch):
    """Every lag misalignment (k C mod 8 samples), windows of one to several tiles of history, the longest window of the
    ring and the first one beyond it (which has to leave the kernel and still be exact)."""
    tile_frames = FLAT_I16_SHAPE[ch][0] * FLAT_I16_SHAPE[ch][1] // ch   # tile of the one-CTA shape (the 512-thread one: 512 x 24 / 32)
    frames = 9 * tile_frames + 77
    x = oracle_mod.fill_i16(frames * ch, 26500 + ch)
    ks = list(range(2, 19)) + [tile_frames - 1, tile_frames, tile_frames + 1, 2 * tile_frames + 3, 3 * tile_frames - 5]
    for k in ks:
        e = oracle_mod.mavg_i16(x, k, ch)
        for threads in (0, 256, 512):      # auto (two 128-thread CTAs per SM while the window allows), one CTA of 256 / 224, of 512
            with mavg.Plan(frames, k, channels=ch, dtype="i16", threads=threads) as plan:
                assert plan.info.path == 1 and plan.info.mode == 6, (ch, k)
                assert np.array_equal(plan.run_host(x), e), (ch, k, threads)
    # beyond the ring: any other path, same result
    with mavg.Plan(1 << 20, 32768, channels=ch, dtype="i16") as plan:
        assert plan.info.mode != 6
    k = 7 * tile_frames + 1
    if k <= 32768:
        with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
            assert np.array_equal(plan.run_host(x), oracle_mod.mavg_i16(x, k, ch)), (ch, k)


def test_few_channel_i16_extremes_and_negative_truncation(mavg, oracle_mod):
    """Saturated inputs (window sums up to 256 * 32768 in magnitude) and sign-alternating ramps whose sums straddle
    zero: the multiply-high division has to truncate toward zero exactly as C's `/` does."""
    for ch in (6, 3, 8, 5, 9, 14, 11):             # flat-stream kernel (6, 3, 8, 5, 9), pair kernel (14), scalar kernel (11)
        frames = 4 * 80 * 32 + 5
        for val in (-32768, 32767, -1, 1):
            x = np.full(frames * ch, val, dtype=np.int16)
            for k in (2, 3, 7, 100, 255, 256):
                with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
                    assert plan.info.mode == (6 if ch in FLAT_I16_CH else 4)
                    assert np.array_equal(plan.run_host(x), oracle_mod.mavg_i16(x, k, ch)), (ch, val, k)
        x = ((np.arange(frames * ch) % 11 - 5) * 997).astype(np.int16)
        for k in (2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 100):
            assert np.array_equal(mavg.moving_average(x, k, channels=ch), oracle_mod.mavg_i16(x, k, ch)), (ch, k)


@pytest.mark.parametrize("case", [(6, 64), (3, 5), (7, 200), (8, 256), (4, 2), (12, 64), (6, 3000), (4, 9000), (3, 4097),
                                  (10, 64), (9, 100), (14, 64), (11, 100), (5, 3001), (16, 1500)])
def test_few_channel_i16_shard_with_halo_bit_exact(mavg, oracle_mod, torch_cuda, case):
    torch = torch_cuda
    ch, k = case
    with mavg.Plan(100_000, k, channels=ch, dtype="i16") as probe:
        tf = int(probe.info.halo_frames)
        assert probe.info.mode == (6 if ch in FLAT_I16_CH else 4) and tf >= k
    frames, cut = 37 * tf + 123, 9 * tf
    x = oracle_mod.fill_i16(frames * ch, 27000 + k)
    dx = torch.from_numpy(x).cuda()
    dz = torch.zeros((frames - cut) * ch, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut) as plan:
        halo = int(plan.info.halo_frames)
        plan.run_device_halo(dx.data_ptr() + 2 * cut * ch, dz.data_ptr(), dx.data_ptr() + 2 * (cut - halo) * ch)
        plan.synchronize()
    assert np.array_equal(dz.cpu().numpy(), oracle_mod.mavg_i16(x, k, ch)[cut * ch:])


# ------------------------------------------------------------------ few-channel kernels, long windows (prefix mode)
@pytest.mark.parametrize("ch", [3, 4, 5, 6, 8, 12, 24, 30])
@pytest.mark.parametrize("k", [257, 272, 273, 512, 545, 1000, 1024, 2048, 4096])
@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_few_channel_long_windows(mavg, oracle_mod, ch, k, dtype):
    """Windows longer than 16 runs: run totals become per-tile prefixes, the window start is a prefix difference
    reaching up to H tiles back.  Whatever does not fit shared memory falls back to the generic kernel; either
    way the result has to match the oracle (bit-exact for int16)."""
    frames = 26000 + 41
    if dtype == "f32":
        x = oracle_mod.fill_f32(frames * ch, 28000 + k + ch)
    else:
        x = oracle_mod.fill_i16(frames * ch, 28000 + k + ch)
    with mavg.Plan(frames, k, channels=ch, dtype=dtype) as plan:
        y = plan.run_host(x)
        i = plan.info
        if dtype == "i16" and ch in FLAT_I16_CH and ch <= 8:
            assert i.path == 1 and i.mode == 6, "expected the flat-stream int16 kernel"
        elif ch <= 8 and k <= 1024:
            assert i.path == 1 and i.mode == 4, "expected a few-channel streaming kernel"
    if dtype == "f32":
        assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL
    else:
        assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))


@pytest.mark.parametrize("dist", ["USYM", "DC1E4"])
def test_few_channel_long_window_conditioning(mavg, oracle_mod, dist):
    """Zero-mean and large-offset inputs through the prefix mode: forward error against sum |x| / k stays tiny
    (prefix differences never leave one tile)."""
    ch, k, frames = 6, 1000, 60000
    x = oracle_mod.fill_f32(frames * ch, 29000, dist=getattr(oracle_mod, "DIST_" + dist))
    with mavg.Plan(frames, k, channels=ch) as plan:
        assert plan.info.mode == 4
        y = plan.run_host(x)
    e = oracle_mod.mavg_f64(x, k, ch)
    scale = oracle_mod.mavg_f64(np.abs(x), k, ch)
    assert np.max(np.abs(y - e) / np.maximum(scale, 1e-30)) < TOL


@pytest.mark.parametrize("case", [("f32", 6, 1000), ("f32", 3, 4096), ("i16", 6, 1000), ("i16", 5, 2048), ("i16", 14, 1000)])
def test_few_channel_long_window_shard_with_halo(mavg, oracle_mod, torch_cuda, case):
    torch = torch_cuda
    dtype, ch, k = case
    es = 4 if dtype == "f32" else 2
    tdt = torch.float32 if dtype == "f32" else torch.int16
    with mavg.Plan(1_000_000, k, channels=ch, dtype=dtype) as probe:
        halo = int(probe.info.halo_frames)
        tiles_back = int(probe.info.history_tiles)
        assert probe.info.mode == (6 if dtype == "i16" and ch in FLAT_I16_CH else 4) and halo >= k and tiles_back >= 1
    tf = halo // tiles_back
    frames, cut = 23 * tf + 123, 7 * tf
    x = (oracle_mod.fill_f32 if dtype == "f32" else oracle_mod.fill_i16)(frames * ch, 30000 + k)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(frames * ch, dtype=tdt, device="cuda")
    dz = torch.zeros((frames - cut) * ch, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch, dtype=dtype) as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype=dtype, first_frame=cut) as plan:
        assert int(plan.info.halo_frames) == halo
        plan.run_device_halo(dx.data_ptr() + es * cut * ch, dz.data_ptr(), dx.data_ptr() + es * (cut - halo) * ch)
        plan.synchronize()
    whole, part = dy.cpu().numpy(), dz.cpu().numpy()
    if dtype == "f32":
        stream_frames = ((frames - cut) * ch // 32) * 32 // ch     # the rest is the generic tail
        assert np.array_equal(part[:stream_frames * ch], whole[cut * ch:(cut + stream_frames) * ch])
        assert _rel(whole, oracle_mod.mavg_f64(x, k, ch)) < TOL
        assert _rel(part, oracle_mod.mavg_f64(x, k, ch)[cut * ch:]) < TOL
    else:
        e = oracle_mod.mavg_i16(x, k, ch)
        assert np.array_equal(whole, e) and np.array_equal(part, e[cut * ch:])


@pytest.mark.parametrize("case", [("f32", 3, 1 << 24, 1000), ("f32", 6, 1 << 23, 64), ("i16", 6, 1 << 23, 700),
                                  ("i16", 5, 1 << 23, 48), ("i16", 14, 1 << 22, 200), ("i16", 10, 1 << 22, 200), ("i16", 9, 1 << 22, 77), ("i16", 8, 1 << 23, 4000),
                                  ("i16", 3, 1 << 24, 77), ("i16", 4, 1 << 23, 1)])
def test_few_channel_many_tiles_per_cta(mavg, oracle_mod, case):
    """Tens of tiles per persistent CTA (ring wrap-around, staging double buffer, chunk boundaries with history
    replay) on 50 M samples; every output checked."""
    dtype, ch, frames, k = case
    if dtype == "f32":
        x = oracle_mod.fill_f32(frames * ch, 31000 + k)
    else:
        x = oracle_mod.fill_i16(frames * ch, 31000 + k)
    with mavg.Plan(frames, k, channels=ch, dtype=dtype) as plan:
        if k >= 2:
            assert plan.info.mode == (6 if dtype == "i16" and ch in FLAT_I16_CH else 4)
        y = plan.run_host(x)
    if dtype == "f32":
        assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL
    else:
        assert np.array_equal(y, oracle_mod.mavg_i16_mt(x, k, ch, 8))


def test_distinct_plans_on_concurrent_threads(mavg, oracle_mod):
    """a plan is not thread-safe, distinct plans are.  Four host threads, each with its own plan
    (different dtype / channels / window, i.e. different kernels), run concurrently on one GPU."""
    import threading
    cases = [("f32", 1, 64, 1 << 21), ("i16", 2, 300, 1 << 20), ("f32", 6, 100, 1 << 18), ("i16", 5, 1000, 1 << 18)]
    inputs = [oracle_mod.fill_f32(f * c, 90 + i) if d == "f32" else oracle_mod.fill_i16(f * c, 90 + i)
              for i, (d, c, k, f) in enumerate(cases)]
    want = [oracle_mod.mavg_f64(x, k, c) if d == "f32" else oracle_mod.mavg_i16(x, k, c)
            for x, (d, c, k, f) in zip(inputs, cases)]
    errors = []

    def work(i):
        try:
            d, c, k, f = cases[i]
            with mavg.Plan(f, k, channels=c, dtype=d) as plan:
                for _ in range(6):
                    y = plan.run_host(inputs[i])
                    if d == "f32":
                        assert _rel(y, want[i]) < TOL
                    else:
                        assert np.array_equal(y, want[i])
        except Exception as exc:                     # surfaced in the main thread
            errors.append((i, repr(exc)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_errors_are_returned_not_fatal(mavg, oracle_mod):
    """Every entry point returns a status: an impossible device allocation, a null shard pointer
    and aliased buffers are reported through MavgError, and the library keeps working afterwards."""
    from digital_signal_processsing_b200 import _lib
    with mavg.Plan(1 << 38, 64) as plan:                       # 1 TiB of float32 per buffer: cudaMalloc must refuse
        with pytest.raises(_lib.MavgError) as ei:
            plan.run_owned()
        assert ei.value.status == _lib.ERR_ALLOC and "cudaMalloc" in str(ei.value)
    with mavg.Plan(1000, 5) as plan:
        with pytest.raises(_lib.MavgError):
            plan.run_device([0], [0])
        x = oracle_mod.fill_f32(1000, 3)
        with pytest.raises(_lib.MavgError):
            plan.run_host(x, out=x)                              # output must not alias input
        assert _rel(plan.run_host(x), oracle_mod.mavg_f64(x, 5)) < TOL     # still usable


# ------------------------------------------------------------------ many-channel interleaved int16 (column kernel over channel pairs)
@pytest.mark.parametrize("ch", [64, 72, 96, 128, 256, 520])
@pytest.mark.parametrize("k", [1, 2, 3, 8, 9, 64, 100, 300, 1000, 1500])
def test_many_channel_interleaved_i16_bit_exact(mavg, oracle_mod, ch, k):
    frames = 5 * 256 + 37
    x = oracle_mod.fill_i16(frames * ch, 32000 + k + ch)
    with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
        y = plan.run_host(x)
        if 2 <= k <= 96:
            assert plan.info.path == 1 and plan.info.mode == 3, "expected the int16 column kernel"
            assert plan.info.threads == 256 and plan.info.run == 32      # 8 warps x 32 frames
    e = oracle_mod.mavg_i16(x, k, ch)
    assert np.array_equal(y, e)
    for tune in (dict(threads=512), dict(direct_max_k=1), dict(threads=512, direct_max_k=1)):
        # 16 warps x 16 frames; stores from registers instead of staging tiles + TMA stores
        with mavg.Plan(frames, k, channels=ch, dtype="i16", **tune) as plan:
            assert np.array_equal(plan.run_host(x), e), tune


@pytest.mark.parametrize("case", [(64, (1 << 18) + 77, 64, {}), (256, (1 << 16) + 5, 8, {}), (128, (1 << 17) + 33, 200, dict(chunks_per_cta=3)),
                                  (64, (1 << 18) + 77, 700, dict(direct_max_k=1)), (1024, (1 << 14) + 3, 64, dict(threads=512))])
def test_many_channel_i16_many_tiles_per_cta(mavg, oracle_mod, case):
    """Several tiles per persistent CTA (ring wrap-around, staging double buffer, history replay at every range
    boundary, carried window sums) on 16 M samples; every output checked."""
    ch, frames, k, tune = case
    x = oracle_mod.fill_i16(frames * ch, 32500 + k + ch)
    with mavg.Plan(frames, k, channels=ch, dtype="i16", **tune) as plan:
        assert plan.info.path == 1 and plan.info.mode == 3
        y = plan.run_host(x)
    assert np.array_equal(y, oracle_mod.mavg_i16_mt(x, k, ch, 8))


def test_many_channel_i16_extremes_shards_and_odd_counts(mavg, oracle_mod, torch_cuda):
    torch = torch_cuda
    ch, frames = 64, 3000
    for val in (-32768, 32767):
        x = np.full(frames * ch, val, dtype=np.int16)
        for k in (2, 7, 64, 255, 1000):
            assert np.array_equal(mavg.moving_average(x, k, channels=ch), oracle_mod.mavg_i16(x, k, ch)), (val, k)
    # channel counts the pair layout cannot take (C % 8 != 0) stay on the generic kernel and stay exact
    for odd in (66, 100, 33):
        x = oracle_mod.fill_i16(2000 * odd, 33000 + odd)
        with mavg.Plan(2000, 50, channels=odd, dtype="i16") as plan:
            assert plan.info.path == 2
            assert np.array_equal(plan.run_host(x), oracle_mod.mavg_i16(x, 50, odd))
    # shard with halo == whole run
    k = 100
    with mavg.Plan(100_000, k, channels=ch, dtype="i16") as probe:
        assert probe.info.mode == 3
        halo, tiles_back = int(probe.info.halo_frames), int(probe.info.history_tiles)
    tf = halo // tiles_back
    frames, cut = 40 * tf + 77, 11 * tf
    x = oracle_mod.fill_i16(frames * ch, 34000)
    dx = torch.from_numpy(x).cuda()
    dz = torch.zeros((frames - cut) * ch, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut) as plan:
        plan.run_device_halo(dx.data_ptr() + 2 * cut * ch, dz.data_ptr(), dx.data_ptr() + 2 * (cut - halo) * ch)
        plan.synchronize()
    assert np.array_equal(dz.cpu().numpy(), oracle_mod.mavg_i16(x, k, ch)[cut * ch:])


@pytest.mark.parametrize("case", [("f32", 1, 600_000, 60_000, "interleaved"), ("i16", 2, 300_000, 40_000, "interleaved"),
                                  ("f32", 3, 200_000, 9_000, "interleaved"), ("f32", 3, 150_000, 70_000, "planar"),
                                  ("i16", 1, 400_000, 150_000, "interleaved"), ("f32", 40, 30_000, 5_000, "interleaved"),
                                  ("f32", 1, 100_000, 4096 * 20 + 1, "interleaved")])
def test_very_long_windows_on_the_generic_kernel(mavg, oracle_mod, case):
    """Windows whose history exceeds shared memory run on the generic kernel (run starts from 64-frame block sums)."""
    dtype, ch, frames, k, layout = case
    n = frames * ch
    x = oracle_mod.fill_f32(n, 35000 + k) if dtype == "f32" else oracle_mod.fill_i16(n, 35000 + k)
    with mavg.Plan(frames, k, channels=ch, dtype=dtype, layout=layout) as plan:
        # mono / planar float32: far-lag streaming kernel; int16 mono / stereo / 4 / 6 / 8 channels up to k = 46 340: its twin
        far = (dtype == "f32" and (ch == 1 or layout == "planar")) or (dtype == "i16" and k <= 46_340 and ch in (1, 2, 4, 6, 8, 12, 16))
        assert (plan.info.path, plan.info.mode) == ((1, 5) if far else (2, plan.info.mode))
        y = plan.run_host(x)
    if layout == "planar":
        for c in range(ch):
            seg = slice(c * frames, (c + 1) * frames)
            assert _rel(y[seg], oracle_mod.mavg_f64(x[seg], k)) < TOL
    elif dtype == "f32":
        assert _rel(y, oracle_mod.mavg_f64(x, k, ch)) < TOL
    else:
        assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch))


# ------------------------------------------------------------------ far-lag kernel (mono / planar float32, k beyond the ring)
@pytest.mark.parametrize("k", [49_153, 50_000, 60_001, 65_536, 65_538, 99_999, 131_075, 300_000, 1_000_000])
def test_far_lag_kernel_windows(mavg, oracle_mod, k):
    """Every lag misalignment (k mod 4, k mod 32), windows from just past the ring up to a million samples, a signal
    that is not a whole number of tiles or rows, several tiles per CTA."""
    n = 148 * 8192 * 3 + 8192 * 5 + 77
    x = oracle_mod.fill_f32(n, 36000 + k % 1000)
    with mavg.Plan(n, k) as plan:
        assert plan.info.path == 1 and plan.info.mode == 5
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mavg_f64(x, k)) < TOL


@pytest.mark.parametrize("dist", ["USYM", "DC1E4"])
def test_far_lag_kernel_conditioning(mavg, oracle_mod, dist):
    n, k = 6_000_000, 77_777
    x = oracle_mod.fill_f32(n, 37000, dist=getattr(oracle_mod, "DIST_" + dist))
    with mavg.Plan(n, k) as plan:
        assert plan.info.mode == 5
        y = plan.run_host(x)
    e = oracle_mod.mavg_f64(x, k)
    scale = oracle_mod.mavg_f64(np.abs(x), k)
    assert np.max(np.abs(y - e) / np.maximum(scale, 1e-30)) < TOL


def test_far_lag_kernel_slices_shards_and_short_signals(mavg, oracle_mod, torch_cuda):
    torch = torch_cuda
    k = 100_003
    # 2^24 samples through mavg_run_host: 8 MiB slices, each reading its left context in place in front of it
    n = 1 << 24
    x = oracle_mod.fill_f32(n, 38000)
    e = oracle_mod.mavg_f64(x, k)
    with mavg.Plan(n, k) as plan:
        y = plan.run_host(x)
        assert plan.info.launches_per_run >= 4
        halo = int(plan.info.halo_frames)
    assert _rel(y, e) < TOL
    # device run in one piece against the sliced host run: the window sum carried from tile to tile (fp64) starts
    # from a different warm-up in every slice, so the two agree to rounding (~1e-7), not bit for bit
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(n, dtype=torch.float32, device="cuda")
    with mavg.Plan(n, k) as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    whole = dy.cpu().numpy()
    assert _rel(whole, e) < TOL and float(np.max(np.abs(whole - y) / np.abs(e))) < 1e-6
    # shard plan whose context sits right in front of it in device memory
    cut = halo + 5 * 8192
    dz = torch.zeros(n - cut, dtype=torch.float32, device="cuda")
    with mavg.Plan(n - cut, k, first_frame=cut) as plan:
        plan.run_device_halo(dx.data_ptr() + 4 * cut, dz.data_ptr(), dx.data_ptr() + 4 * (cut - halo))
        plan.synchronize()
    assert _rel(dz.cpu().numpy(), e[cut:]) < TOL
    # window longer than the signal (pure warm-up), and a planar batch of three signals
    xs = oracle_mod.fill_f32(70_000, 38001)
    assert _rel(mavg.moving_average(xs, 90_000), oracle_mod.mavg_f64(xs, 90_000)) < TOL
    xp = oracle_mod.fill_f32(3 * 200_000, 38002)
    with mavg.Plan(200_000, 60_000, channels=3, layout="planar") as plan:
        assert plan.info.mode == 5
        yp = plan.run_host(xp)
    for c in range(3):
        seg = slice(c * 200_000, (c + 1) * 200_000)
        assert _rel(yp[seg], oracle_mod.mavg_f64(xp[seg], 60_000)) < TOL


@pytest.mark.parametrize("k", [24_577, 30_000, 50_001, 65_536, 200_000])
def test_far_lag_kernel_stereo(mavg, oracle_mod, k):
    """Interleaved stereo float32 beyond the ring: lag distance 2k flat samples, one carried sum per channel."""
    frames = 148 * 4096 * 2 + 4096 * 3 + 19
    x = oracle_mod.fill_f32(2 * frames, 39000 + k % 1000)
    with mavg.Plan(frames, k, channels=2) as plan:
        assert plan.info.path == 1 and plan.info.mode == 5
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mavg_f64(x, k, 2)) < TOL


@pytest.mark.parametrize("threads", [384, 512])
@pytest.mark.parametrize("L", [60_013, 100_003])
def test_far_lag_kernel_matches_numpy_model_bitwise(mavg, oracle_mod, L, threads):
    """The far-lag kernel performs exactly the operations of tests/algo_model.far_lag_model (one tile per chunk here:
    fewer tiles than SMs), so the two agree bit for bit on every sample the streaming kernel produces.  Tiles of
    384 x 16 samples by default, 512 x 16 with tuning.threads = 512."""
    from algo_model import far_lag_model
    n = 23 * 8192 + 123
    x = oracle_mod.fill_f32(n, 777)
    with mavg.Plan(n, L, path="stream", threads=threads) as plan:
        assert plan.info.mode == 5, "windows of 60 000 samples and more are beyond the ring of the ordinary kernel"
        assert plan.info.threads == threads
        y = plan.run_host(x)
    m = far_lag_model(x, L, chunk_tiles=1, NT=threads)
    whole_rows = n // 32 * 32                  # the last partial row belongs to tail_kernel
    assert np.array_equal(y[:whole_rows], m[:whole_rows])


@pytest.mark.parametrize("ch,k", [(1, 60_001), (2, 30_000), (1, 300_000)])
def test_far_lag_kernel_with_context_in_another_allocation(mavg, oracle_mod, torch_cuda, ch, k):
    """A far-lag shard whose left context is NOT contiguous with it (a peer's tail, a staged halo): the first
    halo_frames frames go through a small plan-owned [context | frames] buffer, the rest finds its context inside the
    shard -- two far-lag launches instead of the generic kernel's quadratic run starts."""
    torch = torch_cuda
    probe = mavg.Plan(1 << 22, k, channels=ch)
    assert probe.info.mode == 5
    halo = int(probe.info.halo_frames)
    probe.close()
    frames = 2 * halo + 8192 * 7 + 33                  # shard longer than the context, ragged end
    cut = halo + 8192 * 3                              # global frame where the shard starts (tile aligned)
    x = oracle_mod.fill_f32((cut + frames) * ch, 39000 + k % 100)
    e = oracle_mod.mavg_f64(x, k, ch)[cut * ch:]
    d_ctx = torch.from_numpy(x[(cut - halo) * ch:cut * ch].copy()).cuda()      # its own allocation
    d_in = torch.from_numpy(x[cut * ch:].copy()).cuda()
    d_out = torch.zeros(frames * ch, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch, first_frame=cut) as plan:
        plan.run_device_halo(d_in.data_ptr(), d_out.data_ptr(), d_ctx.data_ptr())
        plan.synchronize()
        assert plan.info.path == 1 and plan.info.mode == 5 and plan.info.launches_per_run >= 2
    assert _rel(d_out.cpu().numpy(), e) < TOL
    # a shard shorter than the context: everything goes through the staging buffer
    short = halo // 2 // 32 * 32
    d_out2 = torch.zeros(short * ch, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(short, k, channels=ch, first_frame=cut) as plan:
        plan.run_device_halo(d_in.data_ptr(), d_out2.data_ptr(), d_ctx.data_ptr())
        plan.synchronize()
    assert _rel(d_out2.cpu().numpy(), e[:short * ch]) < TOL


# ------------------------------------------------------------------ far-lag int16 kernel (windows beyond the ring, exact)
@pytest.mark.parametrize("ch,k", [(1, 32_769), (1, 32_770), (1, 32_771), (1, 32_772), (1, 32_773), (1, 32_774), (1, 32_775),
                                  (1, 32_776), (1, 46_340), (2, 24_577), (2, 24_578), (2, 24_579), (2, 24_580), (2, 40_000),
                                  (2, 46_340), (4, 12_289), (4, 12_290), (4, 20_001), (4, 46_340), (6, 8_065), (6, 8_066),
                                  (6, 8_067), (6, 8_068), (6, 19_200), (6, 30_001), (8, 6_145), (8, 19_200), (8, 40_000),
                                  (12, 4_033), (12, 4_034), (12, 20_000), (16, 3_073), (16, 12_000)])
def test_far_lag_i16_kernel_bit_exact(mavg, oracle_mod, ch, k):
    """stream_far_i16_kernel: every lag misalignment a channel count can have, the first windows beyond the ring of the
    flat-stream kernel, the longest exact window (k = 46 340), several tiles per CTA, a ragged end -- bit-identical to
    the reference CPU function."""
    frames = ((148 * 3 + 5) * 16384 + 77 * ch) // ch
    x = oracle_mod.fill_i16(frames * ch, 43_000 + k % 1000 + ch)
    with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
        assert plan.info.path == 1 and plan.info.mode == 5
        y = plan.run_host(x)
    assert np.array_equal(y, oracle_mod.mavg_i16_mt(x, k, ch, 8))


def test_far_lag_i16_slices_shards_and_planar_bit_identical(mavg, oracle_mod, torch_cuda):
    """int32 window sums are exact, so -- unlike the float32 far-lag kernel -- sliced host runs, shards with their
    context in front of them or in another allocation, and the whole run agree bit for bit."""
    torch = torch_cuda
    ch, k = 2, 30_000
    frames = (1 << 22) + 333
    x = oracle_mod.fill_i16(frames * ch, 44_000)
    e = oracle_mod.mavg_i16_mt(x, k, ch, 8)
    with mavg.Plan(frames, k, channels=ch, dtype="i16", slice_bytes=1 << 20) as plan:
        assert plan.info.mode == 5
        assert np.array_equal(plan.run_host(x), e)
        assert plan.info.launches_per_run >= 4
        halo = int(plan.info.halo_frames)
    assert halo >= k
    dx = torch.from_numpy(x).cuda()
    cut = halo + 5 * 8192 + 1000                         # not a multiple of a tile: the shard has its own tile grid
    dz = torch.zeros((frames - cut) * ch, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype="i16", first_frame=cut) as plan:
        plan.run_device_halo(dx.data_ptr() + 2 * cut * ch, dz.data_ptr(), dx.data_ptr() + 2 * (cut - halo) * ch)
        plan.synchronize()
        assert np.array_equal(dz.cpu().numpy(), e[cut * ch:])
        d_ctx = torch.from_numpy(x[(cut - halo) * ch:cut * ch].copy()).cuda()      # context in its own allocation
        d_in = torch.from_numpy(x[cut * ch:].copy()).cuda()
        dz.zero_()
        torch.cuda.synchronize()
        plan.run_device_halo(d_in.data_ptr(), dz.data_ptr(), d_ctx.data_ptr())
        plan.synchronize()
        assert plan.info.launches_per_run >= 2
        assert np.array_equal(dz.cpu().numpy(), e[cut * ch:])
    # saturated input: window sums of 32768 k in magnitude
    for val in (-32768, 32767):
        xs = np.full(300_000 * 4, val, dtype=np.int16)
        for kk in (20_000, 46_340):
            assert np.array_equal(mavg.moving_average(xs, kk, channels=4), oracle_mod.mavg_i16(xs, kk, 4)), (val, kk)
    # window longer than the signal, and a planar batch of three mono signals
    xs = oracle_mod.fill_i16(70_000, 44_001)
    assert np.array_equal(mavg.moving_average(xs, 45_000), oracle_mod.mavg_i16(xs, 45_000))
    xp = oracle_mod.fill_i16(3 * 200_000, 44_002)
    with mavg.Plan(200_000, 40_000, channels=3, layout="planar", dtype="i16") as plan:
        assert plan.info.mode == 5
        yp = plan.run_host(xp)
    for c in range(3):
        seg = slice(c * 200_000, (c + 1) * 200_000)
        assert np.array_equal(yp[seg], oracle_mod.mavg_i16(xp[seg], 40_000))


@pytest.mark.parametrize("dtype,ch,k", [("i16", 2, 50_000), ("i16", 1, 70_000), ("i16", 5, 12_000), ("f32", 3, 20_000),
                                        ("i16", 2, 3_000_000)])
def test_prefix_difference_path_far_windows(mavg, oracle_mod, torch_cuda, dtype, ch, k):
    """Far windows that no streaming kernel takes run as single-pass prefix sum + difference (info.mode 7): whole
    signal, sliced host run, and a shard whose left context lives in another allocation -- int16 bit-exact."""
    torch = torch_cuda
    frames = 1_200_000 // ch + 13
    n = frames * ch
    x = oracle_mod.fill_i16(n, 41000 + k % 97) if dtype == "i16" else oracle_mod.fill_f32(n, 41000 + k % 97, oracle_mod.DIST_USYM)
    e = oracle_mod.mavg_i16(x, k, ch) if dtype == "i16" else oracle_mod.mavg_f64(x, k, ch)

    def check(y, exp):
        if dtype == "i16":
            assert np.array_equal(y, exp)
        else:
            scale = oracle_mod.mavg_f64(np.abs(x), k, ch)[-exp.size:]
            assert np.max(np.abs(y - exp) / np.maximum(scale, 1e-30)) < TOL

    with mavg.Plan(frames, k, channels=ch, dtype=dtype) as plan:
        assert plan.info.path == 2 and plan.info.mode == 7
        check(plan.run_host(x), e)
    with mavg.Plan(frames, k, channels=ch, dtype=dtype, slice_bytes=1 << 20) as plan:      # several slices, each with a context
        check(plan.run_host(x), e)
    # shard with a separate context buffer
    with mavg.Plan(frames, k, channels=ch, dtype=dtype) as probe:
        halo = int(probe.info.halo_frames)
    cut = min(frames // 2, max(halo, 1000) + 7) if halo < frames // 2 else frames // 3
    tdt = torch.int16 if dtype == "i16" else torch.float32
    lo = max(0, cut - halo)
    ctx = np.zeros(halo * ch, dtype=x.dtype)                      # context shorter than the halo: zeros in front (signal start)
    ctx[(halo - (cut - lo)) * ch:] = x[lo * ch:cut * ch]
    d_ctx = torch.from_numpy(ctx).cuda()
    d_in = torch.from_numpy(x[cut * ch:].copy()).cuda()
    d_out = torch.zeros((frames - cut) * ch, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    with mavg.Plan(frames - cut, k, channels=ch, dtype=dtype, first_frame=cut) as plan:
        plan.run_device_halo(d_in.data_ptr(), d_out.data_ptr(), d_ctx.data_ptr())
        plan.synchronize()
    check(d_out.cpu().numpy(), e[cut * ch:])


# ------------------------------------------------------------------ size-independent properties at full size, no oracle in the loop
@pytest.mark.parametrize("case", [("i16", 2, 1 << 27, 4096, 12_345), ("i16", 6, (1 << 27) // 6, 64, 1_001), ("i16", 8, 1 << 24, 1000, 77),
                                  ("i16", 64, 1 << 21, 64, 333), ("i16", 256, 1 << 19, 8, 50), ("i16", 5, (1 << 26) // 5, 300, 4_097),
                                  ("i16", 1, 1 << 27, 32768, 9)])
def test_full_size_shift_invariance_and_impulse_response_i16(mavg, torch_cuda, case):
    """A causal filter commutes with a delay: filtering the signal delayed by d frames (zeros shifted in) gives the
    output delayed by d frames -- bit for bit for int16, whatever tile, run or tile-range boundary the delay moves the
    samples across (d is never a multiple of a tile).  And the response to a single full-scale impulse is
    trunc(32767 / k) for exactly k frames.  Both checked on the device over the whole 2^27-sample signal."""
    torch = torch_cuda
    dtype, ch, frames, k, d = case
    n = frames * ch
    x = torch.empty(n, dtype=torch.int16, device="cuda")
    mavg.fill_synthetic_device(x.data_ptr(), "i16", n, 0, 4242 + k)
    x2 = torch.zeros(n, dtype=torch.int16, device="cuda")
    x2[d * ch:] = x[:n - d * ch]
    y, y2 = torch.empty_like(x), torch.empty_like(x)
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch, dtype="i16") as plan:
        assert plan.info.path == 1
        plan.run_device([x.data_ptr()], [y.data_ptr()])
        plan.run_device([x2.data_ptr()], [y2.data_ptr()])
        plan.synchronize()
        assert int(torch.count_nonzero(y2[:d * ch])) == 0
        assert torch.equal(y2[d * ch:], y[:n - d * ch])
        # impulse of 32767 in channel c at frame f0 = c * 1000 + 123_457 (mod frames)
        x2.zero_()
        pos = [((c * 1000 + 123_457) % (frames - k - 1), c) for c in range(ch)]
        for f0, c in pos:
            x2[f0 * ch + c] = 32767
        plan.run_device([x2.data_ptr()], [y2.data_ptr()])
        plan.synchronize()
    yv = y2.view(frames, ch)
    for f0, c in pos:
        col = yv[:, c]
        assert int(torch.count_nonzero(col)) == (k if 32767 // k else 0)
        assert bool(torch.all(col[f0:f0 + k] == 32767 // k))


@pytest.mark.parametrize("case", [(1, 1 << 27, 4096, 12_345), (2, 1 << 26, 100, 7), (1, 1 << 27, 60_000, 1_001), (6, (1 << 26) // 6, 64, 99),
                                  (256, 1 << 18, 64, 11)])
def test_full_size_shift_invariance_f32(mavg, torch_cuda, case):
    """The same delay property for float32, where a delay that is not a multiple of a tile changes the order of the
    additions: the two outputs agree to 1e-5 relative (inputs in [0.5, 1.5), so every window sum is far from zero)."""
    torch = torch_cuda
    ch, frames, k, d = case
    n = frames * ch
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    mavg.fill_synthetic_device(x.data_ptr(), "f32", n, 0, 4343 + k)
    x += 0.5
    x2 = torch.zeros(n, dtype=torch.float32, device="cuda")
    x2[d * ch:] = x[:n - d * ch]
    y, y2 = torch.empty_like(x), torch.empty_like(x)
    torch.cuda.synchronize()
    with mavg.Plan(frames, k, channels=ch) as plan:
        plan.run_device([x.data_ptr()], [y.data_ptr()])
        plan.run_device([x2.data_ptr()], [y2.data_ptr()])
        plan.synchronize()
    assert float(y2[:d * ch].abs().max()) == 0.0
    a, b = y2[d * ch:], y[:n - d * ch]
    assert float(((a - b).abs() / b.abs().clamp_min(1e-30)).max()) < TOL
