"""Randomised differential test: libmavg through the C ABI against the CPU oracle on seeded random shapes.

Every streaming kernel and the generic kernel are reached by construction (the plan picks the path from dtype,
channel count, layout and window); shapes are small enough for the oracle to check EVERY output.  Seeds are fixed:
a failure prints the case so that it can be replayed.  int16: bit-exact.  float32: <= 1e-5 relative on U[0,1)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _draw(rng):
    dtype = "f32" if rng.random() < 0.5 else "i16"
    u = rng.random()
    if u < 0.30:
        ch = int(rng.integers(1, 3))                 # mono / stereo streaming kernels
    elif u < 0.65:
        ch = int(rng.integers(3, 32))                # few-channel kernels
    elif u < 0.85:
        ch = int(rng.choice([32, 36, 40, 64, 96, 128]))   # column kernel (float32) / generic (int16)
    else:
        ch = int(rng.integers(33, 70))               # generic for most
    layout = "planar" if (ch > 1 and rng.random() < 0.15) else "interleaved"
    v = rng.random()
    if v < 0.35:
        k = int(rng.integers(1, 18))
    elif v < 0.70:
        k = int(rng.integers(18, 300))
    elif v < 0.93:
        k = int(rng.integers(300, 5000))
    else:
        k = int(rng.integers(5000, 40000))
    budget = int(rng.choice([3_000, 40_000, 400_000, 1_500_000]))     # total samples
    frames = max(1, budget // ch + int(rng.integers(0, 97)))
    return dtype, ch, layout, k, frames


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_against_oracle(mavg, oracle_mod, seed):
    rng = np.random.default_rng(1000 + seed)
    seen = set()
    for case in range(25):
        dtype, ch, layout, k, frames = _draw(rng)
        n = frames * ch
        x = oracle_mod.fill_f32(n, 50_000 + 100 * seed + case) if dtype == "f32" else \
            oracle_mod.fill_i16(n, 50_000 + 100 * seed + case)
        with mavg.Plan(frames, k, channels=ch, dtype=dtype, layout=layout) as plan:
            y = plan.run_host(x)
            seen.add((int(plan.info.path), int(plan.info.mode)))
        tag = (seed, case, dtype, ch, layout, k, frames)
        if layout == "planar":
            for c in range(ch):
                seg = slice(c * frames, (c + 1) * frames)
                if dtype == "f32":
                    e = oracle_mod.mavg_f64(x[seg], k)
                    assert np.max(np.abs(y[seg] - e) / np.abs(e)) < TOL, tag
                else:
                    assert np.array_equal(y[seg], oracle_mod.mavg_i16(x[seg], k)), tag
        elif dtype == "f32":
            e = oracle_mod.mavg_f64(x, k, ch)
            assert np.max(np.abs(y - e) / np.abs(e)) < TOL, tag
        else:
            assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch)), tag
    assert len(seen) >= 3, seen          # several kernels were exercised by this seed


@pytest.mark.parametrize("seed", range(6))
def test_random_shards_with_halo_against_oracle(mavg, oracle_mod, seed):
    """Shard plans (first_frame > 0) fed through mavg_run_host with the left context in front of the buffer:
    any cut that respects the plan's alignment reproduces the whole-signal result."""
    rng = np.random.default_rng(2000 + seed)
    ran = 0
    for case in range(10):
        dtype, ch, layout, k, frames = _draw(rng)
        if layout == "planar":
            layout = "interleaved"
        frames = max(frames, 2000)
        n = frames * ch
        x = oracle_mod.fill_f32(n, 60_000 + 100 * seed + case) if dtype == "f32" else \
            oracle_mod.fill_i16(n, 60_000 + 100 * seed + case)
        with mavg.Plan(frames, k, channels=ch, dtype=dtype) as whole:
            align = int(whole.info.halo_frames) // max(1, int(whole.info.history_tiles)) if whole.info.path == 1 else 1
            halo = int(whole.info.halo_frames)
        if align <= 0:
            align = 1
        halo_up = (halo + align - 1) // align * align
        if halo_up + align >= frames:
            continue
        cut = halo_up + int(rng.integers(0, (frames - halo_up) // align)) * align
        if cut >= frames:
            continue
        with mavg.Plan(frames - cut, k, channels=ch, dtype=dtype, first_frame=cut) as plan:
            if int(plan.info.halo_frames) > cut:
                continue
            h = int(plan.info.halo_frames)
            buf = np.ascontiguousarray(x[(cut - h) * ch:])
            y = plan.run_host_with_context(buf, h)
        tag = (seed, case, dtype, ch, k, frames, cut)
        if dtype == "f32":
            e = oracle_mod.mavg_f64(x, k, ch)[cut * ch:]
            assert np.max(np.abs(y - e) / np.abs(e)) < TOL, tag
        else:
            assert np.array_equal(y, oracle_mod.mavg_i16(x, k, ch)[cut * ch:]), tag
        ran += 1
    assert ran >= 5, ran


@pytest.mark.parametrize("seed", range(4))
def test_random_shapes_on_two_shards_bit_identical(mavg, oracle_mod, seed):
    """One process driving two shards (two GPUs when present, else the same GPU twice): the result is the
    single-shard result bit for bit on every streaming path (tile grid anchored at frame 0) and for int16 everywhere
    (exact arithmetic).  The generic float32 kernel builds its run-start sums in fp64 from differently partitioned
    partial sums on a shard (and so does the far-lag kernel with its carried window sum), so those are held to 1e-6
    of the single-shard result."""
    rng = np.random.default_rng(3000 + seed)
    devs = [0, 1] if mavg.device_count() >= 2 else [0, 0]
    bad = []
    for case in range(12):
        dtype, ch, layout, k, frames = _draw(rng)
        n = frames * ch
        x = oracle_mod.fill_f32(n, 70_000 + 100 * seed + case) if dtype == "f32" else \
            oracle_mod.fill_i16(n, 70_000 + 100 * seed + case)
        with mavg.Plan(frames, k, channels=ch, dtype=dtype, layout=layout) as plan:
            y1 = plan.run_host(x)
            stream = int(plan.info.path) == 1 and int(plan.info.mode) != 5   # the far-lag kernel carries fp64 sums per chunk
        try:
            with mavg.Plan(frames, k, channels=ch, dtype=dtype, layout=layout, devices=devs) as plan:
                y2 = plan.run_host(x)
        except Exception as exc:                     # a shard must hold its neighbour's context: refused, not wrong
            assert "too short to shard" in str(exc), exc
            continue
        if dtype == "i16" or stream:
            ok = np.array_equal(y1, y2)
        else:
            ok = bool(np.max(np.abs(y1 - y2) / np.abs(y1)) < 1e-6)
        if not ok:
            bad.append((seed, case, dtype, ch, layout, k, frames, stream, int(np.count_nonzero(y1 != y2))))
    assert not bad, bad


@pytest.mark.parametrize("seed", range(4))
def test_random_tiny_and_degenerate_shapes(mavg, oracle_mod, torch_cuda_mod, seed):
    """Signals shorter than a tile, a row or the window itself; windows longer than the signal; device pointers
    that are not 16-byte aligned (the streaming kernels hand those to the generic kernel)."""
    torch = torch_cuda_mod
    rng = np.random.default_rng(4000 + seed)
    for case in range(40):
        dtype = "f32" if rng.random() < 0.5 else "i16"
        ch = int(rng.integers(1, 41))
        frames = int(rng.integers(1, 300))
        k = int(rng.integers(1, 600))
        layout = "planar" if (ch > 1 and rng.random() < 0.2) else "interleaved"
        n = frames * ch
        x = oracle_mod.fill_f32(n, 80_000 + 100 * seed + case) if dtype == "f32" else \
            oracle_mod.fill_i16(n, 80_000 + 100 * seed + case)
        tag = (seed, case, dtype, ch, layout, k, frames)

        def expect(seg, c):
            return oracle_mod.mavg_f64(seg, k, c) if dtype == "f32" else oracle_mod.mavg_i16(seg, k, c)

        def check(y):
            parts = [(slice(c * frames, (c + 1) * frames), 1) for c in range(ch)] if layout == "planar" else [(slice(0, n), ch)]
            for seg, c in parts:
                e = expect(x[seg], c)
                if dtype == "f32":
                    assert np.max(np.abs(y[seg] - e) / np.abs(e)) < TOL, tag
                else:
                    assert np.array_equal(y[seg], e), tag

        with mavg.Plan(frames, k, channels=ch, dtype=dtype, layout=layout) as plan:
            check(plan.run_host(x))
            # the same plan on device buffers shifted off 16-byte alignment by one element
            tdt = torch.float32 if dtype == "f32" else torch.int16
            dx = torch.zeros(n + 8, dtype=tdt, device="cuda")
            dy = torch.zeros(n + 8, dtype=tdt, device="cuda")
            dx[1:n + 1] = torch.from_numpy(x).cuda()
            es = 4 if dtype == "f32" else 2
            plan.run_device([dx.data_ptr() + es], [dy.data_ptr() + es])
            plan.synchronize()
            check(dy[1:n + 1].cpu().numpy())
            assert float(dy[0]) == 0 and float(dy[n + 1]) == 0, tag      # nothing written outside the signal


@pytest.fixture(scope="module")
def torch_cuda_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch
