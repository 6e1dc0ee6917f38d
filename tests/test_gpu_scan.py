"""GPU tests of mavg_prefix_sum, the single-pass decoupled-look-back prefix-sum primitive (what the
reference's recursive_hillis_steele / recursive_blelloch compute: basics/hillis_steele_averager.cu:69-84,
basics/blelloch_scan_averager.cu:134-167).  int16 -> int64 is checked bit-exactly against numpy's int64
cumsum; float32 -> float64 within 1e-12 relative (fp64 association differs between tiles)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _ref(x, ch):
    return np.cumsum(x.reshape(-1, ch).astype(np.int64 if x.dtype == np.int16 else np.float64), axis=0).reshape(-1)


@pytest.mark.parametrize("ch", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("frames", [1, 15, 16, 17, 4096, 4097, 100_000, 1_000_003])
def test_prefix_sum_i16_exact(mavg, oracle_mod, torch_cuda, ch, frames):
    torch = torch_cuda
    x = oracle_mod.fill_i16(frames * ch, 21000 + frames + ch)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(frames * ch, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    mavg.prefix_sum_device(dx.data_ptr(), dy.data_ptr(), "i16", frames, ch)
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), _ref(x, ch))


@pytest.mark.parametrize("ch", [1, 2, 3, 6, 7, 8])
@pytest.mark.parametrize("frames", [33, 4096 * 3 + 5, 3_000_001])
def test_prefix_sum_f32(mavg, oracle_mod, torch_cuda, ch, frames):
    torch = torch_cuda
    x = oracle_mod.fill_f32(frames * ch, 22000 + frames + ch, oracle_mod.DIST_USYM)
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(frames * ch, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    mavg.prefix_sum_device(dx.data_ptr(), dy.data_ptr(), "f32", frames, ch)
    torch.cuda.synchronize()
    e = _ref(x, ch)
    scale = np.cumsum(np.abs(x.reshape(-1, ch).astype(np.float64)), axis=0).reshape(-1)
    assert np.max(np.abs(dy.cpu().numpy() - e) / scale) < 1e-12


def test_prefix_sum_large_and_prefix_difference_equals_moving_average(mavg, oracle_mod, torch_cuda):
    """2^27 samples (32768 tiles of look-back), and the reference's scan-binary identity:
    (P[i] - P[i-k]) / k, truncated, equals the moving average (hillis_steele_averager.cu:87-100 done exactly)."""
    torch = torch_cuda
    n, k = 1 << 27, 1000
    dx = torch.empty(n, dtype=torch.int16, device="cuda")
    mavg.fill_synthetic_device(dx.data_ptr(), "i16", n, 0, 77)
    dp = torch.empty(n, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    mavg.prefix_sum_device(dx.data_ptr(), dp.data_ptr(), "i16", n, 1)
    torch.cuda.synchronize()
    assert torch.equal(dp, torch.cumsum(dx.to(torch.int64), 0))
    w = dp.clone()
    w[k:] -= dp[:-k]
    y_scan = torch.div(w, k, rounding_mode="trunc").to(torch.int16)
    dy = torch.empty_like(dx)
    torch.cuda.synchronize()
    with mavg.Plan(n, k, dtype="i16") as plan:
        plan.run_device([dx.data_ptr()], [dy.data_ptr()])
        plan.synchronize()
    assert torch.equal(dy, y_scan)


def test_prefix_sum_saturated_input_many_chunks(mavg, torch_cuda):
    """All samples at -32768 / 32767: chunk aggregates at their extreme (16384 * 32768 = 2^29 fits the int32 chunk-local
    type), prefixes far beyond 2^31; 3 channels so that chunk boundaries fall inside odd-length runs."""
    torch = torch_cuda
    for ch, val in ((1, -32768), (3, 32767), (2, -32768)):
        frames = (1 << 22) // ch + 7
        dx = torch.full((frames * ch,), val, dtype=torch.int16, device="cuda")
        dy = torch.zeros(frames * ch, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        mavg.prefix_sum_device(dx.data_ptr(), dy.data_ptr(), "i16", frames, ch)
        torch.cuda.synchronize()
        e = (torch.arange(frames, device="cuda", dtype=torch.int64) + 1).repeat_interleave(ch) * val
        assert torch.equal(dy, e)


def test_prefix_sum_repeated_calls_reuse_pool_scratch(mavg, torch_cuda):
    """Scratch comes from a library-owned stream-ordered pool (the device's default pool is not touched): repeated
    calls on one stream work back to back."""
    torch = torch_cuda
    dx = torch.zeros(1 << 16, dtype=torch.int16, device="cuda")
    dy = torch.zeros(1 << 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        mavg.prefix_sum_device(dx.data_ptr(), dy.data_ptr(), "i16", 1 << 16, 1)
    torch.cuda.synchronize()
    assert int(dy[-1]) == 0


def test_prefix_sum_argument_errors(mavg):
    with pytest.raises(mavg.MavgError) as e:
        mavg.prefix_sum_device(0, 0, "i16", 10, 9)       # 1..8 interleaved channels
    assert e.value.status == -2
    with pytest.raises(mavg.MavgError) as e:
        mavg.prefix_sum_device(0, 0, "i16", 10, 0)
    assert e.value.status == -2
    mavg.prefix_sum_device(0, 0, "i16", 0, 1)   # empty: nothing to do
