"""Moving RMS (mavg_op MAVG_OP_RMS; SURVEY.md section 8(f) row 4: the windowed reduction beside the average).
float32 on the TMA streaming kernel (squares on load, root on store) against the fp64 oracle within 1e-5 relative;
int16 and every other shape on the generic kernel, int16 bit-exact against oracle_mrms_i16."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(y, e):
    return float(np.max(np.abs(y.astype(np.float64) - e) / np.maximum(np.abs(e), 1e-30)))


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 9, 16, 17, 64, 255, 256, 257, 1000, 1024, 4095, 4096, 20000])
@pytest.mark.parametrize("ch", [1, 2])
def test_rms_f32_stream_vs_fp64_oracle(mavg, oracle_mod, k, ch):
    frames = (5 * 8192 + 77) // ch + 3                      # several tiles, ragged tail
    for dist in (oracle_mod.DIST_U01, oracle_mod.DIST_USYM):
        x = oracle_mod.fill_f32(frames * ch, 31000 + k + ch, dist)
        with mavg.Plan(frames, k, channels=ch, op="rms") as plan:
            if k * ch <= 8192:      # beyond, the ring of squares no longer fits shared memory: generic kernel
                assert plan.info.path == 1, "float32 mono/stereo RMS must take the TMA streaming kernel"
            y = plan.run_host(x)
        e = oracle_mod.mrms_f64(x, k, ch)
        # U[-1,1): an RMS is a root of a sum of squares, never near zero by cancellation
        assert _rel(y, e) < TOL, (k, ch, dist)


def test_rms_constant_signal_is_its_magnitude(mavg):
    n, k = 1 << 22, 300
    x = np.full(n, -0.75, dtype=np.float32)
    y = mavg.moving_rms(x, k)
    assert np.all(np.abs(y[k:] - 0.75) < 1e-6)
    # warm-up divides by the full k like the average does: sqrt(j/k) * |c|
    j = np.arange(1, k + 1, dtype=np.float64)
    assert np.max(np.abs(y[:k] - 0.75 * np.sqrt(j / k))) < 1e-6


def test_rms_f32_large_offset_signal(mavg, oracle_mod):
    """1e4 DC offset: squares of 1e8, window sums of 1e11 -- the tile-local rebasing keeps the slide accurate."""
    n, k = 1 << 21, 4096
    x = oracle_mod.fill_f32(n, 5, oracle_mod.DIST_DC1E4)
    y = mavg.moving_rms(x, k)
    assert _rel(y, oracle_mod.mrms_f64(x, k)) < TOL


def test_rms_planar_batch_and_sharded_bit_identity(mavg, oracle_mod):
    ch, frames, k = 6, 3 * 8192 + 5, 100
    x = oracle_mod.fill_f32(ch * frames, 77, oracle_mod.DIST_USYM)
    with mavg.Plan(frames, k, channels=ch, layout="planar", op="rms") as plan:
        assert plan.info.path == 1
        y = plan.run_host(x)
    for c in range(ch):
        e = oracle_mod.mrms_f64(x[c * frames:(c + 1) * frames], k)
        assert _rel(y[c * frames:(c + 1) * frames], e) < TOL
    # slices of mavg_run_host never change a bit
    n = 1 << 23
    x = oracle_mod.fill_f32(n, 78)
    with mavg.Plan(n, 1000, op="rms", slice_bytes=4 << 20) as a, mavg.Plan(n, 1000, op="rms", slice_bytes=64 << 20) as b:
        assert np.array_equal(a.run_host(x), b.run_host(x))


@pytest.mark.parametrize("ch,k", [(1, 1), (1, 5), (2, 41), (2, 1000), (3, 64), (6, 300), (40, 17)])
def test_rms_i16_bit_exact(mavg, oracle_mod, ch, k):
    frames = 70_000 // ch + 11
    x = oracle_mod.fill_i16(frames * ch, 900 + k + ch)
    y = mavg.moving_average(x, k, channels=ch, op="rms")
    assert np.array_equal(y, oracle_mod.mrms_i16(x, k, ch))


def test_rms_i16_saturates_at_full_scale(mavg, oracle_mod):
    x = np.full(50_000, -32768, dtype=np.int16)
    for k in (1, 7, 4096):
        y = mavg.moving_rms(x, k)
        assert np.array_equal(y, oracle_mod.mrms_i16(x, k))
        assert y[-1] == 32767


@pytest.mark.parametrize("ch,k", [(3, 100), (5, 3), (34, 64), (64, 2048)])
def test_rms_f32_other_shapes_take_the_generic_kernel(mavg, oracle_mod, ch, k):
    frames = 40_000 // ch + 9
    x = oracle_mod.fill_f32(frames * ch, 1200 + k + ch, oracle_mod.DIST_USYM)
    with mavg.Plan(frames, k, channels=ch, op="rms") as plan:
        assert plan.info.path == 2
        y = plan.run_host(x)
    assert _rel(y, oracle_mod.mrms_f64(x, k, ch)) < TOL


def test_rms_unknown_op_is_rejected(mavg):
    import ctypes
    from digital_signal_processsing_b200 import _lib
    d = _lib.Desc()
    d.struct_size = ctypes.sizeof(_lib.Desc)
    d.channels, d.frames, d.window, d.op = 1, 100, 3, 7
    h = ctypes.c_void_p()
    assert _lib.load().mavg_plan_create(ctypes.byref(d), ctypes.byref(h)) == _lib.ERR_INVALID_ARG
