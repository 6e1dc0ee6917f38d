"""CPU tests of the oracle: pinned against the reference's own outputs and exact arithmetic.

Reference followed: basics/profilable_moving_averager.cpp:14-37.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases(npz):
    names = sorted({k.rsplit("__", 1)[0] for k in npz.files})
    for n in names:
        frames, ch, k = (int(v) for v in npz[n + "__meta"])
        yield n, npz[n + "__x"], npz[n + "__y"], frames, ch, k


def test_known_answer(oracle_mod):
    # derived from the definition: x=1..6, k=3 -> trunc([1,3,6,9,12,15]/3)
    x = np.arange(1, 7, dtype=np.int16)
    assert oracle_mod.mavg_i16(x, 3).tolist() == [0, 1, 2, 3, 4, 5]
    # truncation toward zero for negative sums (C++ integer division), not floor
    x = np.array([-1, -1, -1, -5], dtype=np.int16)
    assert oracle_mod.mavg_i16(x, 2).tolist() == [0, -1, -1, -3]


def test_i16_matches_reference_golden(oracle_mod):
    g = np.load(os.path.join(GOLD, "golden_i16.npz"))
    n = 0
    for name, x, y, frames, ch, k in _cases(g):
        got = oracle_mod.mavg_i16(x, k, ch)
        assert np.array_equal(got, y), name
        n += 1
    assert n >= 12


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libref_cpu.so")),
                    reason="oracle/_ref not built (needs /root/reference once)")
def test_i16_matches_reference_live(oracle_mod):
    """The restatement against the UNMODIFIED reference function, fresh random inputs."""
    rng = np.random.default_rng(1234)
    for ch in (1, 2, 3):
        for k in (1, 2, 3, 7, 41, 64, 850, 1000):
            frames = int(rng.integers(k, k + 3000))
            x = rng.integers(-32768, 32768, size=frames * ch, dtype=np.int64).astype(np.int16)
            assert np.array_equal(oracle_mod.mavg_i16(x, k, ch), oracle_mod.ref_mavg_i16(x, k, ch)), (ch, k)


def test_f64_matches_exact_golden(oracle_mod):
    g = np.load(os.path.join(GOLD, "golden_f32.npz"))
    for name, x, y, frames, ch, k in _cases(g):
        got = oracle_mod.mavg_f64(x, k, ch)
        np.testing.assert_allclose(got, y, rtol=1e-13, atol=1e-13, err_msg=name)


def test_threaded_variants_agree(oracle_mod):
    x16 = oracle_mod.fill_i16(50_000 * 2, 5)
    for k in (1, 5, 977):
        assert np.array_equal(oracle_mod.mavg_i16_mt(x16, k, 2, 4), oracle_mod.mavg_i16(x16, k, 2))
    xf = oracle_mod.fill_f32(60_001, 6)
    y1 = oracle_mod.mavg_f32_running(xf, 64)
    y4 = oracle_mod.mavg_f32_running(xf, 64, threads=4)
    np.testing.assert_allclose(y4, y1, rtol=2e-5)


def test_fp32_port_is_close_but_not_the_oracle(oracle_mod):
    x = oracle_mod.fill_f32(200_000, 9)
    y = oracle_mod.mavg_f32_running(x, 16)
    e = oracle_mod.mavg_f64(x, 16)
    assert np.max(np.abs(y - e)) < 1e-4


def test_properties(oracle_mod):
    x = oracle_mod.fill_f32(5000, 3, oracle_mod.DIST_USYM)
    z = oracle_mod.fill_f32(5000, 4, oracle_mod.DIST_USYM)
    k = 33
    # linearity
    np.testing.assert_allclose(oracle_mod.mavg_f64((x + z).astype(np.float32), k),
                               oracle_mod.mavg_f64(x, k) + oracle_mod.mavg_f64(z, k), atol=1e-6)
    # constant signal: ramps i/k then constant
    c = np.full(200, 2.0, dtype=np.float32)
    y = oracle_mod.mavg_f64(c, 8)
    np.testing.assert_allclose(y[:8], 2.0 * np.arange(1, 9) / 8)
    np.testing.assert_allclose(y[8:], 2.0)
    # channel independence
    inter = np.empty(2 * 5000, dtype=np.float32)
    inter[0::2], inter[1::2] = x, z
    yi = oracle_mod.mavg_f64(inter, k, 2)
    np.testing.assert_allclose(yi[0::2], oracle_mod.mavg_f64(x, k))
    np.testing.assert_allclose(yi[1::2], oracle_mod.mavg_f64(z, k))
    # k = 1 is the identity, k > frames is pure warm-up
    np.testing.assert_array_equal(oracle_mod.mavg_f64(x, 1), x.astype(np.float64))
    np.testing.assert_allclose(oracle_mod.mavg_f64(x[:10], 50), np.cumsum(x[:10].astype(np.float64)) / 50)


def test_generator_is_counter_based(oracle_mod):
    a = oracle_mod.fill_f32(1000, 77)
    b = oracle_mod.fill_f32(400, 77, first_index=600)
    assert np.array_equal(a[600:], b)
    assert a.min() >= 0.0 and a.max() < 1.0
    s = oracle_mod.fill_f32(1000, 77, oracle_mod.DIST_USYM)
    assert s.min() >= -1.0 and s.max() < 1.0
    i = oracle_mod.fill_f32(1000, 77, oracle_mod.DIST_I16)
    assert np.array_equal(i, np.round(i)) and i.min() >= -32768 and i.max() <= 32767
    assert np.array_equal(oracle_mod.fill_i16(1000, 77).astype(np.float32), i)
    # point evaluation agrees with the array evaluation
    x = oracle_mod.fill_f32(3000, 5)
    e = oracle_mod.mavg_f64(x, 100)
    for idx in (0, 50, 99, 100, 2999):
        assert abs(oracle_mod.point_f64(idx, 100, 5) - e[idx]) < 1e-12


def test_moving_rms_oracle_against_numpy(oracle_mod):
    """oracle_mrms_*: sqrt of the windowed mean of squares, same window / padding / divide-by-k conventions as a1."""
    import math
    x = oracle_mod.fill_f32(6000, 3, oracle_mod.DIST_USYM)
    for ch, k in ((1, 1), (1, 7), (2, 64), (3, 5000)):
        sq = x[: (x.size // ch) * ch].astype(np.float64).reshape(-1, ch) ** 2
        e = np.array([sq[max(0, i + 1 - k):i + 1].sum(axis=0) / k for i in range(sq.shape[0])])   # fresh window sums
        y = oracle_mod.mrms_f64(x[: sq.size], k, ch)
        assert np.max(np.abs(y - np.sqrt(e).reshape(-1))) < 1e-12
    xi = oracle_mod.fill_i16(9000, 4)
    for ch, k in ((1, 3), (2, 100), (3, 4000)):
        sq = xi.astype(np.int64).reshape(-1, ch) ** 2
        c = np.cumsum(np.vstack([np.zeros((1, ch), dtype=np.int64), sq]), axis=0)
        e = np.array([[min(32767, math.isqrt(int(v) // k)) for v in (c[i + 1] - c[max(0, i + 1 - k)])]
                      for i in range(sq.shape[0])]).reshape(-1)      # floor(sqrt(w / k)) == isqrt(w // k)
        assert np.array_equal(oracle_mod.mrms_i16(xi, k, ch).astype(np.int64), e)
    assert oracle_mod.mrms_i16(np.full(10, -32768, dtype=np.int16), 1)[0] == 32767
