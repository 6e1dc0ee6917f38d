"""CPU check of the streaming kernel's arithmetic (tests/algo_model.py) against exact sums.

Tolerance: the north-star bound, max relative error 1e-5 against fp64 on U[0,1) input.
"""
import numpy as np
import pytest

from algo_model import geometry, stream_model

TOL = 1e-5


def _exact(x, k):
    s = np.round(x.astype(np.float64) * 2.0**24).astype(np.int64)
    c = np.cumsum(s)
    w = c.copy()
    if len(x) > k:
        w[k:] -= c[:-k]
    return w / (k * 2.0**24)


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 9, 16, 17, 63, 64, 255, 256, 257, 1000, 1024, 4096, 4097])
@pytest.mark.parametrize("shape", [(256, 16), (256, 32)])
def test_model_within_tolerance(oracle_mod, k, shape):
    NT, R = shape
    x = oracle_mod.fill_f32(2 * NT * R + 1234, 100 + k)
    y = stream_model(x, k, NT, R)
    e = _exact(x, k)
    rel = np.max(np.abs(y - e) / np.abs(e))
    assert rel < TOL, (k, rel)


def test_geometry_identities():
    for R in (16, 32):
        for k in range(1, 600):
            g = geometry(k, 256, R)
            if g["mode"] == 2:
                continue
            # whole groups + tail of the lag group make exactly k samples
            assert g["n_full"] * R + g["m_part"] == k
            assert 1 <= g["m_part"] <= R
            assert 0 <= g["mis"] <= 3 and (g["mis"] + k) % 4 == 0
            assert g["H"] * g["T"] >= k


def _mulhi_s32(a, m):
    return (a * m) >> 32            # numpy int64 / Python ints: arithmetic shift = floor, like mul.hi.s32


def test_int16_mulhi_division_is_exact():
    """Host restatement of the int16 kernel's division (plan_stream_i16 / div_trunc_mulhi in csrc):
        trunc(w / k) == (t >> s) + (t >>> 31),  t = mulhi_s32(w, M),  M = floor(2^(30+L) / k) + 1,  s = L - 2
    with L = ceil(log2 k), for every k in 3..32768 and every reachable window sum |w| <= 32768 k.  The quotient
    only changes at multiples of k, so every multiple and both neighbours are checked for the k listed, plus
    random sums.  k == 2 runs with doubled weights (w -> 2 w) and the constants of k = 4."""
    rng = np.random.default_rng(5)
    ks = list(range(3, 300)) + [511, 512, 513, 1000, 1023, 1024, 1025, 4095, 4096, 4097, 16383, 16384, 16385,
                                30000, 32767, 32768]

    def check(k, w, scale=1):
        kd = k * scale
        L = (kd - 1).bit_length()
        M = (1 << (30 + L)) // kd + 1
        assert L >= 2 and M < 2**31
        ws = w * scale
        assert np.abs(ws).max() < 2**31
        t = _mulhi_s32(ws, M)
        got = (t >> (L - 2)) + ((t >> 31) & 1)
        want = np.sign(w) * (np.abs(w) // k)
        assert np.array_equal(got, want), k

    for k in ks:
        q = np.arange(-32768, 32768, dtype=np.int64) * k
        w = np.concatenate([q, q - 1, q + 1, q + k - 1, np.array([-32768 * k, 32767 * k]),
                            rng.integers(-32768 * k, 32767 * k + 1, size=2000)])
        w = w[(w >= -32768 * k) & (w <= 32767 * k)]
        check(k, w)
    check(2, np.arange(-65536, 65535, dtype=np.int64), scale=2)


def test_int16_head_weights_cover_exactly_the_head():
    """The dp2a byte-weight table the host builds for the lag run (csrc/mavg.cu) selects exactly the first m_part
    run elements, split by channel."""
    R = 32
    for C in (1, 2):
        for k in (2, 3, 5, 16, 17, 31, 100, 255, 1000, 4096):
            Lk = k * C
            s = (R - Lk % R) % R
            m_part = R - s
            mis = 8 * ((Lk + 7) // 8) - Lk
            tab = [[0] * 20 for _ in range(2)]
            for wi in range(20):
                for hh in range(2):
                    r = 2 * wi + hh - mis
                    if 0 <= r < m_part:
                        tab[r % C][wi] |= 1 << (8 * hh)
            picked = sorted((2 * wi + hh - mis, c) for c in range(2) for wi in range(20) for hh in range(2)
                            if tab[c][wi] >> (8 * hh) & 1)
            assert picked == [(r, r % C) for r in range(m_part)]
            assert mis % C == 0 and 0 <= mis < 8


@pytest.mark.parametrize("dist", ["U01", "USYM", "DC1E4"])
def test_far_lag_model_error_budget(oracle_mod, dist):
    """The far-lag kernel's arithmetic (fp64 carried window sum, fp32 differences of 16-sample sums) in NumPy:
    forward error relative to the mean absolute window content stays orders of magnitude inside 1e-5, also for
    zero-mean and large-offset inputs and across chunk boundaries."""
    from algo_model import far_lag_model
    n, L = 11 * 8192 + 123, 20_011
    x = oracle_mod.fill_f32(n, 777, dist=getattr(oracle_mod, "DIST_" + dist))
    y = far_lag_model(x, L, chunk_tiles=3)
    e = oracle_mod.mavg_f64(x, L)
    scale = oracle_mod.mavg_f64(np.abs(x), L)
    err = float(np.max(np.abs(y - e) / np.maximum(scale, 1e-30)))
    assert err < 2e-6, err
