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


def test_int16_biased_division_is_exact():
    """Host restatement of the int16 kernel's division (plan_stream_i16 / div_biased in csrc): with every sample
    biased by +32768 the window sum is w' = w + 32768 k, and
        trunc(w / k) + 32768 == umulhi(w' + (w' < 32768 k ? k - 1 : 0), M) >> (L - 1)
    with L = ceil(log2 k), M = ceil(2^(31+L) / k), for every k in 2..32768 and every reachable w'."""
    rng = np.random.default_rng(5)
    ks = list(range(2, 300)) + [511, 512, 513, 1000, 1024, 4095, 4096, 4097, 16384, 30000, 32767, 32768]
    for k in ks:
        L = (k - 1).bit_length()
        M = ((1 << (31 + L)) + k - 1) // k
        assert M < 2**32 and L >= 1
        B = 32768 * k
        wb = np.concatenate([np.array([0, 1, k - 1, k, B - 1, B, B + 1, B - k, B + k, 65535 * k, 65535 * k - 1], dtype=np.int64),
                             rng.integers(0, 65535 * k + 1, size=400)])
        w = wb - B
        want = np.sign(w) * (np.abs(w) // k) + 32768          # truncation toward zero, biased
        u = wb + np.where(wb < B, k - 1, 0)
        assert u.max() < 2**32
        got = np.array([((int(v) * M) >> 32) >> (L - 1) for v in u], dtype=np.int64)
        assert np.array_equal(got, want), k
        assert got.min() >= 0 and got.max() <= 65535


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
