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
