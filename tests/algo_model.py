"""NumPy model of the arithmetic of libmavg's streaming kernel (stream_f32_kernel).

TEST INFRASTRUCTURE.  It mirrors, operation for operation and in fp32, how
digital_signal_processsing_b200/csrc/mavg_kernels.cuh forms each output:

  * tiles of NT*R samples on a grid anchored at sample 0, zero padded on the left;
  * each thread owns R consecutive outputs; group total in fixed pairwise order;
  * MODE 0 (k <= direct_max): window sum at the run start = n_full preceding group
    totals (nearest first) + the first m_part samples of the lag run;
  * MODE 1: the n_full groups come from a tile-rebased prefix scan of group totals
    (warp-level Hillis-Steele, exclusive warp offsets, tile totals);
  * then w += x[i] - x[i-k], y = w * (1/k).

`far_lag_model` restates stream_far_f32_kernel the same way (fp64 carried window sum, fp32
differences of 16-sample run totals, per-tile scan, fp32 slide).

It lets the CPU-only test-suite check the index algebra (n_full, m_part, lag
misalignment, history tiles) and the 1e-5 error budget for every k without a GPU; the
GPU tests check the CUDA implementation itself.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def geometry(k: int, NT: int = 256, R: int = 16, direct_max: int = 256):
    s = (R - k % R) % R
    m_part = R - s
    n_full = (k + s) // R - 1
    lag_chunks = (k + 3) // 4
    mis = 4 * lag_chunks - k
    T = NT * R
    H = ((n_full + 1) * R + T - 1) // T
    mode = 2 if k <= 8 else (0 if k <= direct_max else 1)
    if mode == 2:
        H, mis = 1, 0
    return dict(s=s, m_part=m_part, n_full=n_full, lag_chunks=lag_chunks, mis=mis, T=T, H=H, mode=mode)


def _pairwise_group_total(X: np.ndarray) -> np.ndarray:
    """X: [NT, R] fp32 -> [NT] with the kernel's association order."""
    R = X.shape[1]
    q = (X[:, 0::4] + X[:, 1::4]) + (X[:, 2::4] + X[:, 3::4])  # [NT, R/4]
    g = (q[:, 0] + q[:, 1]) + (q[:, 2] + q[:, 3])
    if R == 32:
        g = g + ((q[:, 4] + q[:, 5]) + (q[:, 6] + q[:, 7]))
    return g.astype(f32)


def _warp_inclusive(v: np.ndarray, width: int = 32) -> np.ndarray:
    """Hillis-Steele inclusive scan inside each group of 32 lanes (shfl_up order)."""
    v = v.astype(f32).copy().reshape(-1, 32)
    d = 1
    while d < width:
        up = np.zeros_like(v)
        up[:, d:] = v[:, :-d]
        add = v + up
        v[:, d:] = add[:, d:]
        d <<= 1
    return v.reshape(-1)


def small_window_model(x: np.ndarray, k: int) -> np.ndarray:
    """MODE 2 (k <= 8): additions only, power-of-two partial windows (small_window_sums)."""
    n = x.size
    v = np.zeros(n + 16, dtype=f32)
    v[k - 1:k - 1 + n] = x                     # v[j] = sample j-(k-1); window of output r = v[r .. r+k-1]
    w2 = (v[:-1] + v[1:]).astype(f32)
    w4 = (w2[:-2] + w2[2:]).astype(f32)
    w8 = (w4[:-4] + w4[4:]).astype(f32)
    r = np.arange(n)
    if k >= 8:
        acc, off = w8[r], 8
    elif k >= 4:
        acc, off = w4[r], 4
    elif k >= 2:
        acc, off = w2[r], 2
    else:
        acc, off = v[r], 1
    if 4 <= k < 8 and (k & 2):
        acc = (acc + w2[r + off]).astype(f32)
        off += 2
    if k >= 2 and (k & 1):
        acc = (acc + v[r + off]).astype(f32)
    return (acc * (f32(1.0) / f32(k))).astype(f32)


def stream_model(x: np.ndarray, k: int, NT: int = 256, R: int = 16, direct_max: int = 256) -> np.ndarray:
    x = np.asarray(x, dtype=f32)
    n = x.size
    g = geometry(k, NT, R, direct_max)
    if g["mode"] == 2:
        return small_window_model(x, k)
    T, H, NW = g["T"], g["H"], NT // 32
    ntiles = (n + T - 1) // T
    pad_l = H * T
    buf = np.zeros(pad_l + ntiles * T, dtype=f32)
    buf[pad_l:pad_l + n] = x
    y = np.zeros(ntiles * T, dtype=f32)
    inv = f32(1.0) / f32(k)

    # per-tile summaries for every tile index in [-H, ntiles)
    gt = {}    # group totals [NT]
    wi = {}    # warp-inclusive prefixes [NT]
    wex = {}   # exclusive warp offsets [NW]
    tt = {}    # tile total
    for t in range(-H, ntiles):
        X = buf[pad_l + t * T: pad_l + (t + 1) * T].reshape(NT, R)
        gt[t] = _pairwise_group_total(X)
        if g["mode"] == 1:
            wi[t] = _warp_inclusive(gt[t])
            raw = wi[t].reshape(NW, 32)[:, 31].copy()
            lanes = np.zeros(32, dtype=f32)
            lanes[:NW] = raw
            winc = _warp_inclusive(lanes, width=NW)[:NW]
            wex[t] = (winc - raw).astype(f32)
            tt[t] = winc[NW - 1]

    tid = np.arange(NT)
    for t in range(ntiles):
        base = pad_l + t * T
        X = buf[base: base + T].reshape(NT, R)
        a = base + tid * R                      # absolute buffer index of each run start
        lag0 = a - k                            # first lag sample of each run
        acc = np.zeros(NT, dtype=f32)
        if g["mode"] == 0:
            for j in range(1, g["n_full"] + 1):
                gidx = tid - j                  # group index relative to this tile
                tt_off = np.floor_divide(gidx, NT)
                vals = np.array([gt[t + int(o)][int(i) % NT] for o, i in zip(tt_off, gidx)], dtype=f32)
                acc = (acc + vals).astype(f32)
        else:
            D = g["n_full"] + 1
            lt = tid - D
            h = np.where(lt < 0, (-lt + NT - 1) // NT, 0)
            lt = lt + h * NT
            e_own = (wex[t][tid >> 5] + (wi[t] - gt[t]).astype(f32)).astype(f32)
            for i in range(NT):
                hi = int(h[i])
                src = t - hi
                cp_lag = f32(wi[src][lt[i]] + wex[src][lt[i] >> 5])
                if hi == 0:
                    acc[i] = f32(e_own[i] - cp_lag)
                else:
                    rest = f32(tt[src] - cp_lag)
                    for v in range(1, hi):
                        rest = f32(rest + tt[src + v])
                    acc[i] = f32(e_own[i] + rest)
        for r in range(R):
            if r < g["m_part"]:
                acc = (acc + buf[lag0 + r]).astype(f32)
        w = acc
        for r in range(R):
            w = (w + (X[:, r] - buf[lag0 + r]).astype(f32)).astype(f32)
            y[t * T + tid * R + r] = w * inv
    return y[:n]


def far_lag_model(x: np.ndarray, L: int, chunk_tiles: int = 4, NT: int = 384) -> np.ndarray:
    """stream_far_f32_kernel (mono), operation for operation in fp32 / fp64: per chunk of `chunk_tiles` tiles the
    window sum W in front of the tile is built by ceil(L / T) masked warm-up tiles and carried in fp64; inside a
    tile every thread forms d = (own run total) - (lag run total) pairwise in fp32, the tile scans d (warp-level
    Hillis-Steele, then the NT / 32 warp totals), a run starts from float(W + exclusive prefix) and slides in fp32."""
    R, NW = 16, NT // 32            # tiles of NT x 16 samples: 384 threads by default, 512 with tuning.threads = 512
    T = NT * R
    x = np.asarray(x, dtype=f32)
    n = x.size
    ntiles = (n + T - 1) // T
    HT = (L + T - 1) // T
    pad = (HT + 1) * T + L
    buf = np.zeros(pad + ntiles * T, dtype=f32)
    buf[pad:pad + n] = x
    y = np.zeros(ntiles * T, dtype=f32)
    inv = f32(1.0) / f32(L)
    tid = np.arange(NT)

    def scan_tile(d):
        incl = _warp_inclusive(d)
        raw = incl.reshape(NW, 32)[:, 31].copy()
        lanes = np.zeros(32, dtype=f32)
        lanes[:NW] = raw
        winc = _warp_inclusive(lanes, width=NW)[:NW]
        own_off = (winc - raw).astype(f32)[tid >> 5]
        return incl, own_off, winc[NW - 1]

    for t0 in range(0, ntiles, chunk_tiles):
        W = np.float64(0.0)
        for j in range(HT):                                   # warm-up tiles in front of the chunk
            u = t0 - HT + j
            X = buf[pad + u * T: pad + (u + 1) * T].reshape(NT, R).copy()
            if j == 0:
                m0 = HT * T - L
                idx = (tid[:, None] * R + np.arange(R)[None, :])
                X[idx < m0] = 0
            d = np.zeros(NT, dtype=f32)
            for r in range(R):
                d = (d + X[:, r]).astype(f32)
            _, _, dtot = scan_tile(d)
            W = W + np.float64(dtot)
        for t in range(t0, min(t0 + chunk_tiles, ntiles)):
            base = pad + t * T
            X = buf[base: base + T].reshape(NT, R)
            XL = buf[base - L: base - L + T].reshape(NT, R)
            d = (_pairwise_group_total(X) - _pairwise_group_total(XL)).astype(f32)
            incl, own_off, dtot = scan_tile(d)
            e_run = (own_off + (incl - d).astype(f32)).astype(f32)
            acc = (W + e_run.astype(np.float64)).astype(f32)
            for r in range(R):
                acc = (acc + (X[:, r] - XL[:, r]).astype(f32)).astype(f32)
                y[t * T + tid * R + r] = acc * inv
            W = W + np.float64(dtot)
    return y[:n]
