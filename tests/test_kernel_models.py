"""CPU restatement of the index algebra of the few-channel and far-lag kernels (csrc/mavg_kernels.cuh) and of
the packed-word dp2a arithmetic of the int16 kernels.  TEST INFRASTRUCTURE: integer NumPy models checked against the
oracle, so that the run/tile/prefix bookkeeping and the byte-weight constants are verified without a GPU; the GPU
tests check the CUDA code itself."""
import numpy as np
import pytest


def fewc_geometry(k, C, R, threads=512):
    """plan_fewc (csrc/mavg.cu): run length R frames, NR runs per tile, n_full whole runs + m_part head frames."""
    s = (R - k % R) % R
    m_part = R - s
    n_full = (k + s) // R - 1
    NR = threads // C
    while NR > 0 and (NR * C) % 16:
        NR -= 1
    H = (n_full + 1 + NR - 1) // NR
    return dict(m_part=m_part, n_full=n_full, NR=NR, H=H)


def fewc_model(x, k, R, NR, long_mode):
    """x: [F, C] integers.  Mirrors the kernel: per tile (NR runs of R frames) run totals, optional in-place prefix,
    window start = whole runs (one by one, or prefix differences reaching back over tiles) + head of the lag run,
    then the slide.  Returns exact window sums [F, C] (division is tested separately)."""
    F, C = x.shape
    g = fewc_geometry(k, C, R)
    n_full, m_part = g["n_full"], g["m_part"]
    T = NR * R
    tiles = (F + T - 1) // T
    xp = np.zeros((tiles * T, C), dtype=np.int64)
    xp[:F] = x
    get = lambda i: xp[i] if i >= 0 else np.zeros(C, dtype=np.int64)      # TMA zero fill = left padding
    tot = xp.reshape(tiles, NR, R, C).sum(axis=2)                         # [tile][run][c]
    pre = np.cumsum(tot, axis=1)                                          # fewc_scan_slot, per tile
    out = np.zeros_like(xp)
    zero = np.zeros(C, dtype=np.int64)
    P = lambda t, r: pre[t, r] if t >= 0 else zero                        # history before the signal is all zero
    for t in range(tiles):
        for run in range(NR):
            own = t * T + run * R
            if long_mode:                                                  # fewc_window_long
                q0 = run - n_full
                if q0 >= 0:
                    acc = P(t, run - 1) - (P(t, q0 - 1) if q0 > 0 else 0)
                else:
                    acc = (P(t, run - 1) if run > 0 else zero).copy()
                    m, tt = -q0, t
                    while m > NR:
                        tt -= 1
                        acc = acc + P(tt, NR - 1)
                        m -= NR
                    tt -= 1
                    acc = acc + P(tt, NR - 1) - (P(tt, NR - 1 - m) if m < NR else 0)
            else:                                                          # one by one: this tile, then the previous one
                acc = zero.copy()
                for w in range(1, n_full + 1):
                    r2, t2 = run - w, t
                    if r2 < 0:
                        r2, t2 = r2 + NR, t - 1
                    assert r2 >= 0, "direct mode looks at most one tile back"
                    if t2 >= 0:
                        acc = acc + tot[t2, r2]
            for r in range(R):
                if r < m_part:
                    acc = acc + get(own - k + r)
            for r in range(R):
                acc = acc + xp[own + r] - get(own - k + r)
                out[own + r] = acc
    return out[:F]


def exact_sums(x, k):
    c = np.cumsum(np.vstack([np.zeros((1, x.shape[1]), dtype=np.int64), x.astype(np.int64)]), axis=0)
    hi = np.arange(1, x.shape[0] + 1)
    lo = np.maximum(hi - k, 0)
    return c[hi] - c[lo]


@pytest.mark.parametrize("C,R,NR", [(3, 4, 5), (6, 4, 3), (5, 8, 4)])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 40, 64, 100, 161])
def test_window_bookkeeping_matches_exact_sums(C, R, NR, k):
    """Small run / tile sizes so that windows span many runs and several tiles."""
    rng = np.random.default_rng(k * 31 + C)
    F = 7 * NR * R + 5
    x = rng.integers(-32768, 32768, size=(F, C))
    g = fewc_geometry(k, C, R)
    want = exact_sums(x, k)
    if g["n_full"] >= 1:                                                  # plan_fewc never picks the prefix mode otherwise
        assert np.array_equal(fewc_model(x, k, R, NR, long_mode=True), want)
    if g["n_full"] <= NR:                                                 # direct mode: H == 1 in the plan
        assert np.array_equal(fewc_model(x, k, R, NR, long_mode=False), want)


def test_plan_geometry_identities():
    for C in range(3, 32):
        for R in (16, 32):
            for k in (2, 9, 16, 17, 255, 256, 257, 1000, 4096):
                g = fewc_geometry(k, C, R)
                if g["NR"] == 0:
                    continue
                assert g["n_full"] * R + g["m_part"] == k and 1 <= g["m_part"] <= R
                assert (g["NR"] * C) % 16 == 0 and g["NR"] * C <= 512
                assert g["H"] * g["NR"] >= g["n_full"] + 1                # the lag run is inside the history tiles


# ---------------------------------------------------------------- packed-word arithmetic of the int16 kernels
def dp2a_lo_s32(a, b, c):
    """PTX dp2a.lo.s32.s32: c + a.lo16 * b.byte0 + a.hi16 * b.byte1, halves and bytes signed."""
    a = np.asarray(a, dtype=np.uint32)
    lo = (a & 0xFFFF).astype(np.int64)
    hi = (a >> 16).astype(np.int64)
    lo = np.where(lo >= 32768, lo - 65536, lo)
    hi = np.where(hi >= 32768, hi - 65536, hi)
    sb = lambda v: v - 256 if v >= 128 else v
    return c + lo * sb(b & 0xFF) + hi * sb((b >> 8) & 0xFF)


@pytest.mark.parametrize("wscale", [1, 2])
def test_dp2a_weights_select_add_and_subtract_halves(wscale):
    rng = np.random.default_rng(3)
    s = rng.integers(-32768, 32768, size=(1000, 2))
    words = ((s[:, 1].astype(np.int64) & 0xFFFF) << 16 | (s[:, 0].astype(np.int64) & 0xFFFF)).astype(np.uint32)
    w_lo, w_hi = wscale, wscale << 8
    n_lo = (-wscale) & 0xFF
    n_hi = n_lo << 8
    assert np.array_equal(dp2a_lo_s32(words, w_lo, 0), wscale * s[:, 0])
    assert np.array_equal(dp2a_lo_s32(words, w_hi, 0), wscale * s[:, 1])
    assert np.array_equal(dp2a_lo_s32(words, n_lo, 7), 7 - wscale * s[:, 0])
    assert np.array_equal(dp2a_lo_s32(words, n_hi, 7), 7 - wscale * s[:, 1])
    assert np.array_equal(dp2a_lo_s32(words, w_lo | w_hi, 0), wscale * (s[:, 0] + s[:, 1]))     # mono run totals


def test_pre_swizzled_offsets_equal_swizzled_addresses():
    """swz(base + x) == base + pre_swz(x) for 1024-byte aligned bases and offsets of either sign (two's complement),
    and the ring wrap decision is the same for x and pre_swz(x)."""
    swz = lambda a: a ^ ((a >> 3) & 0x70)
    rng = np.random.default_rng(4)
    ring, tb, S = 3 * 1024, 30 * 1024, 5
    for _ in range(20000):
        st = int(rng.integers(0, S))
        x = int(rng.integers(-3 * tb, tb))
        if st * tb + x < -(S - 1) * tb:
            continue
        o = st * tb + x
        want = swz(ring + (o + S * tb if o < 0 else o))
        xs = x ^ ((x >> 3) & 0x70)                                         # Python ints: arithmetic shift, like the kernel
        b0 = ring + st * tb
        got = xs + (b0 + S * tb if xs < -(st * tb) else b0)
        assert got == want


# ---------------------------------------------------------------- far-lag kernel (stream_far_f32_kernel) geometry
@pytest.mark.parametrize("NT", [384, 512])
def test_far_lag_boxes_and_warm_up_cover_exactly_the_window(NT):
    """Index algebra of the far-lag kernel with its real constants (tiles of NT x 16 samples: 384 threads = 192 rows of
    32 with ONE 193-row lag box per tile, 512 threads = 256 rows with two 129-row boxes; koff / lag_rows / MIS from
    launch_far): every thread's lag run is x[i - L .. i - L + 16) for its own run start i, and the masked warm-up tiles
    in front of a chunk sum exactly the L samples in front of it.  Integer data, so equality is exact."""
    ROW, R = 32, 16
    T = NT * R
    ROWS = T // ROW
    NBOX = 2 if ROWS + 1 > 256 else 1                                   # far_lag_boxes()
    BOXROWS = ROWS // NBOX + 1
    TPB = NT // NBOX
    rng = np.random.default_rng(9)
    for L in (T, T + 1, 9001, 12_346, 16_384, 20_007, 40_000):         # lag distance in flat samples (k, or 2k stereo)
        n = 7 * T
        x = rng.integers(-1000, 1000, size=n).astype(np.int64)
        koff = (ROW - L % ROW) % ROW
        lag_rows = (L + koff) // ROW
        mis = (4 - L % 4) % 4
        assert koff % 4 == mis and (L + koff) % ROW == 0
        HT = (L + T - 1) // T
        rows = n // ROW
        xr = x.reshape(rows, ROW)

        def box(r0, nrows):                                             # TMA box with zero fill outside the tensor
            out = np.zeros((nrows, ROW), dtype=np.int64)
            lo, hi = max(r0, 0), min(r0 + nrows, rows)
            if lo < hi:
                out[lo - r0:hi - r0] = xr[lo:hi]
            return out.reshape(-1)

        for j in (0, 1, 3, 6):
            r0 = j * (T // ROW) - lag_rows
            halves = [box(r0 + b * (ROWS // 2), BOXROWS) for b in range(NBOX)]
            for t in (0, 1, TPB - 1, TPB % NT, 300, NT - 1):
                c0 = koff // 4 + 4 * (t % TPB)                          # first 16-byte chunk the thread loads
                chunks = halves[t // TPB][4 * c0: 4 * (c0 + 5)]
                lag_run = chunks[mis: mis + R]
                i = j * T + R * t
                want = np.array([x[i - L + r] if i - L + r >= 0 else 0 for r in range(R)])
                assert np.array_equal(lag_run, want), (L, j, t)
                assert 4 * (c0 + 4 + (1 if mis else 0)) <= BOXROWS * ROW   # the chunks stay inside the lag box
        # warm-up in front of a chunk that starts at tile t0: HT tiles, the first one masked from m0 on
        for t0 in (0, 2, 5):
            W = 0
            for jj in range(HT):
                u = t0 - HT + jj
                tile = box(u * (T // ROW), T // ROW)
                m0 = HT * T - L if jj == 0 else 0
                W += int(tile[m0:].sum())
            lo = t0 * T - L
            assert W == int(x[max(lo, 0): t0 * T].sum()), (L, t0)


# ---------------------------------------------------------------- flat-stream int16 kernel (stream_i16_kernel), any channel count
FLAT_SHAPES = [(512, 32, 1), (512, 32, 2), (128, 72, 3), (224, 72, 3), (512, 24, 3), (256, 64, 4), (384, 40, 5), (192, 40, 5), (128, 72, 6),
               (224, 72, 6), (256, 56, 7), (128, 56, 7), (128, 64, 8), (256, 64, 8), (224, 72, 12), (256, 64, 16), (384, 32, 2), (192, 64, 8)]


@pytest.mark.parametrize("NT,R,C", FLAT_SHAPES)
def test_flat_i16_delta_scan_matches_exact_window_sums(NT, R, C):
    """stream_i16_kernel's arithmetic restated on integers: a run of R flat samples holds whole frames (channel of
    element r = r % C); s[r] = running sum of x[i] - x[i - L] per channel inside the run, d = its last value per
    channel, start(t) = W + exclusive scan of d over the tile, W carried from tile to tile and built by masked warm-up
    tiles in front of a tile range.  The result has to equal the exact causal window sums for every lag distance,
    including lags that are not whole runs or whole tiles, and ranges that start anywhere."""
    assert R % C == 0 and R % 8 == 0
    T = NT * R
    rng = np.random.default_rng(100 + NT + R + C)
    ntiles = 5
    n = ntiles * T
    x = rng.integers(-32768, 32768, size=n).astype(np.int64)
    xp = np.concatenate([np.zeros(4 * T, dtype=np.int64), x])          # zero padding in front (TMA out-of-bounds fill)
    base = 4 * T
    for k in (2, 3, R // C, R // C + 1, T // C - 1, T // C, T // C + 5, 2 * T // C + 7):
        L = k * C
        H = (L + R - 1) // R * R                                        # (n_full + 1) * R of plan_stream_i16
        H = (H + T - 1) // T
        exact = np.zeros(n, dtype=np.int64)                             # exact window sums per flat sample
        for c in range(C):
            col = xp[c::C].cumsum()
            colp = np.concatenate([np.zeros(k, dtype=np.int64), col])
            win = colp[k:] - colp[:-k]
            exact_c = win[(base + c) // C if (base % C) == 0 else 0:]
            exact[c::C] = exact_c[:len(exact[c::C])]
        for t0 in (0, 2):                                               # a tile range that starts at tile t0
            W = np.zeros(C, dtype=np.int64)
            for j in range(H):                                          # warm-up tiles, the first one masked
                u = t0 - H + j
                tile = xp[base + u * T: base + (u + 1) * T]
                idx = np.arange(T)
                rel = (t0 - u) * T - L                                  # samples at or behind (range start - L) count
                m = np.where(idx >= rel, tile, 0)
                for c in range(C):
                    W[c] += m[c::C].sum()
            for t in range(t0, ntiles):
                own = xp[base + t * T: base + (t + 1) * T].reshape(NT, R)
                lag = xp[base + t * T - L: base + (t + 1) * T - L].reshape(NT, R)
                delta = own - lag
                s_run = np.zeros((NT, R), dtype=np.int64)
                for c in range(C):
                    s_run[:, c::C] = delta[:, c::C].cumsum(axis=1)
                d = np.stack([s_run[:, R - C + c] for c in range(C)], axis=1)     # last value per channel = run delta
                excl = np.cumsum(d, axis=0) - d
                for c in range(C):
                    y = W[c] + excl[:, c][:, None] + s_run[:, c::C]
                    got = y.reshape(-1)
                    want = exact[t * T: (t + 1) * T].reshape(NT, R)[:, c::C].reshape(-1)
                    assert np.array_equal(got, want), (NT, R, C, k, t0, t, c)
                W += d.sum(axis=0)


@pytest.mark.parametrize("NT,R,C", FLAT_SHAPES)
def test_flat_i16_cross_warp_layout_and_lag_alignment(NT, R, C):
    """(i) The [channel][warp] array of the cross-warp scan: lane = (warp w' = lane % NW, group g = lane // NW) reads
    word j * 32 + lane, which has to be the total that warp w' wrote for channel g + j * G -- every (warp, channel) pair
    exactly once, inside the 32 * CJ words the kernel reserves.  (ii) plan_stream_i16's lag addressing: lag_chunks
    16-byte chunks back from the own run plus MIS samples is exactly L samples back, for every window."""
    NW = NT // 32
    G = 32 // NW
    CJ = (C + G - 1) // G
    seen = {}
    for warp in range(NW):
        for c in range(C):
            w = (c // G) * 32 + (c % G) * NW + warp                    # where lane 31 of `warp` stores channel c
            assert 0 <= w < 32 * CJ and w not in seen
            seen[w] = (warp, c)
    for lane in range(G * NW):
        wq, g = lane % NW, lane // NW
        for j in range(CJ):
            c = g + j * G
            if c < C:
                assert seen[j * 32 + lane] == (wq, c)
    for c in range(C):                                                  # the lane a thread of `warp` fetches channel c from
        for warp in range(NW):
            src = (c % G) * NW + warp
            assert src < 32 and src % NW == warp and src // NW == c % G
    for k in range(2, 200):
        L = k * C
        lag_chunks = (L + 7) // 8
        mis = 8 * lag_chunks - L
        assert 0 <= mis < 8 and lag_chunks * 8 - mis == L
        if C % 8 == 0:
            assert mis == 0
        elif C % 4 == 0:
            assert mis in (0, 4)
        elif C % 2 == 0:
            assert mis % 2 == 0


# ---------------------------------------------------------------- far-lag int16 kernel (stream_far_i16_kernel) geometry
@pytest.mark.parametrize("NT,R,C", [(384, 32, 1), (384, 32, 2), (192, 64, 4), (224, 72, 6), (192, 64, 8), (224, 72, 12), (192, 64, 16)])
def test_far_lag_i16_boxes_lag_runs_and_exact_range(NT, R, C):
    """Index algebra of launch_far / stream_far_i16_kernel in int16 samples and 64-sample rows: the lag box of a tile
    starts lag_rows rows in front of it, a thread's lag run starts koff / 8 + (R / 8) * (t % TPB) chunks into its box plus
    MIS samples, which has to be exactly x[i - L .. i - L + R); the warm-up tiles cover exactly the L samples in front
    of a tile range; and the int32 carry / multiply-high division are exact up to k = 46 340 (plan_far_i16)."""
    ROW = 64
    T = NT * R
    ROWS = T // ROW
    NBOX = 2 if ROWS + 1 > 256 else 1
    BOXROWS = ROWS // NBOX + 1
    TPB = NT // NBOX
    CH_OWN = R // 8
    rng = np.random.default_rng(17 + NT + C)
    n = 6 * T
    x = rng.integers(-32768, 32768, size=n).astype(np.int64)
    xr = x.reshape(n // ROW, ROW)

    def box(r0, nrows):
        out = np.zeros((nrows, ROW), dtype=np.int64)
        lo, hi = max(r0, 0), min(r0 + nrows, n // ROW)
        if lo < hi:
            out[lo - r0:hi - r0] = xr[lo:hi]
        return out.reshape(-1)

    for k in (T // C + 1, T // C + 3, (2 * T + 5 * C) // C, 3 * T // C + 7):
        L = k * C
        koff = (ROW - L % ROW) % ROW
        lag_rows = (L + koff) // ROW
        lag_chunks = (L + 7) // 8
        mis = 8 * lag_chunks - L
        assert koff % 8 == mis and (L + koff) % ROW == 0
        for j in (0, 2, 5):
            r0 = j * ROWS - lag_rows
            boxes = [box(r0 + b * (ROWS // 2), BOXROWS) for b in range(NBOX)]
            for t in (0, 1, TPB - 1, TPB % NT, NT // 3, NT - 1):
                c0 = koff // 8 + CH_OWN * (t % TPB)
                chunks = boxes[t // TPB][8 * c0: 8 * (c0 + CH_OWN + 1)]
                lag_run = chunks[mis: mis + R]
                i = j * T + R * t
                want = np.array([x[i - L + r] if i - L + r >= 0 else 0 for r in range(R)])
                assert np.array_equal(lag_run, want), (k, j, t)
                assert 8 * (c0 + CH_OWN + (1 if mis else 0)) <= BOXROWS * ROW
        HT = (L + T - 1) // T
        for t0 in (1, 4):
            W = np.zeros(C, dtype=np.int64)
            for jj in range(HT):
                u = t0 - HT + jj
                tile = box(u * ROWS, ROWS)
                rel = (t0 - u) * T - L
                m = np.where(np.arange(T) >= rel, tile, 0)
                for c in range(C):
                    W[c] += m[c::C].sum()
            lo = t0 * T - L
            seg = x[max(lo, 0): t0 * T]
            off = (max(lo, 0)) % C
            for c in range(C):
                assert W[c] == int(seg[(c - off) % C::C].sum()), (k, t0, c)
    # exact range of the int32 carry and of the multiply-high division (i16_mulhi_consts): every k with k * k < 2^31
    for k in (32_769, 40_000, 46_340):
        assert k * k < 2 ** 31 and 32768 * k < 2 ** 31
        lg = (k - 1).bit_length()
        M = (1 << (30 + lg)) // k + 1
        assert M < 2 ** 31
        e = M * k - (1 << (30 + lg))
        assert 0 < e <= k and 32768 * k * e < (1 << (30 + lg))
        for w in (0, 1, k - 1, k, k + 1, 32768 * k - 1, 32768 * k, 12345 * k + 7, -1, -k, -k - 1, -32768 * k, -(32767 * k + 5)):
            t = (w * M) >> 32                                  # mulhi_s32 (arithmetic shift of the 64-bit product)
            y = (t >> (lg - 2)) + (1 if t < 0 else 0)
            want = abs(w) // k * (1 if w >= 0 else -1)         # C's truncating division
            assert y == want, (k, w)
