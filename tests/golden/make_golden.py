#!/usr/bin/env python
"""Generate tests/golden/*.npz -- run HERE (the container that has /root/reference).

The reference ships no tests and no golden vectors, so the
pins are minted from the reference itself:

* golden_i16.npz  -- outputs of the UNMODIFIED reference function
  profilable_cpu_computations (basics/profilable_moving_averager.cpp:14-37),
  compiled from /root/reference by oracle/Makefile (`make ref`) and called through
  oracle/_ref/libref_cpu.so.  Inputs are stored next to the outputs, so the
  fixtures are self-contained on the GPU box (where /root/reference is absent).
* golden_f32.npz  -- the float extension has no reference implementation; its pin is
  the a1 definition evaluated in EXACT integer arithmetic: the inputs live on a
  2^-24 lattice, so window sums are exact int64 and the single fp64 division is
  correctly rounded.

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import oracle  # noqa: E402

I16_CASES = [
    # (name, frames, channels, k, seed)
    ("kat_mono", 6, 1, 3, None),
    ("k1_identity", 257, 1, 1, 11),
    ("mono_k2", 1000, 1, 2, 12),
    ("mono_k5", 4096, 1, 5, 13),
    ("mono_k41", 4099, 1, 41, 14),      # k=41: the grade where the reference GPU variants drift by 1 LSB
    ("mono_k850", 4000, 1, 850, 15),
    ("mono_k_eq_frames", 300, 1, 300, 16),
    ("stereo_k5", 2048, 2, 5, 17),
    ("stereo_k64", 3001, 2, 64, 18),
    ("stereo_k1000", 2500, 2, 1000, 19),
    ("three_ch_k7", 999, 3, 7, 20),
    ("eight_ch_k33", 512, 8, 33, 21),
]

F32_CASES = [
    # (name, frames, channels, k, seed, dist)
    ("u01_k3", 4096, 1, 3, 31, oracle.DIST_U01),
    ("u01_k5", 4096, 1, 5, 32, oracle.DIST_U01),
    ("u01_k16", 5000, 1, 16, 33, oracle.DIST_U01),
    ("u01_k64", 5000, 1, 64, 34, oracle.DIST_U01),
    ("u01_k256", 9000, 1, 256, 35, oracle.DIST_U01),
    ("u01_k1024", 9000, 1, 1024, 36, oracle.DIST_U01),
    ("u01_k4096", 12000, 1, 4096, 37, oracle.DIST_U01),
    ("usym_k100", 6000, 1, 100, 38, oracle.DIST_USYM),
    ("i16val_k41", 6000, 1, 41, 39, oracle.DIST_I16),
    ("u01_stereo_k9", 3000, 2, 9, 40, oracle.DIST_U01),
    ("u01_k_gt_frames", 100, 1, 250, 41, oracle.DIST_U01),
]


def exact_f64(x: np.ndarray, k: int, channels: int) -> np.ndarray:
    """a1 lifted to reals, exact: x is on the 2^-24 lattice (or integer valued)."""
    scaled = np.round(x.astype(np.float64) * 2.0**24).astype(np.int64)
    assert np.array_equal(scaled.astype(np.float64) / 2.0**24, x.astype(np.float64)), "input not on the lattice"
    frames = x.size // channels
    s = scaled.reshape(frames, channels)
    c = np.cumsum(s, axis=0, dtype=np.int64)
    w = c.copy()
    if frames > k:
        w[k:] -= c[:-k]
    assert np.abs(w).max() < 2**53
    return (w.astype(np.float64) / (float(k) * 2.0**24)).reshape(-1)


def main() -> None:
    oracle.build(ref=True)
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libref_cpu.so missing: /root/reference not present?")

    out = {}
    for name, frames, ch, k, seed in I16_CASES:
        if seed is None:
            x = np.arange(1, frames * ch + 1, dtype=np.int16)
        else:
            x = oracle.fill_i16(frames * ch, seed)
        y = oracle.ref_mavg_i16(x, k, ch)          # the reference itself
        out[f"{name}__x"] = x
        out[f"{name}__y"] = y
        out[f"{name}__meta"] = np.array([frames, ch, k], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_i16.npz"), **out)

    out = {}
    for name, frames, ch, k, seed, dist in F32_CASES:
        x = oracle.fill_f32(frames * ch, seed, dist)
        out[f"{name}__x"] = x
        out[f"{name}__y"] = exact_f64(x, k, ch)
        out[f"{name}__meta"] = np.array([frames, ch, k], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "golden_f32.npz"), **out)
    print("wrote golden_i16.npz (%d cases, from the reference) and golden_f32.npz (%d cases, exact arithmetic)"
          % (len(I16_CASES), len(F32_CASES)))


if __name__ == "__main__":
    main()
