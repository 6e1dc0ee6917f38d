/*
 * mavg_oracle.h -- CPU oracle for the moving-average hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it, and only as the checker or the CPU yardstick.
 * libmavg never links or calls it.
 *
 * What it restates: the definition of the op in the reference's CPU path,
 *   basics/profilable_moving_averager.cpp:14-37  (profilable_cpu_computations)
 * with the input/zero-padding conventions of
 *   wav_header.h:26-48 (interleaved samples) and gpu_utils.h:112-123 (zero halo).
 *
 * Parity pin: the reference ships no tests and no golden vectors (SURVEY.md
 * section 8c), so the int16 restatement is pinned against the reference's own
 * function compiled from /root/reference (oracle/_ref/libref_cpu.so, built by
 * oracle/Makefile) -- see tests/test_oracle.py and tests/golden/make_golden.py.
 * The float path is an extension the reference does not have; it is pinned
 * against exact integer arithmetic (tests/golden/make_golden.py).
 */
#ifndef MAVG_ORACLE_H
#define MAVG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Distributions of the counter-based synthetic generator. */
enum {
    ORACLE_DIST_U01 = 0,      /* D1: U[0,1), 24-bit lattice                  */
    ORACLE_DIST_USYM = 1,     /* D2: U[-1,1), zero mean                      */
    ORACLE_DIST_I16 = 2,      /* D3: integer-valued U[-32768,32767]          */
    ORACLE_DIST_DC1E4 = 3     /* D4: 1e4 DC offset + U[-1,1) noise           */
};

/* a1, bit-exact: int16 in, int64 running sum per channel, C integer division
 * (truncation toward zero), divides by k during warm-up too.
 * basics/profilable_moving_averager.cpp:14-37.  The reference reads out of
 * bounds when frames < k (its warm-up loop is unguarded, :19-21); here the
 * warm-up loop stops at `frames` (pure warm-up), which is the only defined
 * reading. */
void oracle_mavg_i16(const int16_t *x, int16_t *y, uint64_t frames,
                     uint32_t channels, uint32_t k);

/* a1 lifted to reals, evaluated in fp64 for fp32 inputs:
 *   y[i,c] = (1/k) * sum_{j=max(0,i-k+1)}^{i} x[j,c]
 * Window sums are re-started from scratch every 4096 frames, so no long-running
 * accumulator drift enters the oracle. */
void oracle_mavg_f32_to_f64(const float *x, double *y, uint64_t frames,
                            uint32_t channels, uint32_t k);

/* Moving RMS, the windowed reduction SURVEY.md section 8(f) row 4 asks for beside the average (the reference has no
 * such binary; the window, padding and divide-by-k-during-warm-up conventions are those of a1):
 *   y[i,c] = sqrt( (1/k) * sum_{j=max(0,i-k+1)}^{i} x[j,c]^2 )
 * float32 input evaluated in fp64 (fresh window sums every 4096 frames); int16 input with an exact int64 sum of
 * squares and  y = (int16) min(32767, trunc( sqrt( (double) sum / k ) ))  (IEEE double division and square root). */
void oracle_mrms_f32_to_f64(const float *x, double *y, uint64_t frames,
                            uint32_t channels, uint32_t k);
void oracle_mrms_i16(const int16_t *x, int16_t *y, uint64_t frames,
                     uint32_t channels, uint32_t k);

/* Port of the same running-sum loop in fp32 (the apples-to-apples CPU baseline
 * for the fp32 GPU path; NOT an accuracy oracle).  Single thread. */
void oracle_mavg_f32_running(const float *x, float *y, uint64_t frames,
                             uint32_t channels, uint32_t k);

/* Same loop, frames split into contiguous chunks with a k-frame halo, one
 * pthread per chunk.  Returns the number of threads actually used. */
int oracle_mavg_f32_running_mt(const float *x, float *y, uint64_t frames,
                               uint32_t channels, uint32_t k, int threads);
int oracle_mavg_i16_mt(const int16_t *x, int16_t *y, uint64_t frames,
                       uint32_t channels, uint32_t k, int threads);

/* Counter-based generator: value = f(seed, global index).  libmavg implements
 * the same published formula on the device (include/mavg.h, mavg_fill_synthetic). */
uint64_t oracle_mix64(uint64_t seed, uint64_t index);
void oracle_fill_f32(float *dst, uint64_t n, uint64_t first_index, uint64_t seed, int dist);
void oracle_fill_i16(int16_t *dst, uint64_t n, uint64_t first_index, uint64_t seed);

/* Point evaluation of the fp64 definition straight from the generator (used to
 * spot-check outputs of signals too large to hold on the host): mono signal,
 * sample index i, window k. */
double oracle_point_f64(uint64_t i, uint32_t k, uint64_t seed, int dist);

/* Wall-clock helper: runs fn-selected loop `iters` times, returns best seconds.
 * which: 0 = i16 single thread, 1 = f32 running single thread,
 *        2 = f32 running mt, 3 = i16 mt. */
double oracle_time_best(int which, const void *x, void *y, uint64_t frames,
                        uint32_t channels, uint32_t k, int threads, int iters);

#ifdef __cplusplus
}
#endif
#endif
