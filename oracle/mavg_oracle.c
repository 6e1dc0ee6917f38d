/*
 * mavg_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see mavg_oracle.h).
 *
 * Restates basics/profilable_moving_averager.cpp:14-37 of the reference:
 * a causal, zero-padded, per-channel moving average over interleaved frames
 * that divides by the full window k even while the window is still filling.
 */
#define _POSIX_C_SOURCE 200809L
#include "mavg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---------------------------------------------------------------- int16 (a1) */

/* One channel-interleaved pass over frames [f0, f1).  `acc` must already hold,
 * per channel, the sum of the (at most k) frames preceding f0.
 * Reference arithmetic: int64 accumulator, `/` truncating toward zero, result
 * narrowed to int16 (profilable_moving_averager.cpp:22-23, :31-33). */
static void i16_span(const int16_t *x, int16_t *y, uint64_t f0, uint64_t f1,
                     uint32_t ch, uint32_t k, int64_t *acc)
{
    const int64_t div = (int64_t)k;
    for (uint64_t f = f0; f < f1; ++f) {
        const int16_t *in = x + f * ch;
        int16_t *out = y + f * ch;
        if (f >= k) {
            const int16_t *old = x + (f - k) * ch;
            for (uint32_t c = 0; c < ch; ++c) {
                acc[c] += (int64_t)in[c] - (int64_t)old[c];
                out[c] = (int16_t)(acc[c] / div);
            }
        } else {
            for (uint32_t c = 0; c < ch; ++c) {
                acc[c] += in[c];
                out[c] = (int16_t)(acc[c] / div);
            }
        }
    }
}

/* Sum of the frames in [max(0,f0-k), f0) for every channel. */
static void i16_prime(const int16_t *x, uint64_t f0, uint32_t ch, uint32_t k, int64_t *acc)
{
    memset(acc, 0, sizeof(int64_t) * ch);
    uint64_t lo = f0 > k ? f0 - k : 0;
    for (uint64_t f = lo; f < f0; ++f)
        for (uint32_t c = 0; c < ch; ++c) acc[c] += x[f * ch + c];
}

void oracle_mavg_i16(const int16_t *x, int16_t *y, uint64_t frames,
                     uint32_t channels, uint32_t k)
{
    if (!frames || !channels || !k) return;
    int64_t *acc = (int64_t *)calloc(channels, sizeof(int64_t));
    i16_span(x, y, 0, frames, channels, k, acc);
    free(acc);
}

/* ------------------------------------------------------------- fp64 oracle */

void oracle_mavg_f32_to_f64(const float *x, double *y, uint64_t frames,
                            uint32_t channels, uint32_t k)
{
    if (!frames || !channels || !k) return;
    const uint64_t RESTART = 4096;
    const double inv = 1.0 / (double)k; /* used only as a check below */
    (void)inv;
    for (uint32_t c = 0; c < channels; ++c) {
        double w = 0.0;
        for (uint64_t f = 0; f < frames; ++f) {
            if (f % RESTART == 0) {
                /* fresh window sum: frames [max(0,f-k+1), f] */
                uint64_t lo = (f + 1 > k) ? f + 1 - k : 0;
                w = 0.0;
                for (uint64_t j = lo; j <= f; ++j) w += (double)x[j * channels + c];
            } else {
                w += (double)x[f * channels + c];
                if (f >= k) w -= (double)x[(f - k) * channels + c];
            }
            y[f * channels + c] = w / (double)k;
        }
    }
}

/* ------------------------------------------------------------- moving RMS */

void oracle_mrms_f32_to_f64(const float *x, double *y, uint64_t frames,
                            uint32_t channels, uint32_t k)
{
    if (!frames || !channels || !k) return;
    const uint64_t RESTART = 4096;
    for (uint32_t c = 0; c < channels; ++c) {
        double w = 0.0;
        for (uint64_t f = 0; f < frames; ++f) {
            if (f % RESTART == 0) {
                uint64_t lo = (f + 1 > k) ? f + 1 - k : 0;
                w = 0.0;
                for (uint64_t j = lo; j <= f; ++j) {
                    const double v = (double)x[j * channels + c];
                    w += v * v;
                }
            } else {
                const double v = (double)x[f * channels + c];
                w += v * v;
                if (f >= k) {
                    const double o = (double)x[(f - k) * channels + c];
                    w -= o * o;
                }
            }
            y[f * channels + c] = sqrt((w > 0.0 ? w : 0.0) / (double)k);
        }
    }
}

void oracle_mrms_i16(const int16_t *x, int16_t *y, uint64_t frames,
                     uint32_t channels, uint32_t k)
{
    if (!frames || !channels || !k) return;
    int64_t *acc = (int64_t *)calloc(channels, sizeof(int64_t));
    for (uint64_t f = 0; f < frames; ++f)
        for (uint32_t c = 0; c < channels; ++c) {
            const int64_t v = x[f * channels + c];
            acc[c] += v * v;
            if (f >= k) {
                const int64_t o = x[(f - k) * channels + c];
                acc[c] -= o * o;
            }
            const double r = sqrt((double)acc[c] / (double)k);   /* 32768 (all samples at -32768) saturates */
            y[f * channels + c] = (int16_t)(r > 32767.0 ? 32767.0 : r);
        }
    free(acc);
}

/* --------------------------------------------------------- fp32 CPU port */

static void f32_span(const float *x, float *y, uint64_t f0, uint64_t f1,
                     uint32_t ch, uint32_t k, float *acc)
{
    const float div = (float)k;
    for (uint64_t f = f0; f < f1; ++f) {
        const float *in = x + f * ch;
        float *out = y + f * ch;
        if (f >= k) {
            const float *old = x + (f - k) * ch;
            for (uint32_t c = 0; c < ch; ++c) {
                acc[c] -= old[c];
                acc[c] += in[c];
                out[c] = acc[c] / div;
            }
        } else {
            for (uint32_t c = 0; c < ch; ++c) {
                acc[c] += in[c];
                out[c] = acc[c] / div;
            }
        }
    }
}

static void f32_prime(const float *x, uint64_t f0, uint32_t ch, uint32_t k, float *acc)
{
    memset(acc, 0, sizeof(float) * ch);
    uint64_t lo = f0 > k ? f0 - k : 0;
    for (uint64_t f = lo; f < f0; ++f)
        for (uint32_t c = 0; c < ch; ++c) acc[c] += x[f * ch + c];
}

void oracle_mavg_f32_running(const float *x, float *y, uint64_t frames,
                             uint32_t channels, uint32_t k)
{
    if (!frames || !channels || !k) return;
    float *acc = (float *)calloc(channels, sizeof(float));
    f32_span(x, y, 0, frames, channels, k, acc);
    free(acc);
}

/* ------------------------------------------------------------ threaded */

typedef struct {
    int is_i16;
    const void *x;
    void *y;
    uint64_t f0, f1;
    uint32_t ch, k;
} span_job;

static void *span_worker(void *p)
{
    span_job *j = (span_job *)p;
    if (j->is_i16) {
        int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * j->ch);
        i16_prime((const int16_t *)j->x, j->f0, j->ch, j->k, acc);
        i16_span((const int16_t *)j->x, (int16_t *)j->y, j->f0, j->f1, j->ch, j->k, acc);
        free(acc);
    } else {
        float *acc = (float *)malloc(sizeof(float) * j->ch);
        f32_prime((const float *)j->x, j->f0, j->ch, j->k, acc);
        f32_span((const float *)j->x, (float *)j->y, j->f0, j->f1, j->ch, j->k, acc);
        free(acc);
    }
    return NULL;
}

static int run_mt(int is_i16, const void *x, void *y, uint64_t frames,
                  uint32_t ch, uint32_t k, int threads)
{
    if (!frames || !ch || !k) return 0;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > frames) threads = (int)frames;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    span_job *jobs = (span_job *)malloc(sizeof(span_job) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t].is_i16 = is_i16;
        jobs[t].x = x;
        jobs[t].y = y;
        jobs[t].f0 = frames * (uint64_t)t / (uint64_t)threads;
        jobs[t].f1 = frames * (uint64_t)(t + 1) / (uint64_t)threads;
        jobs[t].ch = ch;
        jobs[t].k = k;
        pthread_create(&tid[t], NULL, span_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    free(jobs);
    free(tid);
    return threads;
}

int oracle_mavg_f32_running_mt(const float *x, float *y, uint64_t frames,
                               uint32_t channels, uint32_t k, int threads)
{
    return run_mt(0, x, y, frames, channels, k, threads);
}

int oracle_mavg_i16_mt(const int16_t *x, int16_t *y, uint64_t frames,
                       uint32_t channels, uint32_t k, int threads)
{
    return run_mt(1, x, y, frames, channels, k, threads);
}

/* ------------------------------------------------------------ generator */

/* splitmix64 finaliser over (index + seed * golden-ratio increment). */
uint64_t oracle_mix64(uint64_t seed, uint64_t index)
{
    uint64_t z = index + seed * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static float gen_f32(uint64_t seed, uint64_t index, int dist)
{
    uint64_t z = oracle_mix64(seed, index);
    float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f); /* 24-bit lattice, exact */
    switch (dist) {
    case ORACLE_DIST_U01: return u;
    case ORACLE_DIST_USYM: return 2.0f * u - 1.0f; /* exact */
    case ORACLE_DIST_I16: return (float)((int32_t)(z >> 48) - 32768);
    case ORACLE_DIST_DC1E4: {
        volatile float n = 2.0f * u - 1.0f; /* keep the single rounding below unfused */
        return 10000.0f + n;
    }
    default: return 0.0f;
    }
}

void oracle_fill_f32(float *dst, uint64_t n, uint64_t first_index, uint64_t seed, int dist)
{
    for (uint64_t i = 0; i < n; ++i) dst[i] = gen_f32(seed, first_index + i, dist);
}

void oracle_fill_i16(int16_t *dst, uint64_t n, uint64_t first_index, uint64_t seed)
{
    for (uint64_t i = 0; i < n; ++i)
        dst[i] = (int16_t)((int32_t)(oracle_mix64(seed, first_index + i) >> 48) - 32768);
}

double oracle_point_f64(uint64_t i, uint32_t k, uint64_t seed, int dist)
{
    uint64_t lo = (i + 1 > k) ? i + 1 - k : 0;
    double w = 0.0;
    for (uint64_t j = lo; j <= i; ++j) w += (double)gen_f32(seed, j, dist);
    return w / (double)k;
}

/* --------------------------------------------------------------- timing */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double oracle_time_best(int which, const void *x, void *y, uint64_t frames,
                        uint32_t channels, uint32_t k, int threads, int iters)
{
    double best = 1e300;
    for (int it = 0; it < iters; ++it) {
        double t0 = now_s();
        switch (which) {
        case 0: oracle_mavg_i16((const int16_t *)x, (int16_t *)y, frames, channels, k); break;
        case 1: oracle_mavg_f32_running((const float *)x, (float *)y, frames, channels, k); break;
        case 2: oracle_mavg_f32_running_mt((const float *)x, (float *)y, frames, channels, k, threads); break;
        case 3: oracle_mavg_i16_mt((const int16_t *)x, (int16_t *)y, frames, channels, k, threads); break;
        default: return -1.0;
        }
        double dt = now_s() - t0;
        if (dt < best) best = dt;
    }
    return best;
}
