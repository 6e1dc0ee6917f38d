// ref_shim.cpp -- C-callable door onto the UNMODIFIED reference CPU function.
//
// TEST INFRASTRUCTURE ONLY (see mavg_oracle.h).  This file contains no
// reference code: it only declares the reference's function
//   profilable_cpu_computations(int, int, const vector<int16_t>&, vector<int16_t>&)
//   (/root/reference/basics/profilable_moving_averager.cpp:14)
// and forwards to it.  oracle/Makefile compiles the reference translation unit
// straight from /root/reference (with the one missing '}' after line 83 streamed
// in by sed -- the file does not compile as shipped) and links
// it with this shim into oracle/_ref/libref_cpu.so.
#include <chrono>
#include <cstdint>
#include <vector>

void profilable_cpu_computations(int numberOfChannels, int point,
                                 const std::vector<int16_t>& samples,
                                 std::vector<int16_t>& processedSamples);

extern "C" {

// Runs the reference on `n` interleaved int16 samples.  The reference indexes
// samples[i*C+ch] for i < point without a bound (its warm-up loop), so callers
// must keep n / channels >= k.  Returns 0, or -1 when that precondition fails.
int ref_cpu_i16(int channels, int k, const int16_t* in, int16_t* out, uint64_t n)
{
    if (channels <= 0 || k <= 0 || n / (uint64_t)channels < (uint64_t)k) return -1;
    std::vector<int16_t> src(in, in + n);
    std::vector<int16_t> dst(n);
    profilable_cpu_computations(channels, k, src, dst);
    for (uint64_t i = 0; i < n; ++i) out[i] = dst[i];
    return 0;
}

// Best-of-`iters` wall-clock seconds of the reference function alone (vectors
// are built outside the timed region, as the reference's own harness does,
// profilable_moving_averager.cpp:61-67).
double ref_cpu_i16_time(int channels, int k, const int16_t* in, int16_t* out,
                        uint64_t n, int iters)
{
    if (channels <= 0 || k <= 0 || n / (uint64_t)channels < (uint64_t)k) return -1.0;
    std::vector<int16_t> src(in, in + n);
    std::vector<int16_t> dst(n);
    double best = 1e300;
    for (int it = 0; it < iters; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        profilable_cpu_computations(channels, k, src, dst);
        auto t1 = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        if (s < best) best = s;
    }
    if (out) for (uint64_t i = 0; i < n; ++i) out[i] = dst[i];
    return best;
}

}  // extern "C"
