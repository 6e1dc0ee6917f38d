// mavg_bench.h -- timing harness of the drop-in binaries.
//
// Interface twin of the reference's benchmark.h (ProfileResult :9-70, GpuTimer :72-96, CpuTimer
// :98-114, benchmark<> :116-132) with the change the north star asks for: the GPU phases are no
// longer four events on the legacy default stream owned by the harness, they are CUDA-event times
// measured inside libmavg on the plan's streams and reported as the MAXIMUM over devices
// (mavg_get_timing).  GpuTimer therefore carries no CUDA state; the pipeline lambda hands it the
// plan's timing after each run.
#pragma once

#include <chrono>
#include <cstdio>

#include "mavg.h"

struct ProfileResult {
    float initialization_ms = 0.0f;
    float transfer_h2d_ms = 0.0f;
    float compute_ms = 0.0f;
    float transfer_d2h_ms = 0.0f;
    float total_ms = 0.0f;

    void operator+=(const ProfileResult& o)
    {
        initialization_ms += o.initialization_ms; transfer_h2d_ms += o.transfer_h2d_ms;
        compute_ms += o.compute_ms; transfer_d2h_ms += o.transfer_d2h_ms; total_ms += o.total_ms;
    }
    void divide(int n)
    {
        if (n <= 0) return;
        const float s = 1.0f / (float)n;
        initialization_ms *= s; transfer_h2d_ms *= s; compute_ms *= s; transfer_d2h_ms *= s; total_ms *= s;
    }
    // Same report sections as the reference (benchmark.h:33-69), plus Gsamples/s and HBM GB/s.
    void print_stats(size_t n, size_t in_bytes, size_t out_bytes = 0) const
    {
        if (!out_bytes) out_bytes = in_bytes;
        const double gb = (double)n * (double)(in_bytes + out_bytes) / 1e9, ms_n = (double)n / 1e6;
        printf("1. LATENCY BREAKDOWN (Steady State)\n");
        if (transfer_h2d_ms > 0) printf("   H2D Transfer:   %.3f ms\n", transfer_h2d_ms);
        printf("   Kernel Compute: %.3f ms\n", compute_ms);
        if (transfer_d2h_ms > 0) printf("   D2H Transfer:   %.3f ms\n", transfer_d2h_ms);
        printf("   -----------------------------\n   TOTAL LATENCY:  %.3f ms\n", total_ms);
        printf("\n2. THROUGHPUT (Steady State)\n");
        if (compute_ms > 0) {
            printf("   Kernel Bandwidth: %.3f GB/s\n", gb / (compute_ms / 1e3));
            printf("   Kernel Speed:   %.3f Mega Samples/s\n", ms_n / (compute_ms / 1e3));
        }
        if (total_ms > 0) {
            printf("   App BandWidth:   %.3f GB/s\n", gb / (total_ms / 1e3));
            printf("   App Speed:      %.3f Mega Samples/s\n", ms_n / (total_ms / 1e3));
            printf("   Cold Start:     %.3f Mega Samples/s (Includes Init)\n", ms_n / ((initialization_ms + total_ms) / 1e3));
        }
        printf("\n3. INITIALIZATION COST (One-time)\n   Allocation:     %.3f ms\n   First Frame:    %.3f ms (Cold Start)\n",
               initialization_ms, initialization_ms + total_ms);
        printf("___________________________________\n\n");
    }
};

// Device-side phase times of the last libmavg run (max over devices).
class GpuTimer {
    mavg_timing t_{0, 0, 0, 0};
public:
    void start() {}
    void mark_h2d() {}
    void mark_compute() {}
    void stop() {}
    void capture(mavg_plan* plan) { mavg_get_timing(plan, &t_); }
    ProfileResult get_result() const
    {
        ProfileResult r;
        r.transfer_h2d_ms = t_.h2d_ms; r.compute_ms = t_.compute_ms; r.transfer_d2h_ms = t_.d2h_ms; r.total_ms = t_.total_ms;
        return r;
    }
};

class CpuTimer {
    std::chrono::steady_clock::time_point a_, b_;
public:
    void start() { a_ = std::chrono::steady_clock::now(); }
    void mark_h2d() {}
    void mark_compute() {}
    void stop() { b_ = std::chrono::steady_clock::now(); }
    ProfileResult get_result() const
    {
        ProfileResult r;
        r.compute_ms = std::chrono::duration<float, std::milli>(b_ - a_).count();
        r.total_ms = r.compute_ms;
        return r;
    }
};

// warm-up rounds, then the mean of `iterations` rounds (benchmark.h:116-132)
template <typename TimerType, typename Func>
ProfileResult benchmark(int iterations, int warmup, Func pipeline)
{
    TimerType timer;
    for (int i = 0; i < warmup; ++i) pipeline(timer);
    ProfileResult mean;
    for (int i = 0; i < iterations; ++i) {
        pipeline(timer);
        mean += timer.get_result();
    }
    mean.divide(iterations);
    return mean;
}
