// Compatibility name for the reference's (unused) benchmark_chrono.h: run_benchmark(numRuns, header, f)
// kept as a thin wall-clock shim; device work is timed by CUDA events inside libmavg (mavg_bench.h).
#pragma once
#include <chrono>
#include <cstdio>
#include "mavg_wav.h"

template <typename Func>
void run_benchmark(int numRuns, WAVHeader header, Func f)
{
    const double samples = header.bitsPerSample ? header.dataBytes / (header.bitsPerSample / 8) : 0;
    double us = 0;
    for (int i = 0; i < numRuns; ++i) {
        auto a = std::chrono::steady_clock::now();
        f();
        us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
    }
    const double mean_us = numRuns > 0 ? us / numRuns : 0;
    printf("--- Performance ---\nAverage Wall Clock Time: %.3f ms (%.1f us)\nThroughput:              %.3f Mega samples/sec\n\n",
           mean_us / 1e3, mean_us, mean_us > 0 ? samples / mean_us : 0.0);
}
