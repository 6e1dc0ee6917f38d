// mavg_workspace.h -- device workspace, CSV logger and error macro of the drop-in binaries.
//
// Interface twin of the reference's gpu_utils.h: CUDA_CHECK (:10-18), VecMode/MemoryMode (:20-29),
// warmupRounds/measurementRounds (:31-32), DspWorkspace (:67-160) and CsvLogger (:162-232).
// Differences, all consequences of moving the device work behind libmavg's C ABI:
//   * DspWorkspace owns a mavg_plan instead of raw cudaMalloc pointers; the zeroed halo the reference
//     allocates in front of the input (:112-123) no longer exists -- the kernel's TMA loads zero-fill
//     everything left of sample 0;
//   * host buffers are page-locked (mavg_host_alloc), replacing MemoryTraits<Standard>'s pageable
//     copies; MemoryMode::Unified is accepted and runs the same path (managed memory buys nothing
//     on a discrete B200);
//   * the CSV keeps the reference's 14 columns, in order, and appends GPUs, Dtype, Layout,
//     Gsamples_s, HBM_GBs, Pct_HBM_nominal, Pct_HBM_measured.
#pragma once

#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#include "mavg.h"
#include "mavg_bench.h"
#include "mavg_wav.h"

#define MAVG_CHECK(call)                                                                          \
    do {                                                                                          \
        int mavg_status_ = (call);                                                                \
        if (mavg_status_ != MAVG_OK) {                                                            \
            fprintf(stderr, "libmavg error: %s (%s) at %s:%d\n", mavg_strerror(mavg_status_),    \
                    mavg_last_error(), __FILE__, __LINE__);                                       \
            exit(EXIT_FAILURE);                                                                   \
        }                                                                                         \
    } while (0)

enum class VecMode { Scalar = 1, Int2 = 4, Int4 = 8 };   // accepted for source compatibility, unused
enum class MemoryMode { Standard, Unified };

const int warmupRounds = 5;         // gpu_utils.h:31
const int measurementRounds = 10;   // gpu_utils.h:32

const double kHbmNominalGBs = 8000.0;   // BASELINE.json north_star
// Measured HBM copy bandwidth the Pct_HBM_measured column is quoted against: the --hbm-peak flag of the binaries
// (run_benchmarks.py passes MEASURED_PEAKS.json's hbm_gbs), else MAVG_HBM_PEAK_GBS in the environment, else the
// figure B200_PROFILING.md gives for a B200.  Never a per-pool constant baked into the binary.
inline double& hbm_measured_override() { static double v = 0.0; return v; }
inline double hbm_measured_gbs()
{
    if (hbm_measured_override() > 0.0) return hbm_measured_override();
    if (const char* e = getenv("MAVG_HBM_PEAK_GBS")) {
        const double v = atof(e);
        if (v > 0.0) return v;
    }
    return 6650.0;
}

template <typename T, MemoryMode Mode = MemoryMode::Standard>
class DspWorkspace {
    static_assert(std::is_same<T, int16_t>::value || std::is_same<T, float>::value,
                  "libmavg filters int16 or float32 samples (no int64 staging is needed any more)");
    mavg_plan* plan_ = nullptr;
    T* h_in_ = nullptr;
    T* h_out_ = nullptr;
    DspWorkspace(const DspWorkspace&) = delete;
    DspWorkspace& operator=(const DspWorkspace&) = delete;

public:
    const size_t num_samples;
    const size_t valid_bytes;
    const size_t halo_elements = 0;   // no physical halo any more

    DspWorkspace(size_t samples, int grade, int num_channels, VecMode = VecMode::Scalar, size_t = 0,
                 int block_size = 0, int gpus = 1, int layout = MAVG_INTERLEAVED)
        : num_samples(samples), valid_bytes(samples * sizeof(T))
    {
        mavg_desc d;
        memset(&d, 0, sizeof d);
        d.struct_size = sizeof d;
        d.dtype = std::is_same<T, float>::value ? MAVG_F32 : MAVG_I16;
        d.layout = (uint32_t)layout;
        d.channels = (uint32_t)num_channels;
        d.frames = num_channels ? samples / (size_t)num_channels : 0;
        d.window = (uint32_t)grade;
        d.block_size = (uint32_t)block_size;
        if (gpus > 1) {
            d.num_devices = (uint32_t)gpus;
            for (int i = 0; i < gpus && i < MAVG_MAX_DEVICES; ++i) d.devices[i] = i;
        }
        MAVG_CHECK(mavg_plan_create(&d, &plan_));
        for (uint32_t r = 0; r < (gpus > 1 ? (uint32_t)gpus : 1u); ++r)
            MAVG_CHECK(mavg_plan_buffers(plan_, r, nullptr, nullptr));   // allocate now, like the reference's constructor
        MAVG_CHECK(mavg_host_alloc(valid_bytes, (void**)&h_in_));
        MAVG_CHECK(mavg_host_alloc(valid_bytes, (void**)&h_out_));
        // a sample count that is not a multiple of the channel count leaves samples % channels trailing elements no
        // frame owns: the reference's value-initialised vector holds zeros there, so does this buffer
        if (valid_bytes) memset(h_out_, 0, valid_bytes);
    }
    ~DspWorkspace()
    {
        mavg_plan_destroy(plan_);
        mavg_host_free(h_in_);
        mavg_host_free(h_out_);
    }
    mavg_plan* plan() const { return plan_; }
    T* host_in() const { return h_in_; }     // page-locked staging the caller fills once
    T* host_out() const { return h_out_; }
    // H2D + kernel + D2H: what every XxxGpuLoad of the reference does
    void run(GpuTimer& t)
    {
        MAVG_CHECK(mavg_run_host(plan_, h_in_, h_out_));
        t.capture(plan_);
    }
};

class CsvLogger {
    std::string filename_;
    static bool exists(const std::string& n) { struct stat b; return stat(n.c_str(), &b) == 0; }

public:
    explicit CsvLogger(const std::string& fname = "benchmark_results.csv") : filename_(fname) {}

    void log(const std::string& algo, const std::string& memory_mode, size_t N, int grade, int block_size,
             const ProfileResult& r, size_t in_bytes, size_t out_bytes = 0, int gpus = 1, const char* dtype = "int16",
             const char* layout = "interleaved")
    {
        if (!out_bytes) out_bytes = in_bytes;
        const bool fresh = !exists(filename_);
        FILE* f = fopen(filename_.c_str(), "a");
        if (!f) { fprintf(stderr, "Error: Could not open CSV file %s\n", filename_.c_str()); return; }
        if (fresh)
            fprintf(f, "Algorithm,MemoryMode,N_Samples,Grade,BlockSize,H2D_ms,Compute_ms,D2H_ms,Total_ms,Init_ms,"
                       "ColdStart_Total_ms,Bandwidth_GBs,Throughput_MSs,ColdStart_MSs,"
                       "GPUs,Dtype,Layout,Gsamples_s,HBM_GBs,Pct_HBM_nominal,Pct_HBM_measured\n");
        const double gb = (double)N * (double)(in_bytes + out_bytes) / 1e9;
        const double steady = r.total_ms / 1e3, cold = (r.initialization_ms + r.total_ms) / 1e3, kern = r.compute_ms / 1e3;
        const double hbm = kern > 0 ? gb / kern : 0.0;
        fprintf(f, "%s,%s,%zu,%d,%d,%g,%g,%g,%g,%g,%g,%g,%g,%g,%d,%s,%s,%g,%g,%g,%g\n", algo.c_str(), memory_mode.c_str(),
                N, grade, block_size, r.transfer_h2d_ms, r.compute_ms, r.transfer_d2h_ms, r.total_ms, r.initialization_ms,
                r.initialization_ms + r.total_ms, steady > 0 ? gb / steady : 0.0, steady > 0 ? N / 1e6 / steady : 0.0,
                cold > 0 ? N / 1e6 / cold : 0.0, gpus, dtype, layout, kern > 0 ? N / 1e9 / kern : 0.0, hbm,
                100.0 * hbm / (kHbmNominalGBs * gpus), 100.0 * hbm / (hbm_measured_gbs() * gpus));
        fclose(f);
        printf(">> Data saved to %s\n", filename_.c_str());
    }
};
