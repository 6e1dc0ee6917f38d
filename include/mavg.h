/*
 * mavg.h -- C ABI of libmavg, the B200 (sm_100a) moving-average filter library.
 *
 * Drop-in boundary.  The reference (SherTheCoder/digital_signal_processsing) has no
 * plugin/FFI layer: every `averager` binary statically contains its kernels.  The seam
 * this ABI occupies is the per-binary `XxxGpuLoad(workspace, grade, blockSize,
 * numOfChannels, GpuTimer&, samples, processedSamples)` function plus the
 * `DspWorkspace` it relies on:
 *
 *   basics/profilable_parallel_averager.cu:25-51   parallelAveragerGpuLoad
 *   basics/profilable_sm_averager.cu:47-74         smAveragerGpuLoad
 *   basics/profilable_sm_vload2.cu:64-92           vload2AveragerGpuLoad
 *   basics/profilable_sm_vload4.cu:90-145          vload4AveragerGpuLoad
 *   basics/hillis_steele_averager.cu:102-130       hillisSteeleAveragerGpuLoad
 *   basics/blelloch_scan_averager.cu:188-232       blellochAveragerGpuLoad
 *   gpu_utils.h:67-160                             DspWorkspace / MemoryTraits
 *   benchmark.h:72-96                              GpuTimer (h2d / compute / d2h ms)
 *
 * Semantics are those of the reference CPU path
 * (basics/profilable_moving_averager.cpp:14-37): causal, zero padded on the left,
 * per channel, same length and layout as the input, divides by the full window k
 * even while the window is filling.  int16 results are bit-identical to that
 * function (int64-exact sum, division truncating toward zero); float32 is the same
 * definition over reals, accurate to <= 1e-5 relative against an fp64 evaluation.
 *
 * Conventions: plain C, POD structs, uint64_t lengths, every entry point returns an
 * int status (0 = MAVG_OK, negative = error) and never exits or throws across the
 * ABI (the reference's CUDA_CHECK calls exit(), gpu_utils.h:10-18).  All filtering
 * runs in CUDA kernels; there is no CPU fallback: without a usable GPU the compute
 * entry points return MAVG_ERR_NO_DEVICE.
 */
#ifndef MAVG_H
#define MAVG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAVG_VERSION_MAJOR 0
#define MAVG_VERSION_MINOR 3
#define MAVG_VERSION_PATCH 0

typedef enum mavg_status {
    MAVG_OK = 0,
    MAVG_ERR_INVALID_ARG = -1,  /* null pointer, k == 0, channels == 0, bad enum ...       */
    MAVG_ERR_UNSUPPORTED = -2,  /* valid request this build cannot run                    */
    MAVG_ERR_CUDA = -3,         /* a CUDA runtime call failed; see mavg_last_error()      */
    MAVG_ERR_NO_DEVICE = -4,    /* no CUDA device / driver                                */
    MAVG_ERR_ALLOC = -5,        /* host or device allocation failed                       */
    MAVG_ERR_BLOCK_SIZE = -6,   /* block_size outside the reference's 32..1024, %32 rule  */
    MAVG_ERR_DRIVER = -7        /* driver entry point (tensor-map encode) unavailable     */
} mavg_status;

typedef enum mavg_dtype {
    MAVG_F32 = 0, /* float32 samples (WAV audioFormat 3, 32 bit) -- north-star extension  */
    MAVG_I16 = 1  /* int16 PCM samples, what wav_header.h:26-48 reads                     */
} mavg_dtype;

typedef enum mavg_layout {
    MAVG_INTERLEAVED = 0, /* [frame][channel], the reference/WAV layout                   */
    MAVG_PLANAR = 1       /* [channel][frame], a batch of independent mono signals        */
} mavg_layout;

/* Kernel family.  AUTO picks STREAM whenever its preconditions hold. */
typedef enum mavg_path {
    MAVG_PATH_AUTO = 0,
    MAVG_PATH_STREAM = 1,  /* TMA-staged shared-memory streaming kernels, float32 and int16:
                              mono / stereo / planar, 3..31 interleaved channels, many
                              interleaved channels (float32: multiples of 4 from 32 on,
                              int16: multiples of 8 from 64 on); direct window sums for
                              short windows, per-tile prefix scans above, a second TMA
                              stream for float32 windows beyond the shared-memory history */
    MAVG_PATH_GENERIC = 2  /* register sliding-window kernel on global memory; any shape  */
} mavg_path;

/* Synthetic input distributions, value = f(seed, global sample index); the same
 * formula is restated on the CPU in oracle/mavg_oracle.c for the tests. */
typedef enum mavg_dist {
    MAVG_DIST_U01 = 0,   /* U[0,1) on a 2^-24 lattice                                    */
    MAVG_DIST_USYM = 1,  /* U[-1,1)                                                      */
    MAVG_DIST_I16 = 2,   /* integer valued U[-32768,32767]                               */
    MAVG_DIST_DC1E4 = 3  /* 1e4 + U[-1,1)                                                */
} mavg_dist;

/* Windowed reduction a plan computes ("other windowed reductions").  Window, zero
 * padding on the left and the division by the full k while the window fills are the same for both. */
typedef enum mavg_op {
    MAVG_OP_MEAN = 0, /* the moving average of the reference (basics/profilable_moving_averager.cpp:14-37)    */
    MAVG_OP_RMS = 1   /* moving RMS: sqrt(mean of squares).  float32: <= 1e-5 relative against fp64; int16: exact
                         int64 sum of squares, y = min(32767, trunc(sqrt((double) sum / k))).  float32 mono, stereo
                         and planar signals run on the TMA streaming kernel (squares on load, root on store); every
                         other shape runs on the generic kernel                                               */
} mavg_op;

#define MAVG_MAX_DEVICES 16

/* Optional tuning overrides; 0 = library default. */
typedef struct mavg_tuning {
    uint32_t threads;        /* threads per CTA.  float32 stream kernel: 256 or 512.  int16 with 3+ interleaved
                                channels: 128 (two CTAs per SM), 256 (one CTA of 224 / 256 / 384 threads with the
                                long runs), 512 (short runs; 3 / 4 / 6 / 8 channels).  int16 column kernel
                                (64+ channels): 512 = 16 warps x 16 frames instead of 8 x 32.  Far-lag kernel: 512 =
                                tiles of 512 x 16 samples instead of 384 x 16                                   */
    uint32_t run;            /* samples per thread run of the float32 stream kernel: 16 or 32                  */
    uint32_t prefetch;       /* tiles in flight ahead of the one being filtered            */
    uint32_t ctas_per_sm;    /* resident CTAs per SM the grid is sized for                 */
    uint32_t chunks_per_cta; /* contiguous tile ranges each CTA walks (>=1)                */
    uint32_t direct_max_k;   /* largest k*channels served by direct group sums (default 256); int16 column
                                kernel: 1 = results stored from registers instead of staging tiles + TMA stores */
    uint32_t slice_bytes;    /* mavg_run_host: bytes per pipelined H2D/kernel/D2H slice (default 16 MiB) */
    uint32_t overlap;        /* programmatic dependent launch of the streaming kernels (the next launch's prologue
                                runs under the previous kernel's tail; nothing is written before the previous
                                kernel has completed).  0 = library default (1); 1 = wait for the previous kernel
                                before the first load -- always safe; 2 = wait only before the first store: the
                                caller asserts that this plan's INPUT (and halo) buffers are not written by work
                                queued earlier on the same stream, so tile loads may start early; 3 = off    */
} mavg_tuning;

/* Plan description.  Replaces the DspWorkspace constructor arguments
 * (gpu_utils.h:91-97: num_samples, grade, num_channels, VecMode, scratch). */
typedef struct mavg_desc {
    uint32_t struct_size;  /* sizeof(mavg_desc), for forward compatibility                */
    uint32_t dtype;        /* mavg_dtype                                                  */
    uint32_t layout;       /* mavg_layout                                                 */
    uint32_t channels;     /* >= 1 (header.numChannels)                                   */
    uint64_t frames;       /* frames per channel; total samples N = frames * channels     */
    uint32_t window;       /* k >= 1 ("grade"/"point"); k > frames is allowed             */
    uint32_t block_size;   /* the reference's <block_size> argv: 0 = not given, otherwise
                              validated like basics/profilable_sm_vload4.cu:231 and then
                              only a hint                                                 */
    uint32_t path;         /* mavg_path                                                   */
    uint32_t num_devices;  /* 0 or 1: the current device; >1: one process drives several
                              GPUs, frames (interleaved/mono) or channels (planar) are
                              sharded contiguously over `devices`                         */
    int32_t devices[MAVG_MAX_DEVICES];
    uint64_t first_frame;  /* global index of frame 0 of this plan inside a longer signal
                              (one-process-per-GPU sharding); 0 otherwise.  When > 0 the
                              left context is supplied through mavg_run_device_halo.      */
    mavg_tuning tuning;
    uint32_t op;           /* mavg_op                                                     */
    uint32_t reserved;     /* must be 0                                                   */
} mavg_desc;

/* Phase times of the last run, as GpuTimer::get_result (benchmark.h:88-96) reports
 * them; with several devices each field is the maximum over devices. */
typedef struct mavg_timing {
    float h2d_ms;
    float compute_ms;
    float d2h_ms;
    float total_ms;
} mavg_timing;

typedef struct mavg_info {
    uint32_t path;             /* mavg_path actually selected                             */
    uint32_t mode;             /* stream kernel arithmetic: 0 direct group sums, 1 tile-rebased
                                  prefix scan, 2 additions-only (k <= 8), 3 column kernel,
                                  4 few-channel kernel (3..31 interleaved channels), 5 far-lag
                                  kernel (mono / planar float32, windows beyond the ring), 6 int16
                                  mono / stereo / planar: exclusive scan of run deltas (any k),
                                  7 (generic path) single-pass prefix sum + difference: far windows,
                                  up to 8 interleaved channels                                  */
    uint32_t threads, run;     /* stream kernel shape                                     */
    uint32_t tile_samples;     /* samples per shared-memory tile                          */
    uint32_t history_tiles;    /* tiles of left context each tile range replays           */
    uint32_t stages;           /* TMA ring depth                                          */
    uint32_t grid;             /* CTAs launched per device                                */
    uint32_t smem_bytes;       /* dynamic shared memory per CTA                           */
    uint32_t launches_per_run; /* kernel launches one mavg_run_device issues, all devices */
    uint32_t num_devices;
    uint32_t reserved;
    uint64_t halo_frames;      /* left-context frames a shard plan wants (>= k)           */
    uint64_t shard_frames[MAVG_MAX_DEVICES]; /* frames (or channels, planar) per device   */
} mavg_info;

typedef struct mavg_plan mavg_plan;

/* Library version as major*10000 + minor*100 + patch. */
int mavg_version(void);

/* Static description of a status code. */
const char *mavg_strerror(int status);

/* Detail of the calling thread's most recent failure ("" if none). */
const char *mavg_last_error(void);

/* Number of usable CUDA devices (0 when there is no driver/GPU); never fails. */
int mavg_device_count(void);

/* Builds a plan: validates the description, picks the kernel family, creates
 * per-device streams/events.  Device buffers are allocated lazily by the first call
 * that needs them (mavg_run_host, mavg_fill_synthetic, mavg_plan_buffers).
 * Replaces DspWorkspace::DspWorkspace (gpu_utils.h:91-125). */
int mavg_plan_create(const mavg_desc *desc, mavg_plan **plan);

/* Releases everything the plan owns (DspWorkspace::~DspWorkspace, gpu_utils.h:127-131). */
int mavg_plan_destroy(mavg_plan *plan);

int mavg_plan_info(const mavg_plan *plan, mavg_info *info);

/* H2D + kernel(s) + D2H from/to caller-owned host memory holding the whole signal in
 * the plan's layout -- what every XxxGpuLoad does (e.g.
 * basics/profilable_sm_vload4.cu:90-145).  Blocks until the output is in h_out.
 * h_out must not alias h_in.  For a shard plan (desc.first_frame > 0) the
 * info.halo_frames frames of left context must sit in host memory immediately before
 * h_in (h_in points into a [halo | shard] buffer). */
int mavg_run_host(mavg_plan *plan, const void *h_in, void *h_out);

/* One upload, many windows -- what a sweep over grades does with one WAV file
 * (basics/run_benchmarks.py:21-47 runs the binaries once per grade on the same input): `count` plans that
 * describe the SAME signal (dtype, layout, channels, frames, first_frame, one and the same device) and differ in
 * window / op / tuning are run over h_in; plan i's output goes to h_out[i].  The input crosses the host link once
 * instead of `count` times: slices of it are uploaded, every plan's kernel runs on the slice as soon as it has
 * arrived, and the `count` results of the slice go back while the next slice arrives.  Results are bit-identical
 * to `count` calls of mavg_run_host.  Interleaved / mono single-device plans; other plans (planar batches,
 * multi-device plans, pageable outputs) are served by calling mavg_run_host once per plan.  Timing of the whole
 * sweep (exposed phase times) is reported through plans[0] (mavg_get_timing).  For shard plans the largest
 * info.halo_frames of the plans must sit in host memory in front of h_in.  Blocks until every output is in place. */
int mavg_run_host_sweep(mavg_plan *const *plans, uint32_t count, const void *h_in, void *const *h_out);

/* Kernel(s) only, on device-resident data: d_in[r] / d_out[r] are device r's shard
 * (info.shard_frames[r] frames; for one device simply the whole signal).  Pointers must
 * be 16-byte aligned for the stream path (otherwise the generic path runs).
 * Asynchronous: work is enqueued on the plan's streams; compute_ms is available after
 * mavg_synchronize.  This is the call the roofline is measured on. */
int mavg_run_device(mavg_plan *plan, const void *const *d_in, void *const *d_out);

/* Single-device shard of a longer signal (desc.first_frame > 0): d_halo points at
 * info.halo_frames frames that immediately precede d_in's first frame in the signal.
 * It may be local memory or a peer-mapped pointer into the left neighbour's shard
 * (P2P over NVLink: the kernel's TMA loads read it in place).  NULL = zeros. */
int mavg_run_device_halo(mavg_plan *plan, const void *d_in, void *d_out, const void *d_halo);

/* Blocks until all work enqueued by the plan has finished. */
int mavg_synchronize(mavg_plan *plan);

/* Phase times of the last completed run (valid after mavg_synchronize / mavg_run_host). */
int mavg_get_timing(mavg_plan *plan, mavg_timing *timing);

/* Switches the plan's own CUDA-event phase timing off (0) or on (non-zero, the default).
 * Off saves four event records per run when the caller times the stream itself. */
int mavg_enable_timing(mavg_plan *plan, int enable);

/* Makes the plan enqueue on a caller-owned cudaStream_t (single-device plans). */
int mavg_set_stream(mavg_plan *plan, void *cuda_stream);

/* Plan-owned device buffers of device index `rank` (allocated on first use). */
int mavg_plan_buffers(mavg_plan *plan, uint32_t rank, void **d_in, void **d_out);

/* Fills the plan-owned input buffers with the synthetic signal (seed, dist), indexed by
 * global sample position, so sharded plans generate exactly the slices of one signal. */
int mavg_fill_synthetic(mavg_plan *plan, uint64_t seed, int dist);

/* Same generator on an arbitrary device buffer of the current device:
 * dst[i] = gen(seed, first_index + i), i < n.  Asynchronous on `cuda_stream`. */
int mavg_fill_synthetic_device(void *d_dst, int dtype, uint64_t n, uint64_t first_index,
                               uint64_t seed, int dist, void *cuda_stream);

/* Runs plan-owned input -> plan-owned output (after mavg_fill_synthetic). */
int mavg_run_owned(mavg_plan *plan);

/* Reusable primitive: per-channel inclusive prefix sum of an interleaved device signal in ONE pass
 * (decoupled look-back), what recursive_hillis_steele / recursive_blelloch compute with multi-level
 * recursion (basics/hillis_steele_averager.cu:69-84, basics/blelloch_scan_averager.cu:134-167):
 *   out[f*C + c] = sum_{j <= f} in[j*C + c]
 * dtype MAVG_I16 -> int64 output (exact), MAVG_F32 -> float64 output.  1 <= channels <= 8.
 * Asynchronous on `cuda_stream`; scratch is taken from and returned to a library-owned stream-ordered memory pool.
 * The moving-average kernels do not call it (they never need a prefix over the whole signal). */
int mavg_prefix_sum(int dtype, const void *d_in, void *d_out, uint64_t frames, uint32_t channels,
                    void *cuda_stream);

/* CUDA IPC helpers for one-process-per-GPU sharding: export a 64-byte handle of a
 * cudaMalloc'ed buffer, open it in the neighbour process, close it. */
int mavg_ipc_export(const void *d_ptr, void *handle64);
int mavg_ipc_open(const void *handle64, void **d_ptr);
int mavg_ipc_close(void *d_ptr);

/* Plain device allocation helpers (cudaMalloc / cudaFree on the current device), so
 * that buffers exported over IPC do not come from a caching allocator. */
int mavg_device_alloc(uint64_t bytes, void **d_ptr);
int mavg_device_free(void *d_ptr);

/* Box-filter cascade: the plan's moving average applied `passes` times in a row on device-resident shards
 * (three passes approximate a Gaussian; what running a reference binary on its own output does, without the
 * round trips).  Pass i reads what pass i-1 wrote, including the left neighbour's tail on sharded plans, so the
 * library orders the passes across devices; the last pass lands in d_out.  d_scratch[r] (same size as the
 * shard) holds intermediates; pass NULL to let the plan allocate them.  int16 truncates after every pass, like
 * the reference would.  Whole-signal plans only (desc.first_frame == 0).  Builds on the kernel launches of
 * basics/profilable_sm_vload4.cu:135 (box-filter cascades). */
int mavg_run_cascade(mavg_plan *plan, const void *const *d_in, void *const *d_out, void *const *d_scratch,
                     uint32_t passes);

/* Page-locked host memory for mavg_run_host buffers (cudaHostAlloc / cudaFreeHost): copies from
 * pinned memory run at full PCIe speed and overlap with the kernels.  Replaces the host half
 * of MemoryTraits (gpu_utils.h:33-65). */
int mavg_host_alloc(uint64_t bytes, void **h_ptr);
int mavg_host_free(void *h_ptr);

/* Page-lock memory the caller already owns (cudaHostRegister / cudaHostUnregister), e.g. the std::vector
 * buffers the reference hands to its XxxGpuLoad functions (basics/profilable_sm_vload4.cu:90-91), whose
 * pageable cudaMemcpy (gpu_utils.h:43,47) runs 7x slower than pinned copies on this hardware (measured:
 * 1.75 Gsamples/s pageable (2.45 since mavg_run_host issues pageable D2H copies from a helper thread), 10.3
 * registered, 11.8 with mavg_host_alloc buffers through mavg_run_host, 2^28
 * float32 samples).  Registering costs about 0.35 s per GiB, i.e. it pays off from the fifth call on the same
 * buffers: do it once per buffer, not per call.  The caller unregisters BEFORE freeing the memory.
 * MAVG_ERR_ALLOC when the range cannot be locked (the buffer then still works, at pageable speed). */
int mavg_host_register(void *h_ptr, uint64_t bytes);
int mavg_host_unregister(void *h_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MAVG_H */
