// mavg_kernels.cuh -- device code of libmavg (sm_100a only).
//
// Replaces every __global__ in the reference's basics/*.cu:
//   averager_kernel (naive / smem / int2 / int4)      -> stream_f32_kernel, MODE 0
//   hillis_steele + uniform_add + prefix-diff averager  -> stream_f32_kernel, MODE 1
//   blelloch_scan_inclusive + blelloch_uniform_add      -> stream_f32_kernel, MODE 1
// plus generic_kernel for every shape the streaming kernel does not take.
//
// Design of stream_f32_kernel (one signal = one row-major [rows][32] float tensor):
//   * persistent CTAs; each walks contiguous ranges of tiles ("chunks") of a signal,
//     so the k-sample left context is re-read once per chunk, not once per tile;
//   * tiles (NT*R samples, 16-32 KB) arrive through a ring of TMA tensor loads
//     (cp.async.bulk.tensor + mbarrier complete_tx, SWIZZLE_128B so that each
//     thread's 64/128-byte run is read with conflict-free LDS.128); rows before
//     sample 0 and after the end are zero-filled by the TMA unit, which IS the
//     reference's zero halo (gpu_utils.h:112-121) without allocating one;
//   * every thread owns a run of R consecutive outputs: window sum at the run start
//     from group totals (MODE 0: direct sum of the <=16 totals between, no
//     cancellation; MODE 1: difference of a tile-rebased prefix scan of the totals,
//     warp-shuffle scan + per-tile totals), then slides w += x[i] - x[i-k];
//   * results go to a swizzled staging tile and leave through a TMA tensor store,
//     which also clips the rows past the end of the signal.
// HBM traffic: 4 B read + 4 B written per sample, plus history tiles per chunk.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mavg {

// ----------------------------------------------------------------------------------
// Synthetic generator (same formula as oracle/mavg_oracle.c: oracle_mix64 / gen_f32)
// ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t index)
{
    uint64_t z = index + seed * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float synth_f32(uint64_t seed, uint64_t index, int dist)
{
    const uint64_t z = mix64(seed, index);
    const float u = __fmul_rn((float)(uint32_t)(z >> 40), 1.0f / 16777216.0f);
    switch (dist) {
    case 0: return u;
    case 1: return __fsub_rn(__fmul_rn(2.0f, u), 1.0f);
    case 2: return (float)((int32_t)(z >> 48) - 32768);
    case 3: return __fadd_rn(10000.0f, __fsub_rn(__fmul_rn(2.0f, u), 1.0f));
    default: return 0.0f;
    }
}

__global__ void fill_f32_kernel(float* __restrict__ dst, uint64_t n, uint64_t first, uint64_t seed, int dist)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = synth_f32(seed, first + i, dist);
}

__global__ void fill_i16_kernel(int16_t* __restrict__ dst, uint64_t n, uint64_t first, uint64_t seed)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = (int16_t)((int32_t)(mix64(seed, first + i) >> 48) - 32768);
}

// ----------------------------------------------------------------------------------
// Generic kernel: any dtype / channel count / alignment / window.
// One thread = one channel x RG consecutive frames; consecutive threads take
// consecutive channels first, so many-channel interleaved signals coalesce.
// float accumulates in fp64 (the k-term fresh sum would otherwise cost k*eps);
// int16 accumulates exactly and divides with C truncation like the reference
// (basics/profilable_moving_averager.cpp:23,33).
// ----------------------------------------------------------------------------------
struct GenericParams {
    uint64_t frames;       // frames of this signal (per channel)
    uint64_t out_begin;    // first frame to produce
    uint64_t out_end;      // one past the last frame to produce
    uint64_t halo_frames;  // frames available in `halo` (left context of frame 0)
    uint64_t sig_stride;   // elements between consecutive planar signals (blockIdx.y)
    uint32_t channels;     // interleave factor inside one signal
    uint32_t k;
    const void* bsum;      // optional: per-channel sums of 64-frame blocks [signal][block][channel] (Acc type)
    uint64_t nblk;         // blocks per signal in bsum
    uint32_t rms;          // mavg_op: 0 moving average, 1 moving RMS (samples squared on load, root of the mean on store)
    uint32_t pad_;
};

template <typename T> struct GenericAcc;
template <> struct GenericAcc<float> {
    typedef double type;
    __device__ static __forceinline__ float finish(double w, uint32_t k, double inv) { (void)k; return (float)(w * inv); }
    __device__ static __forceinline__ float finish_rms(double w, uint32_t k, double inv)
    {
        (void)k;
        return (float)sqrt(fmax(w, 0.0) * inv);
    }
};
template <> struct GenericAcc<int16_t> {
    typedef long long type;
    __device__ static __forceinline__ int16_t finish(long long w, uint32_t k, double inv)
    {
        (void)inv;
        // |w| <= k * 32768: 32-bit division whenever that fits.
        if (k < 65536u) return (int16_t)((int)w / (int)k);
        return (int16_t)(w / (long long)k);
    }
    // exact int64 sum of squares; IEEE double division and square root, truncated, saturated (oracle_mrms_i16)
    __device__ static __forceinline__ int16_t finish_rms(long long w, uint32_t k, double inv)
    {
        (void)inv;
        const double r = sqrt((double)w / (double)k);
        return (int16_t)(r > 32767.0 ? 32767.0 : r);
    }
};
// a sample as it enters a window sum: itself, or its square for the moving RMS
template <typename Acc, bool RMS, typename T>
__device__ __forceinline__ Acc gen_term(T v)
{
    const Acc a = (Acc)v;
    if constexpr (RMS) return a * a;
    else return a;
}

// Sums of RG-frame blocks per channel, so that long windows start from k/RG block sums instead of k
// samples.  Same thread mapping as generic_kernel: consecutive threads = consecutive channels.
template <typename T, int RG, bool RMS = false>
__global__ void __launch_bounds__(256) block_sums_kernel(const T* __restrict__ x, typename GenericAcc<T>::type* __restrict__ bs,
                                                         uint64_t frames, uint64_t sig_stride, uint32_t C, uint64_t nblk)
{
    typedef typename GenericAcc<T>::type Acc;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = (uint32_t)(gid % C);
    const uint64_t b = gid / C;
    if (b >= nblk) return;
    x += (uint64_t)blockIdx.y * sig_stride;
    const uint64_t f0 = b * RG;
    const uint64_t f1 = (f0 + RG < frames) ? f0 + RG : frames;
    Acc a = 0;
    for (uint64_t f = f0; f < f1; ++f) a += gen_term<Acc, RMS>(x[f * C + c]);
    bs[((uint64_t)blockIdx.y * nblk + b) * C + c] = a;
}

// F32SLIDE (float only, k >= 9): the k-term start sum is still formed in fp64, but the RG sliding
// updates run in fp32 (error <= RG * 2^-24 relative to the window sum), which removes three
// fp32<->fp64 conversions per sample.  Tiny windows keep the fp64 update: their sums can be ~0.
template <typename T, int RG, bool F32SLIDE, bool RMS = false>
__global__ void __launch_bounds__(256) generic_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                      const T* __restrict__ halo, const GenericParams p)
{
    typedef typename GenericAcc<T>::type Acc;
    const uint32_t C = p.channels;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = (uint32_t)(gid % C);
    const uint64_t run = gid / C;
    const uint64_t f0 = p.out_begin + run * (uint64_t)RG;
    if (f0 >= p.out_end) return;
    const uint64_t f1 = (f0 + RG < p.out_end) ? f0 + RG : p.out_end;
    x += (uint64_t)blockIdx.y * p.sig_stride;
    y += (uint64_t)blockIdx.y * p.sig_stride;
    const uint64_t k = p.k;
    const double inv = 1.0 / (double)p.k;

    // window sum over frames [f0-k, f0): negative frames come from the halo, else zero
    Acc w = 0;
    if (p.bsum != nullptr && f0 >= k && f0 % RG == 0) {
        // whole RG-frame blocks from the block-sum table, the ragged head from the samples
        const Acc* bs = (const Acc*)p.bsum + (uint64_t)blockIdx.y * p.nblk * C;
        const uint64_t lo = f0 - k;
        const uint64_t jb = (lo + RG - 1) / RG;
        for (uint64_t f = lo; f < jb * RG; ++f) w += gen_term<Acc, RMS>(x[f * C + c]);
        for (uint64_t j = jb; j < f0 / RG; ++j) w += bs[j * C + c];
    } else {
        const long long lo = (long long)f0 - (long long)k;
        long long j = lo;
        if (j < 0) {
            if (halo != nullptr) {
                long long hlo = -(long long)p.halo_frames;
                if (j < hlo) j = hlo;
                const long long hend = (long long)f0 < 0 ? (long long)f0 : 0;
                for (; j < hend; ++j) w += gen_term<Acc, RMS>(halo[(uint64_t)(j + (long long)p.halo_frames) * C + c]);
            }
            j = 0;
        }
        for (; j < (long long)f0; ++j) w += gen_term<Acc, RMS>(x[(uint64_t)j * C + c]);
    }
    if constexpr (F32SLIDE) {
        float wf = (float)w;
        const float invf = 1.0f / (float)p.k;
        if (f0 >= k) {  // steady state: no halo / zero-padding cases inside the run
            const T* xo = x + (f0 - k) * C + c;
            const T* xn = x + f0 * C + c;
            T* yo = y + f0 * C + c;
            const uint32_t cnt = (uint32_t)(f1 - f0);
#pragma unroll 8
            for (uint32_t i = 0; i < cnt; ++i) {
                wf += (float)xn[(uint64_t)i * C] - (float)xo[(uint64_t)i * C];
                yo[(uint64_t)i * C] = (T)(wf * invf);
            }
            return;
        }
        for (uint64_t f = f0; f < f1; ++f) {
            float old = 0.f;
            if (f >= k) {
                old = (float)x[(f - k) * C + c];
            } else if (halo != nullptr) {
                const uint64_t back = k - f;
                if (back <= p.halo_frames) old = (float)halo[(p.halo_frames - back) * C + c];
            }
            wf += (float)x[f * C + c] - old;
            y[f * C + c] = (T)(wf * invf);
        }
    } else {
    for (uint64_t f = f0; f < f1; ++f) {
        Acc old = 0;
        if (f >= k) {
            old = gen_term<Acc, RMS>(x[(f - k) * C + c]);
        } else if (halo != nullptr) {
            const uint64_t back = k - f;  // frames before frame 0
            if (back <= p.halo_frames) old = gen_term<Acc, RMS>(halo[(p.halo_frames - back) * C + c]);
        }
        w += gen_term<Acc, RMS>(x[f * C + c]) - old;
        y[f * C + c] = RMS ? GenericAcc<T>::finish_rms(w, p.k, inv) : GenericAcc<T>::finish(w, p.k, inv);
    }
    }
}

// The few trailing frames a streaming kernel leaves (samples past the last whole 128-byte row): one CTA per signal
// sums the k frames in front of them IN PARALLEL (256 threads, fp64 / int64, fixed partition and order, so the
// result does not depend on how the signal was sliced), then one thread slides over the handful of outputs.
// (generic_kernel would give the whole k-term start sum to a single thread: 0.14 ms at k = 4096, milliseconds for
// the far-lag kernel's windows.)
template <typename T, bool RMS = false>
__global__ void __launch_bounds__(256) tail_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ halo,
                                                   const GenericParams p)
{
    typedef typename GenericAcc<T>::type Acc;
    __shared__ Acc red[8];
    const uint32_t C = p.channels;
    x += (uint64_t)blockIdx.x * p.sig_stride;
    y += (uint64_t)blockIdx.x * p.sig_stride;
    const long long k = (long long)p.k;
    const long long hf = (long long)p.halo_frames;
    const double inv = 1.0 / (double)p.k;
    for (uint32_t c = 0; c < C; ++c) {
        Acc a = 0;
        for (long long f = (long long)p.out_begin - k + threadIdx.x; f < (long long)p.out_begin; f += 256) {
            if (f >= 0) a += gen_term<Acc, RMS>(x[(uint64_t)f * C + c]);
            else if (halo != nullptr && f >= -hf) a += gen_term<Acc, RMS>(halo[(uint64_t)(f + hf) * C + c]);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) a += __shfl_down_sync(0xffffffffu, a, d);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            Acc w = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
            for (uint64_t f = p.out_begin; f < p.out_end; ++f) {
                Acc old = 0;
                const long long fo = (long long)f - k;
                if (fo >= 0) old = gen_term<Acc, RMS>(x[(uint64_t)fo * C + c]);
                else if (halo != nullptr && fo >= -hf) old = gen_term<Acc, RMS>(halo[(uint64_t)(fo + hf) * C + c]);
                w += gen_term<Acc, RMS>(x[f * C + c]) - old;
                y[f * C + c] = RMS ? GenericAcc<T>::finish_rms(w, p.k, inv) : GenericAcc<T>::finish(w, p.k, inv);
            }
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------
// PTX helpers (mbarrier, TMA tensor load/store, proxy fence)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            uint64_t hint)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, uint64_t hint)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Programmatic dependent launch: `launch_dependents` lets the next kernel in the stream become resident as soon as
// every CTA of this grid has issued it (or exited); `wait` blocks the calling thread until the PREVIOUS grid has
// completed and its memory operations are visible.  Both are no-ops in a launch without the attribute.
__device__ __forceinline__ void gdc_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gdc_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// SWIZZLE_128B: 16-byte chunk index bits [4:6] ^= 128-byte row index bits [7:9].
// Valid on absolute shared addresses because every tile buffer is 1024-byte aligned.
__device__ __forceinline__ uint32_t swz(uint32_t addr) { return addr ^ ((addr >> 3) & 0x70u); }

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;  // L2 cache hint: streaming input
constexpr uint64_t kEvictNormal = 0x1000000000000000ull; // L2 cache hint: data that is read again (far-lag kernel)
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;   // L2 cache hint: keep (far-lag kernel: own tiles come back as lag boxes)

// Pre-swizzled tile offsets (few-channel kernels' scalar accesses, int16 kernel's 16-byte chunks).  Tile, staging and ring
// bases are multiples of 1024 bytes, so the SWIZZLE_128B XOR (bits 4-6 with bits 7-9) of base + x only depends
// on x: swz(base + x) = base + (x ^ ((x >> 3) & 0x70)), also for negative x in two's complement (offsets into
// history tiles).  Computing these once per thread removes ~18 of 33 instructions per sample (ncu source view of
// the first version: LOP3 + SHF + IMAD + VIADD + ISETP address arithmetic).
__device__ __forceinline__ int pre_swz(int x) { return x ^ ((x >> 3) & 0x70); }
// address of pre-swizzled offset xs (relative to the current tile, negative = history) in the stage ring:
// b0 = ring + st * tile_bytes, b1 = b0 + ring_bytes (wrapped), neg_st = -(st * tile_bytes)
__device__ __forceinline__ uint32_t ring_addr(int xs, uint32_t b0, uint32_t b1, int neg_st)
{
    return (uint32_t)xs + (xs < neg_st ? b1 : b0);
}


// ----------------------------------------------------------------------------------
// Streaming kernel
// ----------------------------------------------------------------------------------
struct StreamParams {
    float inv_k;
    uint32_t k;
    uint32_t n_full;       // whole thread-groups strictly between the lag group and the own group
    uint32_t m_part;       // leading elements of the lag run that complete the window at the run start (1..R)
    uint32_t lag_chunks;   // ceil(k / 4): 16-byte chunks from the lag run's aligned start to the own run
    int32_t tiles_per_signal;
    int32_t chunk_tiles;        // output tiles per chunk
    int32_t chunks_per_signal;
    int32_t total_chunks;       // chunks_per_signal * signals
    int32_t hist_tiles;         // H: history tiles replayed (loaded, summarised, not written) per chunk
    int32_t stages;             // S = H + 1 + prefetch
    int32_t prefetch;           // P
    int32_t has_halo;           // tiles before tile 0 come from halo_map instead of zero fill
    int32_t pdl;                // programmatic dependent launch: 0 off, 1 wait for the previous grid before the first
                                // load, 2 before the first store only (caller: inputs not written by earlier work)
    // int16 paths: multiply-high constants of the exact truncating division (plan_stream_i16 / plan_fewc)
    uint32_t div_mul;
    uint32_t div_shift;
    uint32_t wscale;            // stream_i16_kernel: dp2a weight magnitude (1; 2 when k == 2, divisor then 4)
    uint32_t wtab[2][20];       // dp2a byte weights selecting, per 32-bit lag word, the halves that belong to the
                                // head of the lag run (first m_part elements) for channel 0 / channel 1
};

// Shared-memory carve-up (bytes), shared by host (size) and device (offsets).
__host__ __device__ inline uint32_t stream_smem_bytes(int NT, int R, int S, int H, int C = 1)
{
    const uint32_t TB = (uint32_t)NT * R * 4;
    return 1024u                           // alignment slack
           + (uint32_t)S * TB              // input ring
           + 2u * TB                       // output staging (double buffered)
           + (uint32_t)(H + 2) * NT * C * 4  // per-thread group totals (MODE 0) / warp-inclusive prefixes (MODE 1)
           + (uint32_t)(H + 2) * 32 * C * 4  // per-tile exclusive warp offsets, [31] = tile total (MODE 1)
           + 2u * 32 * C * 4               // raw warp totals, double buffered by iteration parity (MODE 1)
           + (uint32_t)S * 8;              // mbarriers
}

// ----------------------------------------------------------------------------------
// TileRing -- the producer/consumer protocol every flat streaming kernel shares:
//   * an S-stage ring of TMA tiles (S = history H + 1 + prefetch P) filled by thread 0, one mbarrier per
//     stage, stages recycled by the single __syncthreads each tile costs;
//   * persistent walking of contiguous tile ranges ("chunks"): H history tiles are replayed per chunk;
//   * a double-buffered output staging tile whose TMA store is issued one iteration late (after the next
//     __syncthreads) and whose reuse waits on cp.async.bulk.wait_group.read.
// Shared-memory carve-up: [ring S*TB][staging 2*TB][kernel-specific summaries ...][S mbarriers].
// ----------------------------------------------------------------------------------
// TB_ / ROWS_ = tile bytes / 128-byte rows per tile as compile-time constants, or 0 for run-time values passed to
// setup() (the few-channel kernel's tile depends on the channel count).
// INPLACE_: no staging tiles -- a kernel that reads nothing but its own run from the current tile writes the results
// over it and the TMA store leaves from the ring stage itself (stages: P loading, 1 computing, 1 storing; S >= P + 2).
// The refill of a stage waits until the store issued one iteration earlier has read it (wait_group.read 1).
template <uint32_t TB_, int ROWS_, bool INPLACE_ = false>
struct TileRing {
    uint32_t tb_rt;
    int rows_rt;
    __device__ __forceinline__ uint32_t tbv() const { return TB_ ? TB_ : tb_rt; }
    __device__ __forceinline__ int rowsv() const { return ROWS_ ? ROWS_ : rows_rt; }
    uint32_t ring, ring_bytes, outb, bars;
    int S, H, P, GS;
    uint32_t it;       // tiles streamed so far by this CTA (history + output), never reset
    int st;            // it % S
    int slot;          // it % GS, the summary slot of the current tile
    uint32_t otiles;   // output tiles produced so far (selects the staging buffer)
    // deferred TMA store (thread 0): st_pending = staged but not issued, st_inflight = issued, read not confirmed
    bool st_pending, st_inflight;
    int st_tile, st_sig;
    uint32_t st_buf;
    const CUtensorMap *in_map, *out_map, *halo_map;
    bool has_halo;
    bool pdl_wait_loads; // thread 0: griddepcontrol.wait before the first tile load (pdl == 1)
    bool gdc_pending;    // thread 0: griddepcontrol.wait still owed before the first global store (pdl == 2)
    int row_base;        // rows in front of row 0 of in_map (far-lag kernel: left context contiguous with the input)
    uint64_t load_hint;  // L2 policy of the tile loads
    uint64_t store_hint; // L2 policy of the tile stores (0 = none)

    // returns the first shared address after the staging tiles (where the kernel puts its summaries)
    __device__ __forceinline__ uint32_t setup(uint32_t smem_base, const StreamParams& p, const CUtensorMap* in,
                                              const CUtensorMap* out, const CUtensorMap* halo, uint32_t tb_runtime = 0,
                                              int rows_runtime = 0)
    {
        tb_rt = tb_runtime;
        rows_rt = rows_runtime;
        ring = (smem_base + 1023u) & ~1023u;
        S = p.stages; H = p.hist_tiles; P = p.prefetch; GS = H + 2;
        ring_bytes = (uint32_t)S * tbv();
        outb = ring + ring_bytes;
        it = 0; st = 0; slot = 0; otiles = 0;
        st_pending = st_inflight = false; st_tile = st_sig = 0; st_buf = 0;
        in_map = in; out_map = out; halo_map = halo; has_halo = p.has_halo != 0;
        gdc_pending = p.pdl == 2;
        if (p.pdl != 0 && threadIdx.x == 0) {
            // the next launch may take this SM as soon as every CTA of this grid got here or exited; its own
            // griddepcontrol.wait keeps it from storing (pdl 2) or loading (pdl 1) before this grid has completed
            gdc_launch_dependents();
        }
        pdl_wait_loads = p.pdl == 1;
        row_base = 0; load_hint = kEvictFirst; store_hint = 0;
        return INPLACE_ ? outb : outb + 2u * tbv();
    }
    __device__ __forceinline__ void init_barriers(uint32_t bars_addr)
    {
        bars = bars_addr;
        if (threadIdx.x == 0) {
            prefetch_tmap(in_map);
            prefetch_tmap(out_map);
            if (has_halo) prefetch_tmap(halo_map);
            for (int s = 0; s < S; ++s) mbar_init(bars + 8u * s, 1);
            fence_mbar_init();
            if (pdl_wait_loads) gdc_wait();   // only thread 0 touches global memory (TMA loads and stores)
        }
        __syncthreads();
    }
    // thread 0: tensor load of tile `tile` of signal `sig` into stage `stage`; tiles before tile 0 come from
    // the halo map (sharded signals) or are zero-filled by the TMA unit
    __device__ __forceinline__ void issue_load(int tile, int sig, int stage) const
    {
        const uint32_t bar = bars + 8u * stage;
        mbar_arrive_expect_tx(bar, tbv());
        if (tile < 0 && has_halo)
            tma_load_3d(ring + (uint32_t)stage * tbv(), halo_map, bar, 0, (tile + H) * rowsv(), 0, load_hint);
        else
            tma_load_3d(ring + (uint32_t)stage * tbv(), in_map, bar, 0, tile * rowsv() + row_base, sig, load_hint);
    }
    __device__ __forceinline__ void prologue(int first, int ntl, int sig) const
    {
        if (threadIdx.x == 0) {
            int s2 = st;
            for (int j = 0; j < P && j < ntl; ++j) {
                issue_load(first + j, sig, s2);
                s2 = (s2 + 1 == S) ? 0 : s2 + 1;
            }
        }
    }
    // all threads: wait for the current tile; returns its shared address
    __device__ __forceinline__ uint32_t wait_tile() const
    {
        mbar_wait(bars + 8u * st, (it / (uint32_t)S) & 1u);
        return ring + (uint32_t)st * tbv();
    }
    __device__ __forceinline__ void before_sync()
    {
        if (!INPLACE_ && threadIdx.x == 0 && st_inflight) {  // the staging buffer about to be rewritten is free again
            tma_wait_read0();
            st_inflight = false;
        }
    }
    // thread 0, right after the tile's __syncthreads: refill the stage nobody reads any more, issue the
    // store of the previous output tile
    __device__ __forceinline__ void after_sync(int j, int ntl, int first, int sig)
    {
        if (threadIdx.x == 0) {
            if constexpr (INPLACE_) {
                flush_store();                   // the previous tile's results leave from their ring stage
                if (j + P < ntl) {
                    int s2 = st + P;
                    if (s2 >= S) s2 -= S;
                    // stage s2 held tile j + P - S <= j - 2, whose store was committed at least one group before the
                    // one just issued
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    issue_load(first + j + P, sig, s2);
                }
            } else {
                if (j + P < ntl) {
                    int s2 = st + P;
                    if (s2 >= S) s2 -= S;
                    issue_load(first + j + P, sig, s2);
                }
                flush_store();
            }
        }
    }
    __device__ __forceinline__ void flush_store()
    {
        if (st_pending) {
            if (gdc_pending) {   // first store of this CTA: the previous grid may still be writing the same buffer
                gdc_wait();
                gdc_pending = false;
            }
            if (store_hint) tma_store_3d_hint(out_map, st_buf, 0, st_tile * rowsv(), st_sig, store_hint);
            else tma_store_3d(out_map, st_buf, 0, st_tile * rowsv(), st_sig);
            tma_commit();
            st_pending = false;
            st_inflight = true;
        }
    }
    __device__ __forceinline__ uint32_t out_tile() const
    {
        return INPLACE_ ? ring + (uint32_t)st * tbv() : outb + (otiles & 1u) * tbv();
    }
    // all threads, after writing their part of the staging tile
    __device__ __forceinline__ void staged(int tile, int sig)
    {
        fence_proxy_async_smem();
        if (threadIdx.x == 0) {
            st_pending = true;
            st_tile = tile;
            st_sig = sig;
            st_buf = out_tile();
        }
        ++otiles;
    }
    __device__ __forceinline__ void advance()
    {
        ++it;
        st = (st + 1 == S) ? 0 : st + 1;
        slot = (slot + 1 == GS) ? 0 : slot + 1;
    }
    // everyone is done reading the ring and writing the staging tile of this chunk
    __device__ __forceinline__ void epilogue()
    {
        __syncthreads();
        if (threadIdx.x == 0) {
            flush_store();
            if (INPLACE_) tma_wait_read0();     // the next chunk's prologue refills stages without further checks
        }
    }
    __device__ __forceinline__ void finish() const
    {
        if (threadIdx.x == 0) tma_wait_all0();
    }
    // shared address of byte offset `off` relative to the current tile (negative = history tiles)
    __device__ __forceinline__ uint32_t rel(int off) const
    {
        int o = (int)((uint32_t)st * tbv()) + off;
        if (o < 0) o += (int)ring_bytes;
        return ring + (uint32_t)o;
    }
};

// decodes chunk -> (signal, tile range); false when the range is empty
__device__ __forceinline__ bool chunk_range(const StreamParams& p, int chunk, int& sig, int& t0, int& t1)
{
    sig = chunk / p.chunks_per_signal;
    t0 = (chunk - sig * p.chunks_per_signal) * p.chunk_tiles;
    t1 = t0 + p.chunk_tiles;
    if (t1 > p.tiles_per_signal) t1 = p.tiles_per_signal;
    return t0 < t1;
}

// Cancellation-free window sums for compile-time K <= 8 (MODE 2): v[] holds samples
// a-K+1 .. a+R-1; windows are assembled from power-of-two partial windows, additions only,
// so the relative error stays ~K*2^-24 even where the window sum is nearly zero.
template <int K, int R, int C>
__device__ __forceinline__ void small_window_sums(const float (&v)[R + 7 * C], float (&w)[R])
{
    // v[0] is flat sample a-(K-1)*C; the window of output r is v[r + j*C], j < K (stride C = one channel)
    float w2[R + 6 * C], w4[R + 4 * C], w8[R];
    if constexpr (K >= 2) {
#pragma unroll
        for (int i = 0; i < R + 6 * C; ++i) w2[i] = v[i] + v[i + C];
    }
    if constexpr (K >= 4) {
#pragma unroll
        for (int i = 0; i < R + 4 * C; ++i) w4[i] = w2[i] + w2[i + 2 * C];
    }
    if constexpr (K >= 8) {
#pragma unroll
        for (int i = 0; i < R; ++i) w8[i] = w4[i] + w4[i + 4 * C];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float acc;
        int off = r;
        if constexpr (K >= 8) { acc = w8[off]; off += 8 * C; }
        else if constexpr (K >= 4) { acc = w4[off]; off += 4 * C; }
        else if constexpr (K >= 2) { acc = w2[off]; off += 2 * C; }
        else { acc = v[off]; off += C; }
        if constexpr (K < 8 && K >= 4 && (K & 2)) { acc += w2[off]; off += 2 * C; }
        if constexpr (K >= 2 && (K & 1)) { acc += v[off]; off += C; }
        w[r] = acc;
    }
}

// MODE 0: direct group sums (9 <= k <= direct_max), MODE 1: tile-rebased prefix scan,
// MODE 2: compile-time K <= 8, additions only (MIS unused).
// C = channels interleaved in the flat sample stream (1 mono/planar, 2 stereo, 4): window stride C,
// lag distance k*C, one running sum per channel per thread.
// RMS = moving root-mean-square instead of the moving average (mavg_op): samples are
// squared as they leave shared memory, the window mean goes through a square root on its way to the staging tile;
// everything between (window sums, tile-local rebasing, slide) is the same code.
template <int NT, int R, int MIS, int MODE, int K, int C = 1, bool RMS = false>
__global__ void __launch_bounds__(NT)
    stream_f32_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                      const __grid_constant__ CUtensorMap halo_map, const StreamParams p)
{
    constexpr int T = NT * R;        // samples per tile
    constexpr uint32_t TB = T * 4;   // bytes per tile
    constexpr int ROWS = T / 32;     // 128-byte rows per tile
    constexpr int NW = NT / 32;
    constexpr int CH_OWN = R / 4;
    constexpr int PRE = (7 * C + 3) / 4;  // MODE 2: 16-byte chunks in front of the run that hold (K-1)*C samples
    constexpr int CH_LAG = (MODE == 2) ? PRE : CH_OWN + (MIS ? 1 : 0);
    static_assert(MODE != 2 || (K >= 1 && K <= 8), "MODE 2 serves K in 1..8");
    static_assert(R % C == 0 && MIS % C == 0, "runs hold whole frames");
    static_assert(NT % 32 == 0 && NW <= 16, "NT must be a multiple of 32, at most 512");
    static_assert(R == 16 || R == 32, "run length");

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    TileRing<TB, ROWS> tr;
    const uint32_t gsum = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &halo_map);   // float [GS][NT][C]
    const int H = tr.H, GS = tr.GS;
    const uint32_t wexc = gsum + (uint32_t)GS * NT * C * 4;   // float [GS][32][C]
    const uint32_t wraw = wexc + (uint32_t)GS * 32 * C * 4;   // float [2][32][C]
    tr.init_barriers(wraw + 2u * 32 * C * 4);                 // u64   [S]

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = tr.wait_tile();
            const int slot = tr.slot;
            const uint32_t it = tr.it;

            // ---- own run: R consecutive samples, conflict-free swizzled LDS.128
            float x[R];
#pragma unroll
            for (int c = 0; c < CH_OWN; ++c) {
                const float4 v = lds128(swz(cur + (uint32_t)tid * (R * 4) + 16u * c));
                x[4 * c + 0] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            if constexpr (RMS) {
#pragma unroll
                for (int i = 0; i < R; ++i) x[i] *= x[i];
            }
            // per-channel group totals, fixed order
            float gtot[C], incl[C];
#pragma unroll
            for (int c = 0; c < C; ++c) gtot[c] = 0.f;
            if constexpr (MODE != 2) {
                if constexpr (C == 1) {  // pairwise
                    float q4[CH_OWN];
#pragma unroll
                    for (int c = 0; c < CH_OWN; ++c) q4[c] = (x[4 * c] + x[4 * c + 1]) + (x[4 * c + 2] + x[4 * c + 3]);
                    gtot[0] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
                    if constexpr (R == 32) gtot[0] += (q4[CH_OWN - 4] + q4[CH_OWN - 3]) + (q4[CH_OWN - 2] + q4[CH_OWN - 1]);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) gtot[r % C] += x[r];
                }
            }
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] = gtot[c];  // MODE 1: inclusive prefix of group totals inside the warp

            if constexpr (MODE == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) sts32(gsum + (((uint32_t)slot * NT + tid) * C + c) * 4u, gtot[c]);
            } else if constexpr (MODE == 1) {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float up = __shfl_up_sync(0xffffffffu, incl[c], d);
                        if (lane >= d) incl[c] += up;
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    sts32(gsum + (((uint32_t)slot * NT + tid) * C + c) * 4u, incl[c]);
                    if (lane == 31) sts32(wraw + (((it & 1u) * 32u + warp) * C + c) * 4u, incl[c]);
                }
            }

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);

            float own_off[C], wex[C];
#pragma unroll
            for (int c = 0; c < C; ++c) own_off[c] = wex[c] = 0.f;
            if constexpr (MODE == 1) {
                // exclusive offsets of the NW warp totals of this tile (every warp redundantly)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float v = (lane < NW) ? lds32(wraw + (((it & 1u) * 32u + lane) * C + c) * 4u) : 0.f;
                    float wi = v;
#pragma unroll
                    for (int d = 1; d < NW; d <<= 1) {
                        const float up = __shfl_up_sync(0xffffffffu, wi, d);
                        if (lane >= d) wi += up;
                    }
                    wex[c] = wi - v;
                    if (warp == 0) {
                        if (lane < NW) sts32(wexc + (((uint32_t)slot * 32u + lane) * C + c) * 4u, wex[c]);
                        if (lane == NW - 1) sts32(wexc + (((uint32_t)slot * 32u + 31u) * C + c) * 4u, wi);  // tile total
                    }
                    own_off[c] = __shfl_sync(0xffffffffu, wex[c], warp);
                }
            }

            if (is_out) {
                // ---- lag run: samples [a-k, a-k+R) (a = run start), loaded as aligned 16-byte chunks
                float xl[CH_LAG * 4];
                {
                    const int back = (MODE == 2) ? PRE : (int)p.lag_chunks;
                    const int lin = (tid * CH_OWN - back) * 16;
#pragma unroll
                    for (int c = 0; c < CH_LAG; ++c) {
                        const float4 v = lds128(swz(tr.rel(lin + 16 * c)));
                        xl[4 * c + 0] = v.x; xl[4 * c + 1] = v.y; xl[4 * c + 2] = v.z; xl[4 * c + 3] = v.w;
                    }
                    if constexpr (RMS) {
#pragma unroll
                        for (int i = 0; i < CH_LAG * 4; ++i) xl[i] *= xl[i];
                    }
                }

                const float inv = p.inv_k;
                const uint32_t ob = tr.out_tile() + (uint32_t)tid * (R * 4);
                if constexpr (MODE == 2) {
                    // ---- additions only: v = samples a-K+1 .. a+R-1
                    float v[R + 7 * C], w[R];
#pragma unroll
                    for (int i = 0; i < R + 7 * C; ++i) v[i] = 0.f;
#pragma unroll
                    for (int i = 0; i < (K - 1) * C; ++i) v[i] = xl[4 * PRE - (K - 1) * C + i];
#pragma unroll
                    for (int i = 0; i < R; ++i) v[(K - 1) * C + i] = x[i];
                    small_window_sums<K, R, C>(v, w);
#pragma unroll
                    for (int c = 0; c < CH_OWN; ++c) {
                        if constexpr (RMS)
                            sts128(swz(ob + 16u * c), sqrtf(w[4 * c] * inv), sqrtf(w[4 * c + 1] * inv),
                                   sqrtf(w[4 * c + 2] * inv), sqrtf(w[4 * c + 3] * inv));   // sums of squares: never negative
                        else
                            sts128(swz(ob + 16u * c), w[4 * c] * inv, w[4 * c + 1] * inv, w[4 * c + 2] * inv,
                                   w[4 * c + 3] * inv);
                    }
                } else {
                    // ---- per-channel window sum over [a-k*C, a): whole groups between, then the tail of the lag group
                    float acc[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = 0.f;
                    if constexpr (MODE == 0) {
                        // the n_full groups in front of the own one, nearest first; the summary ring wraps at
                        // most once, so the walk is two straight runs without a per-step wrap test
                        const int gi = slot * NT + tid;
                        const int n1 = ((int)p.n_full < gi) ? (int)p.n_full : gi;
                        uint32_t ga = gsum + (uint32_t)gi * (C * 4u);
                        for (int n = 0; n < n1; ++n) {
                            ga -= C * 4u;
#pragma unroll
                            for (int c = 0; c < C; ++c) acc[c] += lds32(ga + 4u * c);
                        }
                        ga = gsum + (uint32_t)(GS * NT) * (C * 4u);
                        for (int n = n1; n < (int)p.n_full; ++n) {
                            ga -= C * 4u;
#pragma unroll
                            for (int c = 0; c < C; ++c) acc[c] += lds32(ga + 4u * c);
                        }
                    } else {
                        int lt = tid - (int)(p.n_full + 1u);  // thread index of the lag group, relative to this tile
                        int h = 0;
                        if (lt < 0) {
                            h = (-lt + NT - 1) / NT;
                            lt += h * NT;
                        }
                        int ls = slot - h;
                        if (ls < 0) ls += GS;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float wsame = __shfl_sync(0xffffffffu, wex[c], lt >> 5);
                            const float wold = lds32(wexc + (((uint32_t)ls * 32u + (uint32_t)(lt >> 5)) * C + c) * 4u);
                            const float cp_lag =
                                lds32(gsum + (((uint32_t)ls * NT + lt) * C + c) * 4u) + (h == 0 ? wsame : wold);
                            const float e_own = own_off[c] + (incl[c] - gtot[c]);
                            if (h == 0) {
                                acc[c] = e_own - cp_lag;
                            } else {
                                float rest = lds32(wexc + (((uint32_t)ls * 32u + 31u) * C + c) * 4u) - cp_lag;  // lag-tile tail
                                int ms = ls;
                                for (int v2 = 1; v2 < h; ++v2) {  // whole tiles strictly between (k*C > tile only)
                                    ms = (ms + 1 == GS) ? 0 : ms + 1;
                                    rest += lds32(wexc + (((uint32_t)ms * 32u + 31u) * C + c) * 4u);
                                }
                                acc[c] = e_own + rest;
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if ((uint32_t)r < p.m_part) acc[r % C] += xl[MIS + r];

                    // ---- slide and scale (acc[c] is the running window sum of channel c)
#pragma unroll
                    for (int c4 = 0; c4 < CH_OWN; ++c4) {
                        float y[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int r = 4 * c4 + q;
                            acc[r % C] += x[r] - xl[MIS + r];
                            y[q] = acc[r % C] * inv;
                            if constexpr (RMS) y[q] = sqrtf(fmaxf(y[q], 0.f));   // a slid sum of squares may round below 0
                        }
                        sts128(swz(ob + 16u * c4), y[0], y[1], y[2], y[3]);
                    }
                }
                tr.staged(tile, sig);
            }
            tr.advance();
        }

        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// int16 streaming kernel -- the reference's own sample format (wav_header.h:26-48), mono or
// interleaved stereo.  Same skeleton as stream_f32_kernel (persistent CTAs, TMA ring, one
// __syncthreads per tile, TMA store); arithmetic is exact: int32 window sums (|w| <= k * 32768,
// k <= 32768, everything modulo 2^32) and C truncating division by multiply-high, so results are
// bit-identical to profilable_cpu_computations (basics/profilable_moving_averager.cpp:14-37),
// unlike the reference GPU kernels that multiply by a float reciprocal.  Interleaved channels are
// handled on the flat sample stream: window stride C, lag distance L = k*C, one running sum per
// channel per thread.  HBM traffic: 2 B read + 2 B written per sample.
//
// One arithmetic for every window (round 2; rounds 0-1 had a direct-sum mode for short windows and a prefix-scan
// mode with a lag look-up for long ones, 13.3 / 16.0 executed instructions per sample).  A thread owns a run of R
// consecutive flat samples and keeps, per channel, the window sum at the run start:
//     start(t) = W + sum_{u < t} d(u),      d(u) = (total of u's run) - (total of u's lag run, the R samples L earlier)
// where W = window sum at the first sample of the tile, carried in registers from tile to tile (W += sum_t d(t)).
// The slide itself yields everything: s[r] = sum_{i <= r} (x[i] - x[i - L]) is computed BEFORE the tile's barrier with
// two dp2a per sample on the packed words and kept in registers, d = s[R-1]; one exclusive scan of d over the CTA
// (warp shuffles + 16 warp totals through shared memory) gives start(t); the outputs are
//     y[r] = trunc((start + s[r]) / k) = (t >> sh) + (t >>> 31),  t = mulhi(start + s[r], M)
// (a 64-bit multiply-add hi32(s[r] * M + start * M) would fold the addition away, but ptxas splits mad.wide into
// IMAD.WIDE + IADD3 + IMAD.X -- three instructions instead of two).  No run totals, no head-of-lag sums, no summary
// ring, no lag look-up: the same instruction count for any k.  A chunk obtains its first W from the H = ceil(L / tile) tiles in front of it:
// their samples at or behind (chunk start - L) are summed with the same scan (the first of them masked).
// ----------------------------------------------------------------------------------
__host__ __device__ inline uint32_t stream_i16_smem_bytes(int NT, int R, int S, int H, int C)
{
    (void)H;
    const uint32_t TB = (uint32_t)NT * R * 2;
    const uint32_t NW = (uint32_t)NT / 32, G = 32 / NW;          // warps; channel groups of the cross-warp scan
    return 1024u + (uint32_t)S * TB + 2u * TB + 2u * 32 * ((C + G - 1) / G) * 4 + (uint32_t)S * 8;
}

// TMA swizzle (bytes: 0 / 32 / 64 / 128) under which the 128-bit shared-memory accesses of a warp whose threads own
// consecutive runs of R int16 samples (R / 8 chunks of 16 bytes) are free of bank conflicts
__host__ __device__ constexpr int i16_swizzle_bytes(int R)
{
    const int chunks = R / 8;
    return chunks % 2 == 1 ? 0 : 128;   // (strides of 6 / 12 chunks would need the 32- / 64-byte modes, whose boxes are narrower than a row)
}
// offset inside a 1024-byte aligned tile under the TMA swizzle of SWZ bytes: address bits [4, 4 + log2(SWZ / 16)) are
// XORed with the bits starting at 7
template <int SWZ>
__device__ __forceinline__ int pre_swz_n(int x)
{
    if constexpr (SWZ == 0) return x;
    else return x ^ ((x >> 3) & (SWZ == 128 ? 0x70 : SWZ == 64 ? 0x30 : 0x10));
}

__device__ __forceinline__ uint4 lds128u(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int lds32i(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32i(uint32_t addr, int v)
{
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32u(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int2 lds64i(uint32_t addr)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64i(uint32_t addr, int a, int b)
{
    asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// dp2a on the PACKED sample words -- d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1, signed -- with byte weights that
// pick a half (+w), drop it (0) or subtract it (-w): the int16 kernels never unpack a sample.
__device__ __forceinline__ int dp2a_s(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// trunc(w / k) for |w| <= 32768 k (times the weight scale), k <= 32768: t = mulhi(w, M); (t >> s) + (t >>> 31)
// (exactness: plan_stream_i16 in mavg.cu)
__device__ __forceinline__ uint32_t div_trunc_mulhi(int w, int mul, uint32_t sh)
{
    const int t = __mulhi(w, mul);
    return (uint32_t)((t >> sh) + (int)((uint32_t)t >> 31));
}

template <int NT, int R, int C, int MIS>
__global__ void __launch_bounds__(NT, (NT * R * 2 <= 18432 ? 2 : 1))   // tiles of 16 / 18 KB: two CTAs per SM
    stream_i16_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                      const __grid_constant__ CUtensorMap halo_map, const StreamParams p)
{
    // R int16 samples per thread run (32 = 64 bytes)
    constexpr int T = NT * R;
    constexpr uint32_t TB = T * 2;
    constexpr int ROWS = T / 64;     // 128-byte rows per tile
    constexpr int NW = NT / 32;
    constexpr int CH_OWN = R / 8;
    constexpr int CH_LAG = CH_OWN + (MIS ? 1 : 0);
    // 64-byte runs (R = 32) need the 128-byte XOR swizzle for conflict-free 128-bit accesses; runs of an odd number of
    // 16-byte chunks (R = 24, 40: 3 / 5 / 6 interleaved channels) are conflict-free in the dense layout and collide in
    // the swizzled one, so their tensor maps are encoded without swizzle
    // (brute-forced over every quarter-warp: strides of 3 / 5 / 9 chunks: dense; 4 / 8: 128-byte swizzle; 6: 2-way
    // conflicts either way, which is why 3 / 6 channels run as 224 threads x 72 samples and not 256 x 48)
    constexpr int SWZ = i16_swizzle_bytes(R);
    static_assert(R % C == 0 && R % 8 == 0 && NW <= 16 && MIS >= 0 && MIS < 8 && T % 64 == 0 && ROWS <= 256, "shape");

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    TileRing<TB, ROWS> tr;
    // Cross-warp scan of the warp totals, once per warp with the lanes working in parallel: lane = (warp w' = lane % NW,
    // channel group g = lane / NW) owns the totals of warp w' for channels g, g + G, g + 2 G ...; the array is laid out
    // [parity][channel][NW] so that the j-th load of a warp is one dense row of 32 words
    constexpr int G = 32 / NW;                // channel groups
    constexpr int CJ = (C + G - 1) / G;       // channels per lane
    const uint32_t wraw = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &halo_map);   // uint32 [2][CJ * G][NW]
    const int H = tr.H;
    tr.init_barriers(wraw + 2u * 32 * CJ * 4);
    const int wq = lane % NW;                 // the warp whose totals this lane scans
    const int wlast = lane - wq + NW - 1;     // the lane of this group that ends up with the inclusive total
    // dp2a byte weights (+w / -w on the low or the high half; w = 2 when k == 2, see plan_stream_i16)
    const uint32_t w_lo = p.wscale, w_hi = p.wscale << 8;
    const uint32_t n_lo = (0u - p.wscale) & 0xffu, n_hi = n_lo << 8;
    const int L = (int)(p.k * (uint32_t)C);              // lag distance in flat samples
    int xo[CH_OWN], xg[CH_LAG];                          // pre-swizzled chunk offsets: own run, lag run
#pragma unroll
    for (int c = 0; c < CH_OWN; ++c) xo[c] = pre_swz_n<SWZ>(tid * (R * 2) + 16 * c);
#pragma unroll
    for (int c = 0; c < CH_LAG; ++c) xg[c] = pre_swz_n<SWZ>((tid * CH_OWN - (int)p.lag_chunks + c) * 16);
    const int mul = (int)p.div_mul;
    const uint32_t sh = p.div_shift;

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);
        uint32_t Wl[CJ];                                 // window sum at the first sample of the tile, channels g + j G,
#pragma unroll                                           // kept by the lanes that scan warp 0 (wq == 0)
        for (int j = 0; j < CJ; ++j) Wl[j] = 0u;

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);                // CTA-uniform
            const uint32_t cur = tr.wait_tile();
            const uint32_t it = tr.it;

            uint32_t xw[R / 2];                          // own run, packed
#pragma unroll
            for (int c = 0; c < CH_OWN; ++c) {
                const uint4 v = lds128u(cur + (uint32_t)xo[c]);
                xw[4 * c] = v.x, xw[4 * c + 1] = v.y, xw[4 * c + 2] = v.z, xw[4 * c + 3] = v.w;
            }
            uint32_t s[R];                               // s[r] = sum_{i <= r, same channel} (x[i] - x[i - L])
            uint32_t d[C];                               // this run's contribution to the carried window sums
            if (is_out) {
                uint32_t xlw[CH_LAG * 4];                // lag run, packed, as aligned 16-byte chunks
                const uint32_t b0 = cur, b1 = cur + tr.ring_bytes;
                const int neg_st = (int)tr.ring - (int)cur;
#pragma unroll
                for (int c = 0; c < CH_LAG; ++c) {
                    const uint4 v = lds128u(ring_addr(xg[c], b0, b1, neg_st));
                    xlw[4 * c] = v.x, xlw[4 * c + 1] = v.y, xlw[4 * c + 2] = v.z, xlw[4 * c + 3] = v.w;
                }
                int a[C];
#pragma unroll
                for (int c = 0; c < C; ++c) a[c] = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int e = MIS + r;               // the lag partner of run element r in the aligned lag words
                    int a2 = dp2a_s(xw[r >> 1], (r & 1) ? w_hi : w_lo, a[r % C]);
                    a2 = dp2a_s(xlw[e >> 1], (e & 1) ? n_hi : n_lo, a2);
                    a[r % C] = a2;
                    s[r] = (uint32_t)a2;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = (uint32_t)a[c];
            } else {
                // warm-up tile in front of the chunk: samples at or behind (chunk start - L) belong to the first
                // output's window; `rel` leading elements of this run lie in front of it (only in the first warm-up tile)
                const int rel = (t0 - tile) * T - L - tid * R;
                int a[C];
#pragma unroll
                for (int c = 0; c < C; ++c) a[c] = 0;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const uint32_t m0 = (2 * q >= rel) ? w_lo : 0u, m1 = (2 * q + 1 >= rel) ? w_hi : 0u;
                    if constexpr (C == 1) {
                        a[0] = dp2a_s(xw[q], m0 | m1, a[0]);
                    } else {
                        a[(2 * q) % C] = dp2a_s(xw[q], m0, a[(2 * q) % C]);
                        a[(2 * q + 1) % C] = dp2a_s(xw[q], m1, a[(2 * q + 1) % C]);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = (uint32_t)a[c];
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] = 0u;
            }

            // ---- exclusive scan of d over the CTA: inside the warp by shuffles, across warps through shared memory
            uint32_t incl[C];
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] = d[c];
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl[c], dd);
                    if (lane >= dd) incl[c] += up;
                }
            }
            const uint32_t wbase = wraw + (it & 1u) * (32u * CJ * 4u);
            if (lane == 31) {
#pragma unroll
                for (int c = 0; c < C; ++c) sts32u(wbase + (uint32_t)((c / G) * 32 + (c % G) * NW + warp) * 4u, incl[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] -= d[c];   // exclusive inside the warp

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);

            uint32_t excl[CJ];                           // W + totals of the warps in front of warp wq
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const uint32_t v = lds32u(wbase + (uint32_t)(jj * 32 + lane) * 4u);
                uint32_t wi = v + (wq == 0 ? Wl[jj] : 0u);
#pragma unroll
                for (int dd = 1; dd < NW; dd <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, wi, dd);
                    if (wq >= dd) wi += up;
                }
                Wl[jj] = __shfl_sync(0xffffffffu, wi, wlast);   // next tile's W (used where wq == 0)
                excl[jj] = wi - v;
            }
            uint32_t start[C];
#pragma unroll
            for (int c = 0; c < C; ++c)
                start[c] = __shfl_sync(0xffffffffu, excl[c / G], (c % G) * NW + warp) + incl[c];

            if (is_out) {
                const uint32_t ob = tr.out_tile();
#pragma unroll
                for (int c = 0; c < CH_OWN; ++c) {
                    uint32_t wds[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t y[2];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int r = 8 * c + 2 * q + hh;
                            y[hh] = div_trunc_mulhi((int)(start[r % C] + s[r]), mul, sh);
                        }
                        wds[q] = __byte_perm(y[0], y[1], 0x5410);
                    }
                    sts128u(ob + (uint32_t)xo[c], wds[0], wds[1], wds[2], wds[3]);
                }
                tr.staged(tile, sig);
            }
            tr.advance();
        }

        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// Far-lag kernel -- float32 mono / planar / interleaved stereo, windows longer than the shared-memory history
// (k >~ 49 000 frames mono, 24 000 stereo).
// The lag samples x[i-k] can no longer wait in the CTA's ring, so they come back through a SECOND TMA stream: for
// output tile j the 257 rows that hold samples [jT-k, jT-k+T) are loaded as two 129-row boxes (a TMA box holds at
// most 256 rows) into a 2-stage lag ring.  The same CTA read those bytes as its own tile k samples earlier
// (148 CTAs x k x 4 B = 35 MB at k = 60 000), so with a normal L2 policy they are still in the 126 MB L2 and DRAM
// traffic stays near 8 B/sample.
// Window sums: W = sum of the k samples in front of the tile is carried per CTA in fp64,
//     W(j+1) = W(j) + sum_t d[t],   d[t] = (total of thread t's own run) - (total of its lag run),
// and a thread's run starts from W + (exclusive scan of d over the tile's 512 threads), then slides in fp32.
// A chunk builds its first W by streaming the ceil(k/T) tiles in front of it through the same loop (own totals only,
// the first of them masked to the window).  Rounding: every fp32 quantity is a difference of 16-sample sums, the
// carried W is fp64, so the relative error of y stays ~1e-7 for any k and any signal length.
// Shards: the left context must be contiguous with the input (row_base rows in front of it), which is how
// mavg_run_host slices and shard plans fed through mavg_run_host present it.
// ----------------------------------------------------------------------------------
struct FarParams {
    StreamParams sp;        // k, inv_k, chunking; ring: stages = 1 + prefetch, hist_tiles = 0
    int32_t warm_tiles;     // HT = ceil(k / T)
    int32_t row_base;       // rows of left context in front of row 0
    int32_t lag_rows;       // (k + koff) / 32: rows from the lag box's first row to the tile's first row
    uint32_t koff;          // (32 - k % 32) % 32: first lag sample inside its 128-byte row
    int32_t lag_prefetch;   // lag boxes in flight: 1 or 2
    int32_t lag_stages;     // stages of the lag ring: 1 or 2 (>= lag_prefetch)
    int32_t hints;          // L2 policies: 0 everything evict-normal (round 1); 1 lag boxes and output evict-first (a lag box
                            // is dead once read, the output is never read back: neither should push the own tiles that
                            // come back as lag boxes out of L2); 2 = 1 + own tiles evict-last
};
// The lag samples of a tile of ROWS 128-byte rows span ROWS + 1 rows: one TMA box, or two boxes of ROWS / 2 + 1 rows
// when that exceeds the 256 rows a box can hold; each box is padded to whole 1024-byte swizzle atoms.
__host__ __device__ constexpr int far_lag_boxes(int rows) { return rows + 1 > 256 ? 2 : 1; }
__host__ __device__ constexpr uint32_t far_lag_box_bytes(int rows)
{
    return ((uint32_t)(rows / far_lag_boxes(rows) + 1) * 128u + 1023u) & ~1023u;
}

// Results are written over the own tile (the kernel reads nothing but the thread's own run from it) and stored from
// the ring stage: S = P + 2 own stages and no staging tiles.
__host__ __device__ inline uint32_t far_smem_bytes(int NT, int R, int S, int SL)
{
    const uint32_t TB = (uint32_t)NT * R * 4;
    const int rows = NT * R / 32;
    return 1024u + (uint32_t)S * TB + (uint32_t)SL * far_lag_boxes(rows) * far_lag_box_bytes(rows) + 2u * 32 * 2 * 4 +
           (uint32_t)SL * 8 + (uint32_t)S * 8 + 64;
}

template <int NT, int R, int MIS, int C = 1>
__global__ void __launch_bounds__(NT)
    stream_far_f32_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                          const __grid_constant__ CUtensorMap lag_map, const FarParams fp)
{
    // C = 2: interleaved stereo; the flat stream is filtered with lag distance L = k * C (sp.k holds L), one
    // carried sum / scan / running sum per channel, channel of run element r = r % C (R % C == 0, MIS % C == 0)
    const StreamParams& p = fp.sp;
    constexpr int T = NT * R;
    constexpr uint32_t TB = T * 4;
    constexpr int ROWS = T / 32;
    constexpr int NW = NT / 32;
    constexpr int CH_OWN = R / 4;
    constexpr int CH_LAG = CH_OWN + (MIS ? 1 : 0);
    constexpr int NBOX = far_lag_boxes(ROWS);                 // lag boxes per tile
    constexpr int BOXROWS = ROWS / NBOX + 1;
    constexpr uint32_t LAGBOX = far_lag_box_bytes(ROWS);
    constexpr int TPB = NT / NBOX;                            // threads whose lag runs live in one box
    const int SL = fp.lag_stages;
    static_assert(R == 16 && MIS >= 0 && MIS < 4 && ROWS % NBOX == 0 && NT % NBOX == 0 && ROWS <= 256, "shape");
    static_assert((C == 1 || C == 2) && MIS % C == 0, "mono or interleaved stereo");

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    TileRing<TB, ROWS, true> tr;
    const uint32_t lagbuf = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &in_map);   // 1024-aligned: ring of whole tiles
    tr.row_base = fp.row_base;
    tr.load_hint = fp.hints == 2 ? kEvictLast : kEvictNormal;
    tr.store_hint = fp.hints ? kEvictFirst : 0;
    const uint64_t lag_hint = fp.hints ? kEvictFirst : kEvictNormal;
    const uint32_t wraw = lagbuf + (uint32_t)SL * NBOX * LAGBOX;      // float [2][32][2]
    const uint32_t lbars = wraw + 2u * 32 * 2 * 4;                    // u64 [SL]
    if (tid == 0) {
        prefetch_tmap(&lag_map);
        for (int s2 = 0; s2 < SL; ++s2) mbar_init(lbars + 8u * s2, 1);
    }
    tr.init_barriers(lbars + (uint32_t)SL * 8);                       // fences the inits above as well, then syncs

    int xo[CH_OWN], xg[CH_LAG];   // pre-swizzled chunk offsets: own run (tile / staging), lag run (lag stage)
#pragma unroll
    for (int c = 0; c < CH_OWN; ++c) xo[c] = pre_swz(tid * (R * 4) + 16 * c);
    {
        const int half = tid / TPB;
        const int c0 = (int)(fp.koff >> 2) + CH_OWN * (tid % TPB);
#pragma unroll
        for (int c = 0; c < CH_LAG; ++c) xg[c] = pre_swz((c0 + c) * 16) + half * (int)LAGBOX;
    }
    const int PL = fp.lag_prefetch;
    auto issue_lag = [&](int tile, int sig, int stage) {              // thread 0
        const uint32_t bar = lbars + 8u * (uint32_t)stage;
        mbar_arrive_expect_tx(bar, (uint32_t)NBOX * BOXROWS * 128u);
        const int r0 = tile * ROWS - fp.lag_rows + fp.row_base;
        const uint32_t dst = lagbuf + (uint32_t)stage * NBOX * LAGBOX;
        tma_load_3d(dst, &lag_map, bar, 0, r0, sig, lag_hint);
        if constexpr (NBOX == 2) tma_load_3d(dst + LAGBOX, &lag_map, bar, 0, r0 + ROWS / 2, sig, lag_hint);
    };
    uint32_t lagit = 0;   // lag boxes consumed so far by this CTA, never reset
    int lst = 0;          // lagit % SL

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int HT = fp.warm_tiles;
        const int first = t0 - HT;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);
        if (tid == 0) {
            int s2 = lst;
            for (int i = 0; i < PL && t0 + i < t1; ++i) {
                issue_lag(t0 + i, sig, s2);
                s2 = (s2 + 1 == SL) ? 0 : s2 + 1;
            }
        }
        double W[C];      // per channel: sum of the k frames in front of the current tile (complete after the warm-up)
#pragma unroll
        for (int c = 0; c < C; ++c) W[c] = 0.0;

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= HT);
            const uint32_t cur = tr.wait_tile();
            const uint32_t it = tr.it;

            float x[R];
#pragma unroll
            for (int c = 0; c < CH_OWN; ++c) {
                const float4 v = lds128(cur + (uint32_t)xo[c]);
                x[4 * c + 0] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            float xl[CH_LAG * 4];
            float d[C];
            if (is_out) {
                mbar_wait(lbars + 8u * (uint32_t)lst, (lagit / (uint32_t)SL) & 1u);
                const uint32_t lb = lagbuf + (uint32_t)lst * NBOX * LAGBOX;
#pragma unroll
                for (int c = 0; c < CH_LAG; ++c) {
                    const float4 v = lds128(lb + (uint32_t)xg[c]);
                    xl[4 * c + 0] = v.x; xl[4 * c + 1] = v.y; xl[4 * c + 2] = v.z; xl[4 * c + 3] = v.w;
                }
                if constexpr (C == 1) {
                    float q4[CH_OWN], l4[CH_OWN];
#pragma unroll
                    for (int c = 0; c < CH_OWN; ++c) {
                        q4[c] = (x[4 * c] + x[4 * c + 1]) + (x[4 * c + 2] + x[4 * c + 3]);
                        l4[c] = (xl[MIS + 4 * c] + xl[MIS + 4 * c + 1]) + (xl[MIS + 4 * c + 2] + xl[MIS + 4 * c + 3]);
                    }
                    d[0] = ((q4[0] + q4[1]) + (q4[2] + q4[3])) - ((l4[0] + l4[1]) + (l4[2] + l4[3]));
                } else {
                    float go[C], gl[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) go[c] = gl[c] = 0.f;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        go[r % C] += x[r];
                        gl[r % C] += xl[MIS + r];
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) d[c] = go[c] - gl[c];
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH_LAG * 4; ++i) xl[i] = 0.f;
                // warm-up: the tile lies in front of the chunk; only its samples inside the first window count
                const int m0 = (j == 0) ? HT * T - (int)p.k : 0;      // local index of the first sample that counts
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) d[r % C] += (tid * R + r >= m0) ? x[r] : 0.f;
            }
            float incl[C];
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] = d[c];
#pragma unroll
            for (int s2 = 1; s2 < 32; s2 <<= 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float up = __shfl_up_sync(0xffffffffu, incl[c], s2);
                    if (lane >= s2) incl[c] += up;
                }
            }
            if (lane == 31) {
#pragma unroll
                for (int c = 0; c < C; ++c) sts32(wraw + (((it & 1u) * 32u + (uint32_t)warp) * 2u + c) * 4u, incl[c]);
            }

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);
            if (tid == 0 && is_out && tile + PL < t1) {
                int s2 = lst + PL;
                if (s2 >= SL) s2 -= SL;
                issue_lag(tile + PL, sig, s2);       // that stage was read into registers before the barrier
            }

            float acc[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float v = (lane < NW) ? lds32(wraw + (((it & 1u) * 32u + (uint32_t)lane) * 2u + c) * 4u) : 0.f;
                float wi = v;
#pragma unroll
                for (int s2 = 1; s2 < NW; s2 <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, wi, s2);
                    if (lane >= s2) wi += up;
                }
                const float dtot = __shfl_sync(0xffffffffu, wi, NW - 1);
                const float own_off = __shfl_sync(0xffffffffu, wi - v, warp);
                acc[c] = (float)(W[c] + (double)(own_off + (incl[c] - d[c])));
                W[c] += (double)dtot;
            }

            if (is_out) {
                const float inv = p.inv_k;
                const uint32_t ob = tr.out_tile();
#pragma unroll
                for (int c = 0; c < CH_OWN; ++c) {
                    float y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = 4 * c + q;
                        acc[r % C] += x[r] - xl[MIS + r];
                        y[q] = acc[r % C] * inv;
                    }
                    sts128(ob + (uint32_t)xo[c], y[0], y[1], y[2], y[3]);
                }
                tr.staged(tile, sig);
                ++lagit;
                lst = (lst + 1 == SL) ? 0 : lst + 1;
            }
            tr.advance();
        }
        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// Far-lag int16 kernel -- stream_i16_kernel's exact delta-scan arithmetic with stream_far_f32_kernel's data movement:
// interleaved int16 with C channels whose lag distance L = k C no longer fits the ring (stereo beyond k = 24 576,
// 8 channels beyond k = 6 144 ...).  The lag run x[i - L .. i - L + R) of every thread comes back through a second TMA
// stream (one or two lag boxes per tile, normally still in L2: the same CTA read those bytes as its own tile L samples
// earlier), the results are written over the own tile and stored from its ring stage (TileRing<..., INPLACE>), the
// window sum W at the first sample of the tile is carried in int32 (|W| <= 32768 k < 2^31 for k <= 46 340, exact, so
// -- unlike the float32 far-lag kernel -- sliced, sharded and whole runs agree bit for bit), a tile range builds its
// first W from the ceil(L / T) masked warm-up tiles in front of it.  fp.sp.k = k (frames); fp.koff / fp.lag_rows
// describe L in int16 samples and 64-sample rows.
// ----------------------------------------------------------------------------------
__host__ __device__ inline uint32_t far_i16_smem_bytes(int NT, int R, int S, int SL, int C)
{
    const uint32_t TB = (uint32_t)NT * R * 2;
    const int rows = NT * R / 64;
    const uint32_t NW = (uint32_t)NT / 32, G = 32 / NW;
    return 2048u + (uint32_t)S * TB + (uint32_t)SL * far_lag_boxes(rows) * far_lag_box_bytes(rows) +
           2u * 32 * ((C + G - 1) / G) * 4 + (uint32_t)SL * 8 + (uint32_t)S * 8 + 64;
}

template <int NT, int R, int C, int MIS>
__global__ void __launch_bounds__(NT)
    stream_far_i16_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                          const __grid_constant__ CUtensorMap lag_map, const FarParams fp)
{
    const StreamParams& p = fp.sp;
    constexpr int T = NT * R;
    constexpr uint32_t TB = T * 2;
    constexpr int ROWS = T / 64;
    constexpr int NW = NT / 32;
    constexpr int CH_OWN = R / 8;
    constexpr int CH_LAG = CH_OWN + (MIS ? 1 : 0);
    constexpr int SWZ = i16_swizzle_bytes(R);
    constexpr int NBOX = far_lag_boxes(ROWS);
    constexpr int BOXROWS = ROWS / NBOX + 1;
    constexpr uint32_t LAGBOX = far_lag_box_bytes(ROWS);
    constexpr int TPB = NT / NBOX;
    constexpr int G = 32 / NW;
    constexpr int CJ = (C + G - 1) / G;
    static_assert(R % C == 0 && R % 8 == 0 && NW <= 16 && MIS >= 0 && MIS < 8 && T % 64 == 0 && ROWS <= 256 &&
                      ROWS % NBOX == 0 && NT % NBOX == 0 && (SWZ == 0 || TB % 1024 == 0),
                  "shape");
    const int SL = fp.lag_stages;

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    TileRing<TB, ROWS, true> tr;
    uint32_t lagbuf = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &in_map);
    lagbuf = (lagbuf + 1023u) & ~1023u;       // (dense shapes: the ring is not a whole number of 1024-byte atoms)
    tr.row_base = fp.row_base;
    tr.load_hint = fp.hints == 2 ? kEvictLast : kEvictNormal;
    tr.store_hint = fp.hints ? kEvictFirst : 0;
    const uint64_t lag_hint = fp.hints ? kEvictFirst : kEvictNormal;
    const uint32_t wraw = lagbuf + (uint32_t)SL * NBOX * LAGBOX;      // uint32 [2][CJ * G][NW] warp totals
    const uint32_t lbars = wraw + 2u * 32 * CJ * 4;                   // u64 [SL]
    if (tid == 0) {
        prefetch_tmap(&lag_map);
        for (int s2 = 0; s2 < SL; ++s2) mbar_init(lbars + 8u * s2, 1);
    }
    tr.init_barriers(lbars + (uint32_t)SL * 8);
    const int wq = lane % NW;
    const int wlast = lane - wq + NW - 1;
    const uint32_t w_lo = p.wscale, w_hi = p.wscale << 8;
    const uint32_t n_lo = (0u - p.wscale) & 0xffu, n_hi = n_lo << 8;
    const int L = (int)(p.k * (uint32_t)C);
    int xo[CH_OWN], xg[CH_LAG];
#pragma unroll
    for (int c = 0; c < CH_OWN; ++c) xo[c] = pre_swz_n<SWZ>(tid * (R * 2) + 16 * c);
    {
        const int half = tid / TPB;
        const int c0 = (int)(fp.koff >> 3) + CH_OWN * (tid % TPB);
#pragma unroll
        for (int c = 0; c < CH_LAG; ++c) xg[c] = pre_swz_n<SWZ>((c0 + c) * 16) + half * (int)LAGBOX;
    }
    const int mul = (int)p.div_mul;
    const uint32_t sh = p.div_shift;
    const int PL = fp.lag_prefetch;
    auto issue_lag = [&](int tile, int sig, int stage) {              // thread 0
        const uint32_t bar = lbars + 8u * (uint32_t)stage;
        mbar_arrive_expect_tx(bar, (uint32_t)NBOX * BOXROWS * 128u);
        const int r0 = tile * ROWS - fp.lag_rows + fp.row_base;
        const uint32_t dst = lagbuf + (uint32_t)stage * NBOX * LAGBOX;
        tma_load_3d(dst, &lag_map, bar, 0, r0, sig, lag_hint);
        if constexpr (NBOX == 2) tma_load_3d(dst + LAGBOX, &lag_map, bar, 0, r0 + ROWS / 2, sig, lag_hint);
    };
    uint32_t lagit = 0;
    int lst = 0;

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int HT = fp.warm_tiles;
        const int first = t0 - HT;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);
        if (tid == 0) {
            int s2 = lst;
            for (int i = 0; i < PL && t0 + i < t1; ++i) {
                issue_lag(t0 + i, sig, s2);
                s2 = (s2 + 1 == SL) ? 0 : s2 + 1;
            }
        }
        uint32_t Wl[CJ];
#pragma unroll
        for (int j = 0; j < CJ; ++j) Wl[j] = 0u;

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= HT);
            const uint32_t cur = tr.wait_tile();
            const uint32_t it = tr.it;

            uint32_t xw[R / 2];
#pragma unroll
            for (int c = 0; c < CH_OWN; ++c) {
                const uint4 v = lds128u(cur + (uint32_t)xo[c]);
                xw[4 * c] = v.x, xw[4 * c + 1] = v.y, xw[4 * c + 2] = v.z, xw[4 * c + 3] = v.w;
            }
            uint32_t s[R];
            uint32_t d[C];
            if (is_out) {
                mbar_wait(lbars + 8u * (uint32_t)lst, (lagit / (uint32_t)SL) & 1u);
                const uint32_t lb = lagbuf + (uint32_t)lst * NBOX * LAGBOX;
                uint32_t xlw[CH_LAG * 4];
#pragma unroll
                for (int c = 0; c < CH_LAG; ++c) {
                    const uint4 v = lds128u(lb + (uint32_t)xg[c]);
                    xlw[4 * c] = v.x, xlw[4 * c + 1] = v.y, xlw[4 * c + 2] = v.z, xlw[4 * c + 3] = v.w;
                }
                int a[C];
#pragma unroll
                for (int c = 0; c < C; ++c) a[c] = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int e = MIS + r;
                    int a2 = dp2a_s(xw[r >> 1], (r & 1) ? w_hi : w_lo, a[r % C]);
                    a2 = dp2a_s(xlw[e >> 1], (e & 1) ? n_hi : n_lo, a2);
                    a[r % C] = a2;
                    s[r] = (uint32_t)a2;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = (uint32_t)a[c];
            } else {
                const int rel = (t0 - tile) * T - L - tid * R;       // leading run elements in front of the first window
                int a[C];
#pragma unroll
                for (int c = 0; c < C; ++c) a[c] = 0;
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const uint32_t m0 = (2 * q >= rel) ? w_lo : 0u, m1 = (2 * q + 1 >= rel) ? w_hi : 0u;
                    if constexpr (C == 1) {
                        a[0] = dp2a_s(xw[q], m0 | m1, a[0]);
                    } else {
                        a[(2 * q) % C] = dp2a_s(xw[q], m0, a[(2 * q) % C]);
                        a[(2 * q + 1) % C] = dp2a_s(xw[q], m1, a[(2 * q + 1) % C]);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) d[c] = (uint32_t)a[c];
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] = 0u;
            }

            uint32_t incl[C];
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] = d[c];
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl[c], dd);
                    if (lane >= dd) incl[c] += up;
                }
            }
            const uint32_t wbase = wraw + (it & 1u) * (32u * CJ * 4u);
            if (lane == 31) {
#pragma unroll
                for (int c = 0; c < C; ++c) sts32u(wbase + (uint32_t)((c / G) * 32 + (c % G) * NW + warp) * 4u, incl[c]);
            }
#pragma unroll
            for (int c = 0; c < C; ++c) incl[c] -= d[c];

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);
            if (tid == 0 && is_out && tile + PL < t1) {
                int s2 = lst + PL;
                if (s2 >= SL) s2 -= SL;
                issue_lag(tile + PL, sig, s2);       // that stage was read into registers before the barrier
            }

            uint32_t excl[CJ];
#pragma unroll
            for (int jj = 0; jj < CJ; ++jj) {
                const uint32_t v = lds32u(wbase + (uint32_t)(jj * 32 + lane) * 4u);
                uint32_t wi = v + (wq == 0 ? Wl[jj] : 0u);
#pragma unroll
                for (int dd = 1; dd < NW; dd <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, wi, dd);
                    if (wq >= dd) wi += up;
                }
                Wl[jj] = __shfl_sync(0xffffffffu, wi, wlast);
                excl[jj] = wi - v;
            }
            uint32_t start[C];
#pragma unroll
            for (int c = 0; c < C; ++c)
                start[c] = __shfl_sync(0xffffffffu, excl[c / G], (c % G) * NW + warp) + incl[c];

            if (is_out) {
                const uint32_t ob = tr.out_tile();   // = cur: the results replace the own tile
#pragma unroll
                for (int c = 0; c < CH_OWN; ++c) {
                    uint32_t wds[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t y[2];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int r = 8 * c + 2 * q + hh;
                            y[hh] = div_trunc_mulhi((int)(start[r % C] + s[r]), mul, sh);
                        }
                        wds[q] = __byte_perm(y[0], y[1], 0x5410);
                    }
                    sts128u(ob + (uint32_t)xo[c], wds[0], wds[1], wds[2], wds[3]);
                }
                tr.staged(tile, sig);
                ++lagit;
                lst = (lst + 1 == SL) ? 0 : lst + 1;
            }
            tr.advance();
        }
        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// Column kernel -- many-channel interleaved float32 ([frame][channel], C >= 32), e.g.
// BASELINE config 5's 256-channel batch in the reference layout.
// One CTA streams [FT frames x 32 channels] tiles of one 32-channel column block down the
// frames (2-D TMA boxes, no swizzle: lane = channel, so every shared-memory access is one
// conflict-free 128-byte row).  Warp w owns frames [w*RF, (w+1)*RF) of the tile, lane l one
// channel.  Window sum at the run start = direct additions of whole-group totals (groups ->
// tile totals, no subtraction anywhere) + head of the lag run; then w += x[f] - x[f-k].
// Outputs are stored straight from registers: each warp store is one full 128-byte row segment.
// ----------------------------------------------------------------------------------
struct ColsParams {
    float inv_k;
    uint32_t k;
    uint32_t n_full;       // whole RF-frame groups strictly between the lag group and the own group
    uint32_t m_part;       // leading frames of the lag run that complete the window at the run start
    uint32_t channels;
    uint64_t frames;       // frames of this shard
    int32_t col_blocks;    // ceil(channels / channels-per-tile)
    int32_t tiles_per_col;
    int32_t chunk_tiles;
    int32_t chunks_per_col;
    int32_t total_chunks;  // col_blocks * chunks_per_col
    int32_t hist_tiles;
    int32_t stages;
    int32_t prefetch;
    int32_t has_halo;
};

// NWT warps in total; tile bytes and summary sizes do not depend on how they are split into columns
__host__ __device__ inline uint32_t cols_smem_bytes(int NWT, int RF, int S, int H, uint32_t total_bytes = 4)
{
    return 1024u + (uint32_t)S * NWT * RF * 128u + (uint32_t)(H + 2) * NWT * 32 * total_bytes +
           (uint32_t)(H + 2) * 256 * total_bytes + (uint32_t)S * 8;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            uint64_t hint)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

// NWT warps per CTA arranged as CWW column-warps (32 channels each, side by side) x NW = NWT / CWW
// frame-warps (RF frames each).  Wider tiles (CWW > 1) read longer contiguous row pieces from HBM but
// hold fewer frames of history, so the host picks the widest shape whose window still fits.
template <int NWT, int RF, int CWW>
__global__ void __launch_bounds__(NWT * 32)
    stream_cols_f32_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap halo_map,
                           float* __restrict__ out, const ColsParams p)
{
    constexpr int NW = NWT / CWW;            // frame-warps
    constexpr int CW = 32 * CWW;             // channels per tile
    constexpr uint32_t ROWB = CW * 4u;       // bytes per tile row (one frame)
    constexpr int FT = NW * RF;              // frames per tile
    constexpr uint32_t TB = FT * ROWB;       // bytes per tile
    static_assert(NWT % CWW == 0 && FT <= 256 && CW <= 256, "a TMA box holds at most 256 x 256 elements");

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = (tid & 31) + 32 * ((tid >> 5) % CWW);   // channel inside the tile
    const int warp = (tid >> 5) / CWW;                       // frame-warp
    const int S = p.stages;
    const int H = p.hist_tiles;
    const int P = p.prefetch;
    const int GS = H + 2;

    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring_bytes = (uint32_t)S * TB;
    const uint32_t gsum = ring + ring_bytes;                    // float [GS][NW][CW]
    const uint32_t ttot = gsum + (uint32_t)GS * NW * CW * 4;    // float [GS][CW]
    const uint32_t bars = ttot + (uint32_t)GS * CW * 4;

    if (tid == 0) {
        prefetch_tmap(&in_map);
        if (p.has_halo) prefetch_tmap(&halo_map);
        for (int s = 0; s < S; ++s) mbar_init(bars + 8u * s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue_load = [&](int tile, int cb, int st) {
        const uint32_t bar = bars + 8u * st;
        mbar_arrive_expect_tx(bar, TB);
        if (tile < 0 && p.has_halo)
            tma_load_2d(ring + (uint32_t)st * TB, &halo_map, bar, cb * CW, (tile + H) * FT, kEvictFirst);
        else
            tma_load_2d(ring + (uint32_t)st * TB, &in_map, bar, cb * CW, tile * FT, kEvictFirst);
    };

    uint32_t it = 0;
    int st = 0, slot = 0;

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        // consecutive chunk ids sweep the column blocks of one frame range, so neighbouring CTAs
        // read neighbouring 128-byte segments of the same rows at about the same time
        const int rng = chunk / p.col_blocks;
        const int cb = chunk - rng * p.col_blocks;
        const int t0 = rng * p.chunk_tiles;
        int t1 = t0 + p.chunk_tiles;
        if (t1 > p.tiles_per_col) t1 = p.tiles_per_col;
        if (t0 >= t1) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        const uint32_t ch = (uint32_t)cb * CW + lane;
        const bool ch_ok = ch < p.channels;

        if (tid == 0) {
            int s2 = st;
            for (int j = 0; j < P && j < ntl; ++j) {
                issue_load(first + j, cb, s2);
                s2 = (s2 + 1 == S) ? 0 : s2 + 1;
            }
        }

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = ring + (uint32_t)st * TB;

            mbar_wait(bars + 8u * st, (it / (uint32_t)S) & 1u);

            float x[RF];
#pragma unroll
            for (int r = 0; r < RF; ++r) x[r] = lds32(cur + ((uint32_t)(warp * RF + r) * CW + lane) * 4u);
            float gtot;
            {
                float q[RF / 4];
#pragma unroll
                for (int c = 0; c < RF / 4; ++c) q[c] = (x[4 * c] + x[4 * c + 1]) + (x[4 * c + 2] + x[4 * c + 3]);
                gtot = (q[0] + q[1]) + (q[2] + q[3]);
                if constexpr (RF == 32) gtot += (q[RF / 4 - 4] + q[RF / 4 - 3]) + (q[RF / 4 - 2] + q[RF / 4 - 1]);
            }
            sts32(gsum + (((uint32_t)slot * NW + warp) * CW + lane) * 4u, gtot);

            __syncthreads();

            if (tid == 0 && j + P < ntl) {
                int s2 = st + P;
                if (s2 >= S) s2 -= S;
                issue_load(first + j + P, cb, s2);
            }

            // where the lag run starts: `lw` = group (warp slot), `h` = tiles back (both warp-uniform)
            int lw = warp - (int)(p.n_full + 1u);
            int h = 0;
            if (lw < 0) {
                h = (-lw + NW - 1) / NW;
                lw += h * NW;
            }
            // groups of this tile in front of the own group: needed when the window reaches into an earlier
            // tile, and by the last warp for the tile total (which only windows longer than a tile consume)
            float e_own = 0.f;
            if (h > 0 || (H > 1 && warp == NW - 1)) {
                const uint32_t g0 = gsum + ((uint32_t)slot * NW * CW + lane) * 4u;
#pragma unroll 4
                for (int w2 = 0; w2 < warp; ++w2) e_own += lds32(g0 + (uint32_t)w2 * ROWB);
            }
            if (H > 1 && warp == NW - 1) sts32(ttot + ((uint32_t)slot * CW + lane) * 4u, e_own + gtot);

            if (is_out && p.k <= 8u) {
                // Tiny windows: additions only, straight from the ring (a subtractive update could exceed 1e-5
                // relative error where a 3-term window sum is nearly zero).  Row r of the own run needs the k-1
                // rows above it; they sit in this tile or at the end of the previous one.
                const int row_base = (int)((uint32_t)st * (uint32_t)FT) + warp * RF;
                const int ring_rows = S * FT;
                const float inv = p.inv_k;
                const uint64_t f_base = (uint64_t)tile * FT + (uint64_t)warp * RF;
                float* dst = out + f_base * p.channels + ch;
                int nvalid = 0;
                if (ch_ok && f_base < p.frames) nvalid = (p.frames - f_base < (uint64_t)RF) ? (int)(p.frames - f_base) : RF;
                const uint32_t a0 = ring + (uint32_t)lane * 4u;
                float prev[7];   // rows row_base-7 .. row_base-1 (oldest first)
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    int rr = row_base - 7 + j;
                    if (rr < 0) rr += ring_rows;
                    prev[j] = lds32(a0 + (uint32_t)rr * ROWB);
                }
#pragma unroll
                for (int r = 0; r < RF; ++r) {
                    // window = x[r], x[r-1], ..., k terms, oldest terms added first in a fixed order
                    float acc = 0.f;
#pragma unroll
                    for (int j = 7; j >= 1; --j) {
                        if ((uint32_t)j < p.k) acc += (r - j >= 0) ? x[(r - j) < 0 ? 0 : (r - j)] : prev[7 + (r - j)];
                    }
                    acc += x[r];
                    if (r < nvalid) dst[(uint32_t)r * p.channels] = acc * inv;
                }
            } else if (is_out) {
                float xl[RF];
                {
                    // 16 consecutive ring rows starting k frames above the own run; the ring wraps at most
                    // once inside the run, and where it does is warp-uniform (lane offsets stay inside a row)
                    int row0 = (int)((uint32_t)st * (uint32_t)FT) + warp * RF - (int)p.k;   // ring row of the first lag frame
                    const int ring_rows = S * FT;
                    if (row0 + RF <= 0) row0 += ring_rows;
                    if (row0 >= 0) {
                        const uint32_t a0 = ring + (uint32_t)row0 * ROWB + (uint32_t)lane * 4u;
#pragma unroll
                        for (int r = 0; r < RF; ++r) xl[r] = lds32(a0 + ROWB * r);
                    } else {  // rows row0..-1 live at the end of the ring
                        const uint32_t a0 = ring + (uint32_t)lane * 4u;
#pragma unroll
                        for (int r = 0; r < RF; ++r) {
                            const int rr = row0 + r;
                            xl[r] = lds32(a0 + (uint32_t)(rr < 0 ? rr + ring_rows : rr) * ROWB);
                        }
                    }
                }
                float acc = 0.f;
                if (h == 0) {
                    const uint32_t g0 = gsum + ((uint32_t)slot * NW * CW + lane) * 4u;
#pragma unroll 4
                    for (int w2 = lw + 1; w2 < warp; ++w2) acc += lds32(g0 + (uint32_t)w2 * ROWB);
                } else {
                    int ls = slot - h;
                    if (ls < 0) ls += GS;
                    const uint32_t g0 = gsum + ((uint32_t)ls * NW * CW + lane) * 4u;
#pragma unroll 4
                    for (int w2 = lw + 1; w2 < NW; ++w2) acc += lds32(g0 + (uint32_t)w2 * ROWB);
                    int ms = ls;
                    for (int v = 1; v < h; ++v) {
                        ms = (ms + 1 == GS) ? 0 : ms + 1;
                        acc += lds32(ttot + ((uint32_t)ms * CW + lane) * 4u);
                    }
                    acc += e_own;
                }
#pragma unroll
                for (int r = 0; r < RF; ++r)
                    if ((uint32_t)r < p.m_part) acc += xl[r];

                const float inv = p.inv_k;
                const uint64_t f_base = (uint64_t)tile * FT + (uint64_t)warp * RF;
                float* dst = out + f_base * p.channels + ch;
                // frames of this run inside the signal (warp-uniform); channels past C store nothing
                int nvalid = 0;
                if (ch_ok && f_base < p.frames) nvalid = (p.frames - f_base < (uint64_t)RF) ? (int)(p.frames - f_base) : RF;
                const uint32_t cstride = p.channels;
                if (nvalid == RF) {
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        acc += x[r] - xl[r];
                        dst[(uint32_t)r * cstride] = acc * inv;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        acc += x[r] - xl[r];
                        if (r < nvalid) dst[(uint32_t)r * cstride] = acc * inv;
                    }
                }
            }

            ++it;
            st = (st + 1 == S) ? 0 : st + 1;
            slot = (slot + 1 == GS) ? 0 : slot + 1;
        }
        __syncthreads();   // ring stages may be refilled by the next chunk's prologue
    }
}

// ----------------------------------------------------------------------------------
// Column kernel for interleaved int16 with many channels (C >= 64, C % 8 == 0; sensor arrays, multichannel PCM):
// stream_cols_f32_kernel's layout over 32-bit WORDS -- a frame is C/2 words, each a pair of neighbouring channels,
// lane = word column, warp = RF-frame run -- with stream_i16_kernel's exact delta-scan arithmetic (round 2; the
// round-1 version summed run totals, group totals and the head of the lag run: 21.9 executed instructions per sample):
//   * before the tile's barrier a thread slides over its RF frames with two dp2a per sample on the packed words,
//     s[r] = sum_{i <= r} (x[i] - x[i - k]) per channel of the pair, and publishes d = s[RF - 1];
//   * after the barrier it adds up the d of the NW frame-warps of its column (all of them: the total advances the
//     carried window sum W; those in front of its own warp give its start), and
//     y[r] = trunc((W + start + s[r]) / k) by multiply-high, two results merged by one byte permute, one 32-bit store;
//   * a chunk builds its first W from the H = ceil(k / tile frames) tiles in front of it (frames at or behind
//     chunk start - k, the first tile masked).
// p.channels = C/2 (words per frame).  Bit-identical to profilable_cpu_computations.  4 B/sample.
// ----------------------------------------------------------------------------------
// TMAST: results leave through a double-buffered staging tile and one TMA store per tile (which also clips the rows
// and columns past the end of the signal) instead of one 32-bit global store per word: a store from registers cost
// 8 of 23 instructions per word (64-bit row addressing, bounds checks), and bursts of them slowed the kernel down
// further (measured: all stores of a run back to back, 0.125 -> 0.163 ms on 2^27 samples).  The staging tiles take two
// ring stages, so the longest windows of a channel count run with TMAST = false.
__host__ __device__ inline uint32_t cols_i16_smem_bytes(int NWT, int RF, int S, bool tmast = false)
{
    return 1024u + (uint32_t)(S + (tmast ? 2 : 0)) * NWT * RF * 128u + 2u * NWT * 32 * 8u + (uint32_t)S * 8;
}

template <int NWT, int RF, int CWW, bool TMAST = false>
__global__ void __launch_bounds__(NWT * 32)
    stream_cols_i16x2_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap halo_map,
                             const __grid_constant__ CUtensorMap out_map, uint32_t* __restrict__ out, const ColsParams p,
                             const uint32_t div_mul, const uint32_t div_shift, const uint32_t wscale)
{
    constexpr int NW = NWT / CWW;            // frame-warps
    constexpr int CW = 32 * CWW;             // words per tile row
    constexpr uint32_t ROWB = CW * 4u;       // bytes per tile row (one frame)
    constexpr uint32_t GROW = CW * 8u;       // bytes per row of the [NW][CW] int2 delta array
    constexpr int FT = NW * RF;              // frames per tile
    constexpr uint32_t TB = FT * ROWB;       // bytes per tile
    static_assert(NWT % CWW == 0 && FT <= 256 && CW <= 256, "a TMA box holds at most 256 x 256 elements");

    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int lane = (tid & 31) + 32 * ((tid >> 5) % CWW);   // word column inside the tile
    const int warp = (tid >> 5) / CWW;                       // frame-warp
    const int S = p.stages;
    const int H = p.hist_tiles;
    const int P = p.prefetch;

    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring_bytes = (uint32_t)S * TB;
    const uint32_t outb = ring + ring_bytes;                    // TMAST: two staging tiles
    const uint32_t dsum = outb + (TMAST ? 2u * TB : 0u);        // int2 [2][NW][CW]: run deltas (low / high channel), by tile parity
    const uint32_t bars = dsum + 2u * NW * GROW;
    const uint32_t w_lo = wscale, w_hi = wscale << 8;
    const uint32_t n_lo = (0u - wscale) & 0xffu, n_hi = n_lo << 8;
    const int mul = (int)div_mul;

    if (tid == 0) {
        prefetch_tmap(&in_map);
        if (p.has_halo) prefetch_tmap(&halo_map);
        if (TMAST) prefetch_tmap(&out_map);
        for (int s = 0; s < S; ++s) mbar_init(bars + 8u * s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue_load = [&](int tile, int cb, int st) {
        const uint32_t bar = bars + 8u * st;
        mbar_arrive_expect_tx(bar, TB);
        if (tile < 0 && p.has_halo)
            tma_load_2d(ring + (uint32_t)st * TB, &halo_map, bar, cb * CW, (tile + H) * FT, kEvictFirst);
        else
            tma_load_2d(ring + (uint32_t)st * TB, &in_map, bar, cb * CW, tile * FT, kEvictFirst);
    };

    uint32_t it = 0;
    int st = 0;
    // TMAST, thread 0: the store of a staged tile is issued one iteration late (behind the next barrier) and its
    // staging buffer is reused two tiles later, after cp.async.bulk.wait_group.read
    uint32_t otiles = 0;
    bool st_pending = false, st_inflight = false;
    int st_tile = 0, st_cb = 0;
    uint32_t st_buf = 0;
    auto flush_store = [&]() {
        if (st_pending) {
            tma_store_2d(&out_map, st_buf, st_cb * CW, st_tile * FT);
            tma_commit();
            st_pending = false;
            st_inflight = true;
        }
    };

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        const int rng = chunk / p.col_blocks;
        const int cb = chunk - rng * p.col_blocks;
        const int t0 = rng * p.chunk_tiles;
        int t1 = t0 + p.chunk_tiles;
        if (t1 > p.tiles_per_col) t1 = p.tiles_per_col;
        if (t0 >= t1) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        const uint32_t ch = (uint32_t)cb * CW + lane;
        const bool ch_ok = ch < p.channels;
        (void)ch_ok;

        if (tid == 0) {
            int s2 = st;
            for (int j = 0; j < P && j < ntl; ++j) {
                issue_load(first + j, cb, s2);
                s2 = (s2 + 1 == S) ? 0 : s2 + 1;
            }
        }
        uint32_t W0 = 0u, W1 = 0u;           // window sums of the column's two channels at the first frame of the tile

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = ring + (uint32_t)st * TB;

            mbar_wait(bars + 8u * st, (it / (uint32_t)S) & 1u);

            uint32_t x[RF];
#pragma unroll
            for (int r = 0; r < RF; ++r) x[r] = lds32u(cur + ((uint32_t)(warp * RF + r) * CW + lane) * 4u);
            uint32_t s0[RF], s1[RF];
            int a0 = 0, a1 = 0;
            if (is_out) {
                uint32_t xl[RF];
                int row0 = (int)((uint32_t)st * (uint32_t)FT) + warp * RF - (int)p.k;   // ring row of the first lag frame
                const int ring_rows = S * FT;
                if (row0 + RF <= 0) row0 += ring_rows;
                if (row0 >= 0) {
                    const uint32_t b0 = ring + (uint32_t)row0 * ROWB + (uint32_t)lane * 4u;
#pragma unroll
                    for (int r = 0; r < RF; ++r) xl[r] = lds32u(b0 + ROWB * r);
                } else {  // rows row0..-1 live at the end of the ring
                    const uint32_t b0 = ring + (uint32_t)lane * 4u;
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        const int rr = row0 + r;
                        xl[r] = lds32u(b0 + (uint32_t)(rr < 0 ? rr + ring_rows : rr) * ROWB);
                    }
                }
#pragma unroll
                for (int r = 0; r < RF; ++r) {
                    a0 = dp2a_s(xl[r], n_lo, dp2a_s(x[r], w_lo, a0));
                    a1 = dp2a_s(xl[r], n_hi, dp2a_s(x[r], w_hi, a1));
                    s0[r] = (uint32_t)a0;
                    s1[r] = (uint32_t)a1;
                }
            } else {
                // warm-up tile: frames at or behind (chunk start - k) belong to the first output's window
                const int rel = (t0 - tile) * FT - (int)p.k - warp * RF;
#pragma unroll
                for (int r = 0; r < RF; ++r) {
                    if (r >= rel) {
                        a0 = dp2a_s(x[r], w_lo, a0);
                        a1 = dp2a_s(x[r], w_hi, a1);
                    }
                    s0[r] = s1[r] = 0u;
                }
            }
            const uint32_t dbase = dsum + (it & 1u) * NW * GROW + (uint32_t)lane * 8u;
            sts64i(dbase + (uint32_t)warp * GROW, a0, a1);

            if (TMAST && tid == 0 && st_inflight) {      // the staging buffer about to be rewritten is free again
                tma_wait_read0();
                st_inflight = false;
            }
            __syncthreads();

            if (tid == 0) {
                if (j + P < ntl) {
                    int s2 = st + P;
                    if (s2 >= S) s2 -= S;
                    issue_load(first + j + P, cb, s2);
                }
                if (TMAST) flush_store();
            }

            uint32_t e0 = 0u, e1 = 0u, tt0 = 0u, tt1 = 0u;   // deltas of the frame-warps in front / of all of them
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) {
                const int2 t = lds64i(dbase + (uint32_t)w2 * GROW);
                tt0 += (uint32_t)t.x, tt1 += (uint32_t)t.y;
                if (w2 < warp) e0 += (uint32_t)t.x, e1 += (uint32_t)t.y;
            }

            if (is_out) {
                const uint32_t b0 = W0 + e0, b1 = W1 + e1;
                if constexpr (TMAST) {
                    const uint32_t ob = outb + (otiles & 1u) * TB + ((uint32_t)(warp * RF) * CW + lane) * 4u;
#pragma unroll
                    for (int r = 0; r < RF; ++r)
                        sts32u(ob + (uint32_t)r * ROWB, __byte_perm(div_trunc_mulhi((int)(b0 + s0[r]), mul, div_shift),
                                                                    div_trunc_mulhi((int)(b1 + s1[r]), mul, div_shift), 0x5410));
                    fence_proxy_async_smem();
                    if (tid == 0) {
                        st_pending = true;
                        st_tile = tile;
                        st_cb = cb;
                        st_buf = outb + (otiles & 1u) * TB;
                    }
                    ++otiles;
                } else {
                    const uint64_t f_base = (uint64_t)tile * FT + (uint64_t)warp * RF;
                    uint32_t* dst = out + f_base * p.channels + ch;
                    int nvalid = 0;
                    if (ch_ok && f_base < p.frames) nvalid = (p.frames - f_base < (uint64_t)RF) ? (int)(p.frames - f_base) : RF;
                    const uint32_t cstride = p.channels;
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        const uint32_t y = __byte_perm(div_trunc_mulhi((int)(b0 + s0[r]), mul, div_shift),
                                                       div_trunc_mulhi((int)(b1 + s1[r]), mul, div_shift), 0x5410);
                        if (r < nvalid) dst[(uint32_t)r * cstride] = y;
                    }
                }
            }
            W0 += tt0, W1 += tt1;

            ++it;
            st = (st + 1 == S) ? 0 : st + 1;
        }
        __syncthreads();   // ring stages may be refilled by the next chunk's prologue
        if (TMAST && tid == 0) flush_store();
    }
    if (TMAST && tid == 0) tma_wait_all0();
}

// ----------------------------------------------------------------------------------
// Few-channel kernel -- interleaved float32 with 3..31 channels (5.1 / 7.1 audio, ...).
// Flat TMA tiles of the interleaved stream exactly like the mono kernel (TileRing, TMA store), but the work is
// split the way the column kernel does it: thread = (run of RF frames, channel), neighbouring threads take
// neighbouring channels of the same run.  NR = floor(512 / C) runs per tile, rounded down so that NR * C is a
// multiple of 16 (tile = whole 1024-byte swizzle atoms), tile = NR * RF frames.  Tiles use SWIZZLE_128B: with a
// dense layout the runs of one warp start 16*C words apart and collide on one or two banks (5-way conflicts for
// C = 3 or 6); the XOR with the 128-byte row index spreads them.  Window start sum = direct additions of the <= 16 preceding
// run totals of the same channel (this tile or the previous one) + head of the lag run; k <= 8 takes the
// additions-only path.  8 B/sample.
// ----------------------------------------------------------------------------------
struct FewcParams {
    StreamParams sp;       // ring geometry, k, inv_k, n_full, m_part (in frames)
    uint32_t channels;
    uint32_t runs;         // NR
    uint32_t long_mode;    // whole runs of the window summed through per-tile prefixes instead of one by one
};

__host__ __device__ inline uint32_t fewc_smem_bytes(uint32_t tile_bytes, int S, int H, uint32_t active_threads,
                                                    uint32_t total_bytes = 4)
{
    return 1024u + (uint32_t)S * tile_bytes + 2u * tile_bytes + (uint32_t)(H + 2) * active_threads * total_bytes +
           (uint32_t)S * 8 + 64;
}

// ---- long windows in the few-channel kernels (more than 16 whole runs between the lag run and the own run) ----
// After the tile's barrier every warp turns the run totals of "its" channels (warp, warp + 16, ...) of the
// current slot into inclusive prefixes over the tile's runs, IN PLACE; a second barrier publishes them.  The sum
// of the n_full runs in front of the own run is then a difference of two prefixes of this tile, or prefix +
// whole-tile totals + (tile total - prefix) of the oldest tile.  Every difference stays inside one tile, so the
// float32 cancellation error is bounded by eps * (tile sum), as in the mono kernel's MODE 1; the tile grid is
// anchored at global frame 0, so shards and slices reproduce the unsharded sums bit for bit.
template <typename T>
__device__ __forceinline__ T lds_t(uint32_t addr);
template <>
__device__ __forceinline__ float lds_t<float>(uint32_t addr) { return lds32(addr); }
template <>
__device__ __forceinline__ int lds_t<int>(uint32_t addr) { return lds32i(addr); }
__device__ __forceinline__ void sts_t(uint32_t addr, float v) { sts32(addr, v); }
__device__ __forceinline__ void sts_t(uint32_t addr, int v) { sts32i(addr, v); }

// slot_base: totals of the current tile, entry (run * C + c) holds NV values of type T
template <typename T, int NV>
__device__ __forceinline__ void fewc_scan_slot(uint32_t slot_base, int NR, uint32_t C)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int q = (NR + 31) >> 5;                          // runs per lane, <= 8 (NR <= 256)
    for (uint32_t c = (uint32_t)warp; c < C; c += (uint32_t)nwarps) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            T loc[8];
            T run_sum = T(0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = lane * q + i;
                loc[i] = T(0);
                if (i < q && r < NR) {
                    run_sum += lds_t<T>(slot_base + (((uint32_t)r * C + c) * NV + v) * 4u);
                    loc[i] = run_sum;
                }
            }
            T incl = run_sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const T excl = incl - run_sum;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = lane * q + i;
                if (i < q && r < NR) sts_t(slot_base + (((uint32_t)r * C + c) * NV + v) * 4u, loc[i] + excl);
            }
        }
    }
}

// sum of the n_full runs in front of run `run` of channel c (prefixes as left by fewc_scan_slot)
template <typename T, int NV>
__device__ __forceinline__ void fewc_window_long(T (&acc)[NV], uint32_t gsum, int slot, int GS, int NR, uint32_t C,
                                                 uint32_t c, int run, int n_full)
{
    const uint32_t slot_bytes = (uint32_t)NR * C * NV * 4u;
    auto at = [&](int s, int r, int v) -> T {
        return lds_t<T>(gsum + (uint32_t)s * slot_bytes + (((uint32_t)r * C + c) * NV + v) * 4u);
    };
    const int q0 = run - n_full;                           // first run of the sum, in this tile's numbering
    if (q0 >= 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            T t = at(slot, run - 1, v);
            if (q0 > 0) t -= at(slot, q0 - 1, v);
            acc[v] = t;
        }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = run > 0 ? at(slot, run - 1, v) : T(0);
        int m = -q0;                                       // runs still to take from older tiles
        int s2 = slot;
        while (m > NR) {
            s2 = (s2 == 0) ? GS - 1 : s2 - 1;
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] += at(s2, NR - 1, v);
            m -= NR;
        }
        s2 = (s2 == 0) ? GS - 1 : s2 - 1;
#pragma unroll
        for (int v = 0; v < NV; ++v) {                     // the last m runs of the oldest tile, 1 <= m <= NR
            T t = at(s2, NR - 1, v);
            if (m < NR) t -= at(s2, NR - 1 - m, v);
            acc[v] += t;
        }
    }
}

template <int RF>
__global__ void __launch_bounds__(512)
    stream_fewc_f32_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                           const __grid_constant__ CUtensorMap halo_map, const FewcParams fp)
{
    const StreamParams& p = fp.sp;
    const uint32_t C = fp.channels;
    const int NR = (int)fp.runs;
    const uint32_t NRC = (uint32_t)NR * C;                 // active threads
    const uint32_t tile_bytes = NRC * RF * 4u;
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const bool active = (uint32_t)tid < NRC;
    const int run = active ? (int)((uint32_t)tid / C) : 0;
    const uint32_t c = active ? (uint32_t)tid - (uint32_t)run * C : 0u;

    TileRing<0, 0> tr;
    const uint32_t gsum = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &halo_map, tile_bytes, (int)(tile_bytes / 128u));
    const int H = p.hist_tiles, GS = p.hist_tiles + 2;      // float [GS][NRC] run totals
    tr.init_barriers((gsum + (uint32_t)GS * NRC * 4u + 7u) & ~7u);
    const uint32_t row_stride = C * 4u;                     // bytes between consecutive frames of one channel
    const bool long_mode = fp.long_mode != 0;
    const uint32_t own = ((uint32_t)run * RF * C + c) * 4u;  // byte offset of the run's first sample in the tile
    int xo[RF], xg[RF];                                     // pre-swizzled offsets: own run, lag run
#pragma unroll
    for (int r = 0; r < RF; ++r) {
        xo[r] = pre_swz((int)(own + (uint32_t)r * row_stride));
        xg[r] = pre_swz((int)own + (r - (int)p.k) * (int)row_stride);
    }

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = tr.wait_tile();
            const int slot = tr.slot;

            float x[RF];
#pragma unroll
            for (int r = 0; r < RF; ++r) x[r] = active ? lds32(cur + (uint32_t)xo[r]) : 0.f;
            float gtot;
            {
                float q[RF / 4];
#pragma unroll
                for (int i = 0; i < RF / 4; ++i) q[i] = (x[4 * i] + x[4 * i + 1]) + (x[4 * i + 2] + x[4 * i + 3]);
                gtot = (q[0] + q[1]) + (q[2] + q[3]);
            }
            if (active) sts32(gsum + ((uint32_t)slot * NRC + tid) * 4u, gtot);

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);
            if (long_mode) {
                fewc_scan_slot<float, 1>(gsum + (uint32_t)slot * NRC * 4u, NR, C);
                __syncthreads();
            }

            if (is_out) {
                const float inv = p.inv_k;
                const uint32_t ob = tr.out_tile();
                if (active && p.k <= 8u) {
                    // additions only: the k-1 frames above the run come from this tile or the end of the previous one
                    float prev[7];
#pragma unroll
                    for (int i = 0; i < 7; ++i) prev[i] = lds32(swz(tr.rel((int)own - (7 - i) * (int)row_stride)));
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        float acc = 0.f;
#pragma unroll
                        for (int i = 7; i >= 1; --i) {
                            if ((uint32_t)i < p.k) acc += (r - i >= 0) ? x[(r - i) < 0 ? 0 : (r - i)] : prev[7 + (r - i)];
                        }
                        acc += x[r];
                        sts32(ob + (uint32_t)xo[r], acc * inv);
                    }
                } else if (active) {
                    float xl[RF];
                    const uint32_t b0 = cur, b1 = cur + tr.ring_bytes;
                    const int neg_st = (int)tr.ring - (int)cur;
#pragma unroll
                    for (int r = 0; r < RF; ++r) xl[r] = lds32(ring_addr(xg[r], b0, b1, neg_st));
                    // run totals of the same channel between the lag run's group and the own run: D - 1 of them,
                    // nearest first; they sit in this tile and, for the first runs, at the end of the previous one
                    const int D = (int)p.n_full + 1;
                    float acc = 0.f;
                    if (long_mode) {
                        float a1[1];
                        fewc_window_long<float, 1>(a1, gsum, slot, GS, NR, C, c, run, D - 1);
                        acc = a1[0];
                    } else {
                        const uint32_t g_cur = gsum + ((uint32_t)slot * NRC + c) * 4u;
                        const int in_cur = (run < D - 1) ? run : D - 1;       // how many of them are in this tile
                        for (int w2 = 1; w2 <= in_cur; ++w2) acc += lds32(g_cur + (uint32_t)(run - w2) * row_stride);
                        if (in_cur < D - 1) {
                            const int ps = (slot == 0) ? GS - 1 : slot - 1;
                            const uint32_t g_prev = gsum + ((uint32_t)ps * NRC + c) * 4u;
                            for (int w2 = 1; w2 <= D - 1 - in_cur; ++w2) acc += lds32(g_prev + (uint32_t)(NR - w2) * row_stride);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RF; ++r)
                        if ((uint32_t)r < p.m_part) acc += xl[r];
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        acc += x[r] - xl[r];
                        sts32(ob + (uint32_t)xo[r], acc * inv);
                    }
                }
                tr.staged(tile, sig);
            }
            tr.advance();
        }
        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// Few-channel int16 twin: 3..31 interleaved int16 channels (multichannel PCM WAV), k >= 2.  Same layout
// and work split as stream_fewc_f32_kernel with 32-frame runs; int32-exact window sums and the multiply-high
// truncating division, so results stay bit-identical to profilable_cpu_computations.  4 B/sample.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ int lds16s(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, int v)
{
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((short)v) : "memory");
}
// C integer division (truncation toward zero) for |w| < 2^31: sign-mask abs, multiply-high, sign restore
__device__ __forceinline__ int div_trunc_i32(int w, uint32_t mul, uint32_t sh)
{
    const int sg = w >> 31;
    const uint32_t a = (uint32_t)((w ^ sg) - sg);
    const int q = (int)(__umulhi(a, mul) >> sh);
    return (q ^ sg) - sg;
}

template <int RF>
__global__ void __launch_bounds__(512)
    stream_fewc_i16_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                           const __grid_constant__ CUtensorMap halo_map, const FewcParams fp)
{
    const StreamParams& p = fp.sp;
    const uint32_t C = fp.channels;
    const int NR = (int)fp.runs;
    const uint32_t NRC = (uint32_t)NR * C;
    const uint32_t tile_bytes = NRC * RF * 2u;
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const bool active = (uint32_t)tid < NRC;
    const int run = active ? (int)((uint32_t)tid / C) : 0;
    const uint32_t c = active ? (uint32_t)tid - (uint32_t)run * C : 0u;

    TileRing<0, 0> tr;
    const uint32_t gsum = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &halo_map, tile_bytes, (int)(tile_bytes / 128u));
    const int H = p.hist_tiles, GS = p.hist_tiles + 2;      // int [GS][NRC] run totals
    tr.init_barriers((gsum + (uint32_t)GS * NRC * 4u + 7u) & ~7u);
    const uint32_t row_stride = C * 2u;                     // bytes between consecutive frames of one channel
    const bool long_mode = fp.long_mode != 0;
    const uint32_t g_stride = C * 4u;                       // bytes between consecutive runs' totals of one channel
    const uint32_t own = ((uint32_t)run * RF * C + c) * 2u;
    int xo[RF], xg[RF];                                     // pre-swizzled offsets: own run, lag run
#pragma unroll
    for (int r = 0; r < RF; ++r) {
        xo[r] = pre_swz((int)(own + (uint32_t)r * row_stride));
        xg[r] = pre_swz((int)own + (r - (int)p.k) * (int)row_stride);
    }

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = tr.wait_tile();
            const int slot = tr.slot;

            int x[RF];
            int gtot = 0;
#pragma unroll
            for (int r = 0; r < RF; ++r) {
                x[r] = active ? lds16s(cur + (uint32_t)xo[r]) : 0;
                gtot += x[r];
            }
            if (active) sts32i(gsum + ((uint32_t)slot * NRC + tid) * 4u, gtot);

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);
            if (long_mode) {
                fewc_scan_slot<int, 1>(gsum + (uint32_t)slot * NRC * 4u, NR, C);
                __syncthreads();
            }

            if (is_out) {
                if (active) {
                    const uint32_t ob = tr.out_tile();
                    const uint32_t b0 = cur, b1 = cur + tr.ring_bytes;
                    const int neg_st = (int)tr.ring - (int)cur;
                    int xl[RF];
#pragma unroll
                    for (int r = 0; r < RF; ++r) xl[r] = lds16s(ring_addr(xg[r], b0, b1, neg_st));
                    const int D = (int)p.n_full + 1;
                    int acc = 0;
                    if (long_mode) {
                        int a1[1];
                        fewc_window_long<int, 1>(a1, gsum, slot, GS, NR, C, c, run, D - 1);
                        acc = a1[0];
                    } else {
                        const uint32_t g_cur = gsum + ((uint32_t)slot * NRC + c) * 4u;
                        const int in_cur = (run < D - 1) ? run : D - 1;
                        for (int w2 = 1; w2 <= in_cur; ++w2) acc += lds32i(g_cur + (uint32_t)(run - w2) * g_stride);
                        if (in_cur < D - 1) {
                            const int ps = (slot == 0) ? GS - 1 : slot - 1;
                            const uint32_t g_prev = gsum + ((uint32_t)ps * NRC + c) * 4u;
                            for (int w2 = 1; w2 <= D - 1 - in_cur; ++w2) acc += lds32i(g_prev + (uint32_t)(NR - w2) * g_stride);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RF; ++r)
                        if ((uint32_t)r < p.m_part) acc += xl[r];
                    const uint32_t mul = p.div_mul, sh = p.div_shift;
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        acc += x[r] - xl[r];
                        sts16(ob + (uint32_t)xo[r], div_trunc_i32(acc, mul, sh));
                    }
                }
                tr.staged(tile, sig);
            }
            tr.advance();
        }
        tr.epilogue();
    }
    tr.finish();
}

// ----------------------------------------------------------------------------------
// Even channel counts (4, 6 = 5.1, 8 = 7.1, ... 30) of interleaved int16: the frame is C/2 32-bit words, each a
// pair of neighbouring channels, so the kernel is stream_fewc_f32_kernel's layout over WORDS (thread = 16-frame
// run of one channel pair, 32-bit shared-memory accesses instead of three 2-byte ones per sample) with
// stream_i16_kernel's arithmetic: dp2a on the packed words for run totals, head and slide, signed multiply-high
// division, one byte permute per output word.  fp.channels = C/2 here.
// ----------------------------------------------------------------------------------
template <int RF>
__global__ void __launch_bounds__(512)
    stream_fewc_i16x2_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                             const __grid_constant__ CUtensorMap halo_map, const FewcParams fp)
{
    const StreamParams& p = fp.sp;
    const uint32_t C = fp.channels;                        // channel PAIRS
    const int NR = (int)fp.runs;
    const uint32_t NRC = (uint32_t)NR * C;
    const uint32_t tile_bytes = NRC * RF * 4u;
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const bool active = (uint32_t)tid < NRC;
    const int run = active ? (int)((uint32_t)tid / C) : 0;
    const uint32_t c = active ? (uint32_t)tid - (uint32_t)run * C : 0u;

    TileRing<0, 0> tr;
    const uint32_t gsum = tr.setup(smem_u32(smem_raw), p, &in_map, &out_map, &halo_map, tile_bytes, (int)(tile_bytes / 128u));
    const int H = p.hist_tiles, GS = p.hist_tiles + 2;      // int [GS][NRC][2] run totals (low / high channel)
    tr.init_barriers((gsum + (uint32_t)GS * NRC * 8u + 7u) & ~7u);
    const uint32_t row_stride = C * 4u;                     // bytes between consecutive frames of one pair
    const bool long_mode = fp.long_mode != 0;
    const uint32_t g_stride = C * 8u;                       // bytes between consecutive runs' totals of one pair
    const uint32_t w_lo = p.wscale, w_hi = p.wscale << 8;
    const uint32_t n_lo = (0u - p.wscale) & 0xffu, n_hi = n_lo << 8;
    const uint32_t own = ((uint32_t)run * RF * C + c) * 4u;
    int xo[RF], xg[RF];                                     // pre-swizzled offsets: own run, lag run
#pragma unroll
    for (int r = 0; r < RF; ++r) {
        xo[r] = pre_swz((int)(own + (uint32_t)r * row_stride));
        xg[r] = pre_swz((int)own + (r - (int)p.k) * (int)row_stride);
    }

    for (int chunk = blockIdx.x; chunk < p.total_chunks; chunk += gridDim.x) {
        int sig, t0, t1;
        if (!chunk_range(p, chunk, sig, t0, t1)) continue;
        const int first = t0 - H;
        const int ntl = t1 - first;
        tr.prologue(first, ntl, sig);

        for (int j = 0; j < ntl; ++j) {
            const int tile = first + j;
            const bool is_out = (j >= H);
            const uint32_t cur = tr.wait_tile();
            const int slot = tr.slot;

            uint32_t x[RF];
            int g0 = 0, g1 = 0;
#pragma unroll
            for (int r = 0; r < RF; ++r) {
                x[r] = active ? lds32u(cur + (uint32_t)xo[r]) : 0u;
                g0 = dp2a_s(x[r], w_lo, g0);
                g1 = dp2a_s(x[r], w_hi, g1);
            }
            if (active) sts64i(gsum + ((uint32_t)slot * NRC + tid) * 8u, g0, g1);

            tr.before_sync();
            __syncthreads();
            tr.after_sync(j, ntl, first, sig);
            if (long_mode) {
                fewc_scan_slot<int, 2>(gsum + (uint32_t)slot * NRC * 8u, NR, C);
                __syncthreads();
            }

            if (is_out) {
                if (active) {
                    const uint32_t ob = tr.out_tile();
                    const uint32_t b0 = cur, b1 = cur + tr.ring_bytes;
                    const int neg_st = (int)tr.ring - (int)cur;
                    uint32_t xl[RF];
#pragma unroll
                    for (int r = 0; r < RF; ++r) xl[r] = lds32u(ring_addr(xg[r], b0, b1, neg_st));
                    const int D = (int)p.n_full + 1;
                    int a0 = 0, a1 = 0;
                    if (long_mode) {
                        int a2[2];
                        fewc_window_long<int, 2>(a2, gsum, slot, GS, NR, C, c, run, D - 1);
                        a0 = a2[0], a1 = a2[1];
                    } else {
                        const uint32_t g_cur = gsum + ((uint32_t)slot * NRC + c) * 8u;
                        const int in_cur = (run < D - 1) ? run : D - 1;
                        for (int w2 = 1; w2 <= in_cur; ++w2) {
                            const int2 t = lds64i(g_cur + (uint32_t)(run - w2) * g_stride);
                            a0 += t.x, a1 += t.y;
                        }
                        if (in_cur < D - 1) {
                            const int ps = (slot == 0) ? GS - 1 : slot - 1;
                            const uint32_t g_prev = gsum + ((uint32_t)ps * NRC + c) * 8u;
                            for (int w2 = 1; w2 <= D - 1 - in_cur; ++w2) {
                                const int2 t = lds64i(g_prev + (uint32_t)(NR - w2) * g_stride);
                                a0 += t.x, a1 += t.y;
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RF; ++r)
                        if ((uint32_t)r < p.m_part) {
                            a0 = dp2a_s(xl[r], w_lo, a0);
                            a1 = dp2a_s(xl[r], w_hi, a1);
                        }
                    const int mul = (int)p.div_mul;
                    const uint32_t sh = p.div_shift;
#pragma unroll
                    for (int r = 0; r < RF; ++r) {
                        a0 = dp2a_s(xl[r], n_lo, dp2a_s(x[r], w_lo, a0));
                        a1 = dp2a_s(xl[r], n_hi, dp2a_s(x[r], w_hi, a1));
                        sts32u(ob + (uint32_t)xo[r],
                               __byte_perm(div_trunc_mulhi(a0, mul, sh), div_trunc_mulhi(a1, mul, sh), 0x5410));
                    }
                }
                tr.staged(tile, sig);
            }
            tr.advance();
        }
        tr.epilogue();
    }
    tr.finish();
}

}  // namespace mavg
