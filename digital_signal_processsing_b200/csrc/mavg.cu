// mavg.cu -- host side of libmavg: plans, geometry, tensor maps, launches, C ABI.
//
// Boundary being replaced (see include/mavg.h): the XxxGpuLoad functions and the
// DspWorkspace of the reference (basics/*.cu, gpu_utils.h:67-160) and the CUDA-event
// phase timing of GpuTimer (benchmark.h:72-96).
#include "../../include/mavg.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "mavg_kernels.cuh"
#include "mavg_scan.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int status, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return status;
}

#define MAVG_CUDA(call)                                                                                \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? MAVG_ERR_NO_DEVICE \
                                                                                       : MAVG_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

#define MAVG_TRY(call)            \
    do {                          \
        int s_ = (call);          \
        if (s_ != MAVG_OK) return s_; \
    } while (0)

// cuTensorMapEncodeTiled is fetched through the runtime so that libmavg.so does not link
// libcuda.so (it must load on a box without a driver for the symbol-export tests).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn lookup_encoder()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(fn);
}

int get_encoder(EncodeTiledFn* out)
{
    static const EncodeTiledFn cached = lookup_encoder();   // initialised once, thread-safe (plans may live on different threads)
    if (!cached) return fail(MAVG_ERR_DRIVER, "cuTensorMapEncodeTiled unavailable in this driver");
    *out = cached;
    return MAVG_OK;
}

// One signal batch as a [signals][rows][128 bytes] tensor (32 floats or 64 int16 per row), boxes of
// [1][tile_rows][row], 128B swizzle.
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t signals, uint64_t signal_stride_bytes,
             uint32_t tile_rows, uint32_t elem_bytes, int swizzle = 128)
{
    EncodeTiledFn enc = nullptr;
    MAVG_TRY(get_encoder(&enc));
    const uint32_t row_elems = 128 / elem_bytes;
    cuuint64_t dims[3] = {row_elems, rows, signals};
    cuuint64_t strides[2] = {128, signals > 1 ? signal_stride_bytes : rows * 128};
    cuuint32_t box[3] = {row_elems, tile_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3,
                     const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                     : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                     : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                     : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(MAVG_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d) rows=%llu signals=%llu stride=%llu", (int)r,
                    (unsigned long long)rows, (unsigned long long)signals, (unsigned long long)signal_stride_bytes);
    return MAVG_OK;
}

size_t elem_size(uint32_t dtype) { return dtype == MAVG_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------
// Geometry of the streaming kernel for one (k, shard) pair
// ---------------------------------------------------------------------------------
struct StreamGeom {
    bool ok = false;
    int NT = 0, R = 0, MIS = 0, mode = 0;
    int H = 0, P = 0, S = 0;
    int ctas_per_sm = 0;
    uint32_t smem = 0;
    uint32_t n_full = 0, m_part = 0, lag_chunks = 0;
    uint32_t elem = 4;        // bytes per sample
    uint32_t C = 1;           // channels interleaved inside one kernel signal (int16 stereo: 2)
    uint32_t div_mul = 0, div_shift = 0, wscale = 1;
    bool long_mode = false;     // few-channel kernels: prefix mode
    int lag_stages = 2;         // far-lag kernel: stages of the lag ring
    bool pair = false;          // few-channel int16 with an even channel count: the kernel works on channel-pair words
    int cww = 1;              // column kernel: 32-channel column-warps side by side in one tile
    int runs = 0;             // few-channel kernel: 16-frame runs per tile
    bool tmast = false;       // int16 column kernel: results leave through staging tiles and TMA stores
};

constexpr uint32_t kMaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100

// int16: 512 threads x 32 samples (64 bytes) per tile of 16384 samples; flat lag distance k * C.
// Constants of the signed multiply-high division (proof: plan_stream_i16), 2 <= k <= 32768
void i16_mulhi_consts(uint32_t k, uint32_t* mul, uint32_t* shift, uint32_t* wscale)
{
    const uint32_t kd = k == 2 ? 4u : k;               // divisor seen by the kernel (weights doubled for k == 2)
    *wscale = k == 2 ? 2u : 1u;
    uint32_t lg = 0;
    while ((1u << lg) < kd) ++lg;                      // ceil(log2 kd) >= 2
    *mul = (uint32_t)((1ull << (30 + lg)) / kd + 1);   // < 2^31
    *shift = lg - 2;
}

StreamGeom plan_stream_i16_shape(uint32_t k, uint32_t C, const mavg_tuning& tu);

// 3+ channels: CTAs of 128 threads, two per SM (their latency phases overlap: 2^27 samples, 6 channels 0.108 -> 0.103 ms,
// 8 channels 0.095 -> 0.093 ms), while the window leaves them two tiles of prefetch; one larger CTA per SM beyond
StreamGeom plan_stream_i16(uint32_t k, uint32_t C, const mavg_tuning& tu)
{
    if (C > 2 && tu.threads == 0 && tu.ctas_per_sm == 0 && tu.prefetch == 0) {
        mavg_tuning t2 = tu;
        t2.threads = 128;
        const StreamGeom g = plan_stream_i16_shape(k, C, t2);
        if (g.ok && g.ctas_per_sm == 2 && g.P == 2) return g;
    }
    return plan_stream_i16_shape(k, C, tu);
}

StreamGeom plan_stream_i16_shape(uint32_t k, uint32_t C, const mavg_tuning& tu)
{
    StreamGeom g;
    g.NT = tu.threads == 256 ? 256 : 512;
    g.R = 32;    // 16-sample runs with two CTAs per SM were measured 10-20 % slower (per-run overheads dominate)
    // 3 / 4 / 6 / 8 interleaved channels (multichannel PCM) run on the same flat-stream kernel: a run holds whole
    // frames (24 samples for 3 and 6 channels), so the channel of run element r is r % C at compile time.  The scan of
    // C channel deltas per run is what these shapes pay for, so by default they run with half the threads and runs
    // twice as long (one CTA of 256 threads per SM; 2^27 samples, k = 64: 8 channels 0.108 -> 0.096 ms, 4 channels
    // 0.095 -> 0.091 ms); tuning.threads = 512 selects the 512-thread shape
    if (C == 4 || C == 8) {
        g.NT = tu.threads == 512 ? 512 : tu.threads == 128 ? 128 : 256;
        g.R = g.NT == 512 ? 32 : 64;
    } else if (C == 3 || C == 6) {                     // runs of an odd number of 16-byte chunks: dense tiles, no swizzle
        g.NT = tu.threads == 512 ? 512 : tu.threads == 128 ? 128 : 224;
        g.R = g.NT == 512 ? 24 : 72;
    } else if (C == 12 || C == 9) {
        g.NT = tu.threads == 128 ? 128 : 224;
        g.R = 72;
    } else if (C == 10) {
        g.NT = tu.threads == 128 ? 192 : 384;
        g.R = 40;
    } else if (C == 16) {
        g.NT = tu.threads == 128 ? 128 : 256;
        g.R = 64;
    } else if (C == 5) {                               // runs of lcm(C, 8) samples: 5 / 7 chunks, dense tiles
        g.NT = tu.threads == 128 ? 192 : 384;          // "128": the two-CTAs-per-SM shape (192 x 40 = 15 KB tiles)
        g.R = 40;
    } else if (C == 7) {
        g.NT = tu.threads == 128 ? 128 : 256;          // 128 x 56 = 14 KB tiles, two CTAs per SM
        g.R = 56;
    }
    g.elem = 2;
    g.C = C;
    // Window sums |w| <= 32768 k fit int32 for k <= 32768, and C's truncating w / k is computed exactly as
    //   t = mulhi_s32(w, M);  y = (t >> s) + (t >>> 31)      with L = ceil(log2 k), M = floor(2^(30+L) / k) + 1, s = L - 2:
    // e = M k - 2^(30+L) is in (0, k], so for 0 <= w: floor(w M / 2^(30+L)) = floor(w / k) because w e < 2^(30+L)
    // (2^15 k * k <= 2^(15+2L) <= 2^(30+L)); for w < 0 the product lies strictly below w / k by less than 1 / k,
    // so its floor is trunc(w / k) - 1 and the sign bit adds the 1 back.  M < 2^31 needs k >= 3; k == 2 runs with
    // every dp2a weight doubled (sums 2 w) and the constants of k = 4.  k == 1 (identity) and longer windows are
    // left to the generic kernel.
    if (k < 2 || k > 32768u || !(C <= 10 || C == 12 || C == 16)) return g;
    const uint64_t L = (uint64_t)k * C;
    const uint32_t R = (uint32_t)g.R;
    const uint32_t s = (uint32_t)((R - L % R) % R);
    g.m_part = R - s;
    g.n_full = (uint32_t)((L + s) / R - 1);
    g.lag_chunks = (uint32_t)((L + 7) / 8);
    g.MIS = (int)(8ull * g.lag_chunks - L);
    g.mode = 6;                                        // one arithmetic for every window: exclusive scan of run deltas
    const uint64_t T = (uint64_t)g.NT * R;
    g.H = (int)(((uint64_t)(g.n_full + 1) * R + T - 1) / T);
    g.ctas_per_sm = tu.ctas_per_sm ? (int)tu.ctas_per_sm : (g.NT * g.R * 2 <= 18432 ? 2 : 1);
    if (g.NT * g.R * 2 > 18432) g.ctas_per_sm = 1;
    g.P = tu.prefetch ? (int)tu.prefetch : 2;
    for (;;) {
        g.S = g.H + 1 + g.P;
        g.smem = mavg::stream_i16_smem_bytes(g.NT, g.R, g.S, g.H, (int)C);
        const uint32_t per_sm = 233472;
        if (g.smem <= kMaxSmem && (uint64_t)(g.smem + 1024) * g.ctas_per_sm <= per_sm) break;
        if (g.P > 1) { --g.P; continue; }
        if (g.ctas_per_sm > 1) { --g.ctas_per_sm; g.P = tu.prefetch ? (int)tu.prefetch : 2; continue; }
        return g;
    }
    i16_mulhi_consts(k, &g.div_mul, &g.div_shift, &g.wscale);
    g.ok = true;
    return g;
}

// float32: C = channels interleaved in the flat stream (1 for mono / planar, 2 for stereo).
StreamGeom plan_stream(uint32_t k, const mavg_tuning& tu, uint32_t C = 1)
{
    StreamGeom g;
    g.NT = tu.threads ? (int)tu.threads : 512;
    g.R = tu.run ? (int)tu.run : 16;
    g.C = C;
    if (!(g.NT == 256 || g.NT == 512) || !(g.R == 16 || g.R == 32)) return g;
    if (g.NT * g.R > 8192) return g;  // a TMA box holds at most 256 rows of 32 floats
    if (C != 1 && !(C == 2 && g.NT == 512 && g.R == 16)) return g;  // stereo is built for the default shape only
    const uint32_t direct_max = tu.direct_max_k ? tu.direct_max_k : 256u;
    const uint64_t L = (uint64_t)k * C;                              // lag distance in flat samples
    if (L > 0xfffffff0ull) return g;
    g.mode = (k <= 8) ? 2 : (L <= direct_max) ? 0 : 1;
    const uint32_t R = (uint32_t)g.R;
    const uint32_t s = (uint32_t)((R - L % R) % R);
    g.m_part = R - s;
    g.n_full = (uint32_t)((L + s) / R - 1);
    g.lag_chunks = (uint32_t)((L + 3) / 4);
    g.MIS = (g.mode == 2) ? 0 : (int)(4ull * g.lag_chunks - L);
    const uint64_t T = (uint64_t)g.NT * R;
    const uint64_t back = (g.mode == 2) ? 4 * ((7 * C + 3) / 4) : (uint64_t)(g.n_full + 1) * R;  // left context a tile can touch
    g.H = (int)((back + T - 1) / T);
    g.ctas_per_sm = tu.ctas_per_sm ? (int)tu.ctas_per_sm : 2;
    g.P = tu.prefetch ? (int)tu.prefetch : 2;
    // shrink the prefetch depth, then the residency, until the ring fits
    for (;;) {
        g.S = g.H + 1 + g.P;
        g.smem = mavg::stream_smem_bytes(g.NT, g.R, g.S, g.H, (int)C);
        const uint32_t per_sm = 233472;  // 228 KB per SM, 1 KB reserved per resident CTA
        if (g.smem <= kMaxSmem && (uint64_t)(g.smem + 1024) * g.ctas_per_sm <= per_sm) break;
        if (g.P > 1) { --g.P; continue; }
        if (g.ctas_per_sm > 1) { --g.ctas_per_sm; g.P = tu.prefetch ? (int)tu.prefetch : 2; continue; }
        return g;  // window too long for the shared-memory history
    }
    g.ok = true;
    return g;
}

// many-channel interleaved float32: 16 warps x 16 frames per tile of [256 frames][32 channels]
constexpr int kColsNW = 16, kColsRF = 16;
// float32: C >= 32, C % 4 == 0.  int16 (i16 = true): the kernel sees C/2 word columns (channel pairs), so
// C >= 64, C % 8 == 0, 2 <= k <= 32768 (exact int32 sums and division).
StreamGeom plan_cols(uint32_t k, uint32_t C, const mavg_tuning& tu, bool i16 = false)
{
    StreamGeom g;
    g.NT = kColsNW * 32;
    g.R = kColsRF;
    g.C = C;
    g.mode = 3;
    if (i16) {
        // int16: 8 warps x 32 frames by default (same tiles; half the threads with runs twice as long pay the per-run
        // work -- barrier, delta exchange, addressing -- half as often); tuning.threads = 512: 16 warps x 16 frames
        if (tu.threads != 512) { g.NT = 8 * 32; g.R = 32; }
        if (C % 8 != 0 || k < 2 || k > 32768u) return g;
        C /= 2;
        g.C = C;
        g.pair = true;
        g.elem = 4;
        i16_mulhi_consts(k, &g.div_mul, &g.div_shift, &g.wscale);
    }
    if (C < 32 || C % 4 != 0) return g;
    const uint32_t R = (uint32_t)g.R;
    const uint32_t s = (R - k % R) % R;
    g.m_part = R - s;
    g.n_full = (k + s) / R - 1;
    g.ctas_per_sm = 1;
    // widest tile (most contiguous bytes per row) whose history still leaves two tiles of prefetch; int16: with the
    // staging tiles of the TMA store while they fit (tuning.direct_max_k = 1 forces the stores from registers)
    for (int tmast = (i16 && tu.direct_max_k != 1) ? 1 : 0; tmast >= 0; --tmast) {
        g.tmast = tmast != 0;
        for (int cww = 8; cww >= 1; cww >>= 1) {
            if (32u * cww > C && cww > 1) continue;
            const uint32_t FT = (kColsNW / cww) * kColsRF;
            g.cww = cww;
            g.H = (int)(((uint64_t)(g.n_full + 1) * R + FT - 1) / FT);
            g.P = tu.prefetch ? (int)tu.prefetch : 2;
            bool fits = false;
            for (;;) {
                g.S = g.H + 1 + g.P;
                g.smem = i16 ? mavg::cols_i16_smem_bytes(g.NT / 32, g.R, g.S, g.tmast)
                             : mavg::cols_smem_bytes(kColsNW, kColsRF, g.S, g.H, 4u);
                if (g.smem <= kMaxSmem) { fits = true; break; }
                if (cww == 1 && g.P > 1) { --g.P; continue; }   // only the narrowest shape trades prefetch for history
                break;
            }
            if (fits) { g.ok = true; return g; }
        }
    }
    return g;
}

// 3..31 interleaved channels: thread = (run of R frames, channel), flat TMA tiles; windows up to 256 frames sum
// the whole runs one by one, longer ones through per-tile prefixes (as many history tiles as fit shared memory).
// float32: 16-frame runs.  int16, odd channel counts: 32-frame runs of 2-byte samples.  int16, even channel
// counts: the kernel sees C/2 "channels" of 32-bit words (channel pairs), 16-frame runs (g.pair, g.C = C/2,
// g.elem = 4).  int16 needs k >= 2 (k == 1, the identity, stays on the generic kernel).
StreamGeom plan_fewc(uint32_t k, uint32_t C, const mavg_tuning& tu, uint32_t elem = 4)
{
    StreamGeom g;
    g.NT = 512;
    g.mode = 4;
    g.C = C;
    if (C < 3 || C > 31 || k > 32768u) return g;     // int32 sums and the int16 divisions are exact up to here
    if (elem == 2 && k < 2) return g;
    const bool i16 = elem == 2;
    if (i16 && C % 2 == 0) {
        g.pair = true;
        C /= 2;
        elem = 4;
    }
    g.R = elem == 4 ? 16 : 32;
    g.C = C;
    g.elem = elem;
    const uint32_t R = (uint32_t)g.R;
    const uint32_t s = (R - k % R) % R;
    g.m_part = R - s;
    g.n_full = (k + s) / R - 1;
    uint32_t NR = 512 / C;
    while (NR > 0 && (NR * C) % 16 != 0) --NR;       // tile = whole 1024-byte swizzle atoms
    if (NR == 0) return g;
    g.runs = (int)NR;
    g.H = (int)((g.n_full + 1 + NR - 1) / NR);       // history tiles: the lag run lies at most H tiles back
    // whole runs of the window: summed one by one up to direct_max_k frames (at most 16 runs, this tile and the
    // previous one), through per-tile prefixes beyond
    g.long_mode = k > (tu.direct_max_k ? tu.direct_max_k : 256u) || g.n_full > 16 || g.H > 1;
    if (g.long_mode && (k <= 8 || g.n_full == 0)) g.long_mode = false;   // nothing to sum through prefixes
    if (!g.long_mode && g.H > 1) return g;
    g.ctas_per_sm = 1;
    g.P = tu.prefetch ? (int)tu.prefetch : 2;
    const uint32_t tile_bytes = NR * C * R * elem;
    for (;;) {
        g.S = g.H + 1 + g.P;
        g.smem = mavg::fewc_smem_bytes(tile_bytes, g.S, g.H, NR * C, g.pair ? 8u : 4u);
        if (g.smem <= kMaxSmem) break;
        if (g.P > 1) { --g.P; continue; }
        return g;
    }
    if (g.pair) {
        i16_mulhi_consts(k, &g.div_mul, &g.div_shift, &g.wscale);
    } else if (i16) {                                 // unsigned form on |w| (div_trunc_i32), exact for |w| < 2^31
        uint32_t lg = 0;
        while ((1u << lg) < k) ++lg;
        g.div_mul = (uint32_t)(((1ull << (31 + lg)) + k - 1) / k);
        g.div_shift = lg - 1;
    }
    g.ok = true;
    return g;
}

// float32 mono / planar with a window too long for the ring of plan_stream: the lag samples come back through a
// second TMA stream (stream_far_f32_kernel).  H = warm-up tiles = left context in whole tiles.
// int16 twin (stream_far_i16_kernel): mono / stereo / 4 / 6 / 8 / 12 / 16 interleaved channels, exact int32 window sums, so
// k * k < 2^31 (k <= 46 340; the multiply-high division is exact while 32768 k e < 2^(30 + ceil(log2 k)) with
// e <= k, i.e. while 2^15 k^2 <= 2^46).  Shapes: 384 x 32 for mono / stereo, 192 x 64 for 4 / 8 channels (24 KB tiles,
// one lag box, three lag stages), 224 x 72 for 6 channels (31.5 KB tiles, two lag stages).
StreamGeom plan_far_i16(uint32_t k, uint32_t C, const mavg_tuning& tu)
{
    StreamGeom g;
    g.mode = 5;
    g.elem = 2;
    g.C = C;
    if (k < 3 || (uint64_t)k * k >= (1ull << 31)) return g;
    if (C <= 2) { g.NT = 384; g.R = 32; }
    else if (C == 4 || C == 8 || C == 16) { g.NT = 192; g.R = 64; }   // 24 KB tiles, one lag box: two lag boxes in flight (8 ch, k = 19 200: 0.158 -> 0.136 ms against 256 x 64)
    else if (C == 6 || C == 12) { g.NT = 224; g.R = 72; }
    else return g;
    const uint64_t T = (uint64_t)g.NT * g.R;
    const uint64_t L = (uint64_t)k * C;
    if (L < T) return g;
    g.H = (int)((L + T - 1) / T);
    g.P = 2;
    g.S = g.P + 2;
    g.lag_stages = 3;                                  // two lag boxes in flight where 227 KB allow it (24 KB tiles)
    g.lag_chunks = (uint32_t)((L + 7) / 8);
    g.MIS = (int)(8ull * g.lag_chunks - L);
    g.ctas_per_sm = 1;
    for (;;) {
        g.smem = mavg::far_i16_smem_bytes(g.NT, g.R, g.S, g.lag_stages, (int)C);
        if (g.smem <= kMaxSmem) break;
        if (g.lag_stages > 2) { --g.lag_stages; continue; }
        return g;
    }
    i16_mulhi_consts(k, &g.div_mul, &g.div_shift, &g.wscale);
    g.ok = true;
    return g;
}

StreamGeom plan_far(uint32_t k, uint32_t C, const mavg_tuning& tu)
{
    StreamGeom g;
    // Tiles of 384 x 16 samples (24 KB; the lag samples of a tile are one 193-row box): results are written over the own
    // tile and stored from its ring stage, so 227 KB hold four own stages (two loading, one computing, one storing)
    // and three lag stages (two loading) -- twice the bytes in flight of the round-1 shape (512 x 16, staging tiles,
    // one own tile and one lag tile in flight), which was latency-bound.  tuning.threads = 512: tiles of 512 x 16.
    g.NT = tu.threads == 512 ? 512 : 384;
    g.R = 16;
    g.C = C;
    g.mode = 5;
    const uint64_t T = (uint64_t)g.NT * g.R;
    const uint64_t L = (uint64_t)k * C;               // lag distance in flat samples
    if (L < T || L > 0x40000000u || C > 2) return g;
    g.H = (int)((L + T - 1) / T);
    g.P = tu.prefetch ? (int)std::min<uint32_t>(tu.prefetch, 3u) : 2;
    g.S = g.P + 2;
    g.lag_stages = g.NT == 512 ? 2 : g.P + 1;         // lag boxes in flight: lag_stages - 1 (at least one)
    g.MIS = (int)((4 - L % 4) % 4);
    g.ctas_per_sm = 1;
    g.smem = mavg::far_smem_bytes(g.NT, g.R, g.S, g.lag_stages);
    g.ok = g.smem <= kMaxSmem;
    return g;
}

typedef void (*StreamKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const mavg::StreamParams);

// mavg_tuning.overlap -> StreamParams.pdl (0 off, 1 wait before the first load, 2 before the first store).
// MAVG_OVERLAP in the environment overrides the library default (for A/B measurements without a rebuild).
int pdl_mode(const mavg_tuning& tu)
{
    uint32_t o = tu.overlap;
    if (o == 0) {
        static const int env_default = [] {
            const char* e = getenv("MAVG_OVERLAP");
            const int v = e ? atoi(e) : 1;
            return (v >= 1 && v <= 3) ? v : 1;
        }();
        o = (uint32_t)env_default;
    }
    return o == 1 ? 1 : o == 2 ? 2 : 0;
}

// Launch of a TileRing kernel, with the programmatic-stream-serialization attribute when the plan overlaps launches:
// the kernel then starts while the previous kernel of the stream drains and orders itself with griddepcontrol.wait.
template <typename P>
cudaError_t launch_ring(void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const P), unsigned grid,
                        unsigned threads, uint32_t smem, cudaStream_t st, int pdl, const CUtensorMap& a,
                        const CUtensorMap& b, const CUtensorMap& c, const P& prm)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, b, c, prm);
}

template <int NT, int R, int C = 1, bool RMS = false>
StreamKernel pick_variant(int mis, int mode, uint32_t k)
{
    using namespace mavg;
    if (mode == 2) {
        switch (k) {
        case 1: return stream_f32_kernel<NT, R, 0, 2, 1, C, RMS>;
        case 2: return stream_f32_kernel<NT, R, 0, 2, 2, C, RMS>;
        case 3: return stream_f32_kernel<NT, R, 0, 2, 3, C, RMS>;
        case 4: return stream_f32_kernel<NT, R, 0, 2, 4, C, RMS>;
        case 5: return stream_f32_kernel<NT, R, 0, 2, 5, C, RMS>;
        case 6: return stream_f32_kernel<NT, R, 0, 2, 6, C, RMS>;
        case 7: return stream_f32_kernel<NT, R, 0, 2, 7, C, RMS>;
        default: return stream_f32_kernel<NT, R, 0, 2, 8, C, RMS>;
        }
    }
    if (mode == 0) {
        if (mis == 0) return stream_f32_kernel<NT, R, 0, 0, 0, C, RMS>;
        if (mis == 2) return stream_f32_kernel<NT, R, 2, 0, 0, C, RMS>;
        if constexpr (C == 1) {
            if (mis == 1) return stream_f32_kernel<NT, R, 1, 0, 0, C, RMS>;
            if (mis == 3) return stream_f32_kernel<NT, R, 3, 0, 0, C, RMS>;
        }
        return nullptr;
    }
    if (mis == 0) return stream_f32_kernel<NT, R, 0, 1, 0, C, RMS>;
    if (mis == 2) return stream_f32_kernel<NT, R, 2, 1, 0, C, RMS>;
    if constexpr (C == 1) {
        if (mis == 1) return stream_f32_kernel<NT, R, 1, 1, 0, C, RMS>;
        if (mis == 3) return stream_f32_kernel<NT, R, 3, 1, 0, C, RMS>;
    }
    return nullptr;
}

// moving RMS is built for the default kernel shape only (512 threads x 16 samples); plan_create rejects the tuning
// overrides of the shape for RMS plans
StreamKernel pick_kernel(const StreamGeom& g, uint32_t k, bool rms)
{
    if (rms) return g.C == 2 ? pick_variant<512, 16, 2, true>(g.MIS, g.mode, k) : pick_variant<512, 16, 1, true>(g.MIS, g.mode, k);
    if (g.C == 2) return pick_variant<512, 16, 2>(g.MIS, g.mode, k);
    if (g.NT == 256 && g.R == 16) return pick_variant<256, 16>(g.MIS, g.mode, k);
    if (g.NT == 256 && g.R == 32) return pick_variant<256, 32>(g.MIS, g.mode, k);
    return pick_variant<512, 16>(g.MIS, g.mode, k);
}

// lag misalignment MIS = (-k C) mod 8 samples: any value for odd C, even for C = 2 and 6, 0 / 4 for C = 4 and 12, 0 for C = 8 and 16
template <int NT, int R, int C>
StreamKernel pick_i16(int mis)
{
    using namespace mavg;
#define MAVG_I16_CASE(M) \
    case M: return (StreamKernel)stream_i16_kernel<NT, R, C, M>;
    switch (mis) {
        MAVG_I16_CASE(0)
    default: break;
    }
    if constexpr (C % 8 != 0) {
        switch (mis) {
            MAVG_I16_CASE(4)
        default: break;
        }
    }
    if constexpr (C % 4 != 0) {
        switch (mis) {
            MAVG_I16_CASE(2)
            MAVG_I16_CASE(6)
        default: break;
        }
    }
    if constexpr (C % 2 == 1) {
        switch (mis) {
            MAVG_I16_CASE(1)
            MAVG_I16_CASE(3)
            MAVG_I16_CASE(5)
            MAVG_I16_CASE(7)
        default: break;
        }
    }
#undef MAVG_I16_CASE
    return nullptr;
}

StreamKernel pick_i16_kernel(const StreamGeom& g)
{
    switch (g.C) {
    case 1: return g.NT == 256 ? pick_i16<256, 32, 1>(g.MIS) : pick_i16<512, 32, 1>(g.MIS);
    case 2: return g.NT == 256 ? pick_i16<256, 32, 2>(g.MIS) : pick_i16<512, 32, 2>(g.MIS);
    case 3: return g.NT == 224 ? pick_i16<224, 72, 3>(g.MIS) : g.NT == 128 ? pick_i16<128, 72, 3>(g.MIS) : pick_i16<512, 24, 3>(g.MIS);
    case 4:
        return g.NT == 256 ? pick_i16<256, 64, 4>(g.MIS) : g.NT == 128 ? pick_i16<128, 64, 4>(g.MIS) : pick_i16<512, 32, 4>(g.MIS);
    case 6: return g.NT == 224 ? pick_i16<224, 72, 6>(g.MIS) : g.NT == 128 ? pick_i16<128, 72, 6>(g.MIS) : pick_i16<512, 24, 6>(g.MIS);
    case 8:
        return g.NT == 256 ? pick_i16<256, 64, 8>(g.MIS) : g.NT == 128 ? pick_i16<128, 64, 8>(g.MIS) : pick_i16<512, 32, 8>(g.MIS);
    case 5: return g.NT == 192 ? pick_i16<192, 40, 5>(g.MIS) : pick_i16<384, 40, 5>(g.MIS);
    case 7: return g.NT == 128 ? pick_i16<128, 56, 7>(g.MIS) : pick_i16<256, 56, 7>(g.MIS);
    case 9: return g.NT == 128 ? pick_i16<128, 72, 9>(g.MIS) : pick_i16<224, 72, 9>(g.MIS);
    case 10: return g.NT == 192 ? pick_i16<192, 40, 10>(g.MIS) : pick_i16<384, 40, 10>(g.MIS);
    case 12: return g.NT == 128 ? pick_i16<128, 72, 12>(g.MIS) : pick_i16<224, 72, 12>(g.MIS);
    case 16: return g.NT == 128 ? pick_i16<128, 64, 16>(g.MIS) : pick_i16<256, 64, 16>(g.MIS);
    default: return nullptr;
    }
}

// ---------------------------------------------------------------------------------
// Plan
// ---------------------------------------------------------------------------------
struct DevCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, h2d done, compute done, d2h done
    // shard geometry
    uint64_t first_frame = 0;   // interleaved/mono: first frame of the shard (plan-relative)
    uint64_t frames = 0;        // frames in the shard
    uint32_t first_channel = 0; // planar: first channel of the shard
    uint32_t channels = 0;      // planar: channels in the shard; interleaved: all channels
    // owned buffers
    void* d_in = nullptr;
    void* d_out = nullptr;
    void* d_halo = nullptr;     // halo_frames * channels elements of left context (frame sharding)
    void* d_bsum = nullptr;     // generic path, long windows: 64-frame block sums
    void* d_scratch = nullptr;  // mavg_run_cascade intermediates
    void* d_sweep_in = nullptr;    // mavg_run_host_sweep: [largest halo of the sweep | shard], uploaded once
    size_t sweep_in_bytes = 0;
    void* d_far_stage = nullptr;   // far-lag kernel, context not contiguous with the shard: [context | first frames]
    size_t far_stage_bytes = 0;
    void* d_prefix = nullptr;      // prefix-difference path: 8-byte prefixes of the shard and of its left context
    size_t prefix_bytes = 0;
    cudaEvent_t ev_pass = nullptr;  // mavg_run_cascade: end of this device's latest pass
    size_t bsum_bytes = 0;
    int sm_count = 0;
    bool timed = false;
    // run_host pipeline: copies on their own streams so H2D, kernels and D2H overlap
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    // the few frames past the last whole 128-byte row run on a side stream, forked in front of the streaming kernel and
    // joined behind it (they depend on the input only): serialised behind the kernel they cost ~10 us per call
    cudaStream_t s_tail = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool tail_forked = false;
    std::vector<cudaEvent_t> pool;  // untimed events ordering slices across the three streams
};

}  // namespace

struct mavg_plan {
    mavg_desc desc;
    uint32_t path = MAVG_PATH_GENERIC;
    StreamGeom geom;
    uint64_t halo_frames = 0;
    std::vector<DevCtx> dev;
    mavg_timing timing = {0, 0, 0, 0};
    uint32_t launches_last_run = 0;
    bool peer_ok = false;
    bool timing_on = true;
    bool prefix_diff = false;   // generic-path plans with a far window and <= 8 interleaved channels: single-pass prefix + difference
};

namespace {

struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

bool planar_batch(const mavg_plan* p) { return p->desc.layout == MAVG_PLANAR && p->desc.channels > 1; }
// frames covered by one shared-memory tile of the stream kernel
uint64_t tile_frames(const mavg_plan* p)
{
    if (p->geom.mode == 3) return (uint64_t)(kColsNW / p->geom.cww) * kColsRF;
    if (p->geom.mode == 4) return (uint64_t)p->geom.runs * p->geom.R;
    return (uint64_t)p->geom.NT * p->geom.R / p->geom.C;
}
bool frame_sharded(const mavg_plan* p) { return !planar_batch(p); }

// samples per signal and signal count as the stream kernel sees one shard
void shard_signals(const mavg_plan* p, const DevCtx& d, uint64_t frames, uint64_t* n, uint64_t* signals,
                   uint64_t* stride_elems)
{
    if (planar_batch(p)) {
        *n = p->desc.frames;
        *signals = d.channels;
        *stride_elems = p->desc.frames;
    } else {
        *n = frames * p->desc.channels;
        *signals = 1;
        *stride_elems = *n;
    }
}

uint64_t shard_elems(const mavg_plan* p, const DevCtx& d)
{
    return planar_batch(p) ? (uint64_t)d.channels * p->desc.frames : d.frames * p->desc.channels;
}

int alloc_owned(mavg_plan* p, DevCtx& d)
{
    if (d.d_in && d.d_out) return MAVG_OK;
    const size_t bytes = std::max<size_t>(shard_elems(p, d) * elem_size(p->desc.dtype), 256);
    MAVG_CUDA(cudaSetDevice(d.device));
    if (!d.d_in && cudaMalloc(&d.d_in, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu input bytes failed on device %d", bytes, d.device);
    }
    if (!d.d_out && cudaMalloc(&d.d_out, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu output bytes failed on device %d", bytes, d.device);
    }
    return MAVG_OK;
}

int alloc_halo(mavg_plan* p, DevCtx& d)
{
    if (d.d_halo || p->halo_frames == 0) return MAVG_OK;
    const size_t bytes = p->halo_frames * p->desc.channels * elem_size(p->desc.dtype);
    MAVG_CUDA(cudaSetDevice(d.device));
    if (cudaMalloc(&d.d_halo, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu halo bytes failed on device %d", bytes, d.device);
    }
    return MAVG_OK;
}

template <typename T>
int launch_generic_t(const mavg_plan* p, DevCtx& d, const T* in, T* out, const T* halo, uint64_t frames,
                     uint64_t out_begin, uint64_t out_end, uint32_t* launches)
{
    if (out_begin >= out_end) return MAVG_OK;
    constexpr int RG = 64;
    mavg::GenericParams gp;
    uint32_t signals = 1;
    if (planar_batch(p)) {
        gp.frames = p->desc.frames;
        gp.channels = 1;
        gp.sig_stride = p->desc.frames;
        signals = d.channels;
    } else {
        gp.frames = frames;
        gp.channels = p->desc.channels;
        gp.sig_stride = 0;
    }
    gp.out_begin = out_begin;
    gp.out_end = out_end;
    gp.halo_frames = halo ? p->halo_frames : 0;
    gp.k = p->desc.window;
    gp.bsum = nullptr;
    gp.nblk = 0;
    gp.rms = p->desc.op == MAVG_OP_RMS ? 1u : 0u;
    gp.pad_ = 0;
    // long windows: one extra pass builds RG-frame block sums so each run starts from k/RG table entries
    if (gp.k >= 256 && out_begin == 0 && gp.frames >= 4 * (uint64_t)RG) {
        typedef typename mavg::GenericAcc<T>::type Acc;
        const uint64_t nblk = (gp.frames + RG - 1) / RG;
        const size_t need = (size_t)nblk * gp.channels * signals * sizeof(Acc);
        if (d.bsum_bytes < need) {
            if (d.d_bsum) MAVG_CUDA(cudaFree(d.d_bsum));
            d.d_bsum = nullptr;
            d.bsum_bytes = 0;
            if (cudaMalloc(&d.d_bsum, need) != cudaSuccess) {
                cudaGetLastError();
                return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu block-sum bytes failed", need);
            }
            d.bsum_bytes = need;
        }
        const uint64_t bthreads = nblk * gp.channels;
        const uint64_t bblocks = (bthreads + 255) / 256;
        if (bblocks <= 0x7fffffffull && signals <= 65535u) {
            dim3 bgrid((unsigned)bblocks, signals, 1);
            if (gp.rms)
                mavg::block_sums_kernel<T, RG, true><<<bgrid, 256, 0, d.stream>>>(in, (Acc*)d.d_bsum, gp.frames, gp.sig_stride,
                                                                                gp.channels, nblk);
            else
                mavg::block_sums_kernel<T, RG, false><<<bgrid, 256, 0, d.stream>>>(in, (Acc*)d.d_bsum, gp.frames, gp.sig_stride,
                                                                                 gp.channels, nblk);
            MAVG_CUDA(cudaGetLastError());
            ++*launches;
            gp.bsum = d.d_bsum;
            gp.nblk = nblk;
        }
    }
    const uint64_t runs = (out_end - out_begin + RG - 1) / RG;
    const uint64_t threads = runs * gp.channels;
    const uint64_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffull) return fail(MAVG_ERR_UNSUPPORTED, "generic path: signal too long");
    // gridDim.y <= 65535: walk planar signals in slabs
    for (uint32_t s0 = 0; s0 < signals; s0 += 65535u) {
        const uint32_t ns = std::min<uint32_t>(65535u, signals - s0);
        dim3 grid((unsigned)blocks, ns, 1);
        const T* in_s = in + (uint64_t)s0 * gp.sig_stride;
        T* out_s = out + (uint64_t)s0 * gp.sig_stride;
        if (gp.rms)
            mavg::generic_kernel<T, RG, false, true><<<grid, 256, 0, d.stream>>>(in_s, out_s, halo, gp);
        else if (std::is_same<T, float>::value && gp.k >= 9)
            mavg::generic_kernel<T, RG, std::is_same<T, float>::value><<<grid, 256, 0, d.stream>>>(in_s, out_s, halo, gp);
        else
            mavg::generic_kernel<T, RG, false><<<grid, 256, 0, d.stream>>>(in_s, out_s, halo, gp);
        MAVG_CUDA(cudaGetLastError());
        ++*launches;
    }
    return MAVG_OK;
}

// The frames a streaming kernel leaves behind (a fraction of a 128-byte row): tail_kernel, one CTA per signal.
template <typename T>
int launch_tail_t(const mavg_plan* p, DevCtx& d, const T* in, T* out, const T* halo, uint64_t frames, uint64_t out_begin,
                  uint64_t out_end, uint32_t* launches)
{
    mavg::GenericParams gp;
    memset(&gp, 0, sizeof gp);
    uint32_t signals = 1;
    if (planar_batch(p)) {
        gp.frames = p->desc.frames;
        gp.channels = 1;
        gp.sig_stride = p->desc.frames;
        signals = d.channels;
    } else {
        gp.frames = frames;
        gp.channels = p->desc.channels;
    }
    gp.out_begin = out_begin;
    gp.out_end = out_end;
    gp.halo_frames = halo ? p->halo_frames : 0;
    gp.k = p->desc.window;
    gp.rms = p->desc.op == MAVG_OP_RMS ? 1u : 0u;
    if (gp.rms) mavg::tail_kernel<T, true><<<signals, 256, 0, d.stream>>>(in, out, halo, gp);
    else mavg::tail_kernel<T, false><<<signals, 256, 0, d.stream>>>(in, out, halo, gp);
    MAVG_CUDA(cudaGetLastError());
    ++*launches;
    return MAVG_OK;
}

int launch_generic(const mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                   uint64_t out_begin, uint64_t out_end, uint32_t* launches);

int launch_tail(const mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                uint64_t out_begin, uint64_t out_end, uint32_t* launches)
{
    if (out_begin >= out_end) return MAVG_OK;
    if (out_end - out_begin > 128) return launch_generic(p, d, in, out, halo, frames, out_begin, out_end, launches);
    if (p->desc.dtype == MAVG_F32)
        return launch_tail_t<float>(p, d, (const float*)in, (float*)out, (const float*)halo, frames, out_begin, out_end, launches);
    return launch_tail_t<int16_t>(p, d, (const int16_t*)in, (int16_t*)out, (const int16_t*)halo, frames, out_begin, out_end,
                                  launches);
}

int launch_generic(const mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                   uint64_t out_begin, uint64_t out_end, uint32_t* launches)
{
    if (p->desc.dtype == MAVG_F32)
        return launch_generic_t<float>(p, d, (const float*)in, (float*)out, (const float*)halo, frames, out_begin,
                                       out_end, launches);
    return launch_generic_t<int16_t>(p, d, (const int16_t*)in, (int16_t*)out, (const int16_t*)halo, frames,
                                     out_begin, out_end, launches);
}

// Tail of a streamed shard (frames [out_begin, frames): less than one 128-byte row): enqueued on the device's side
// stream BEFORE the streaming kernel is launched, so the two run concurrently; tail_join() makes the main stream wait
// for it.  Capture-safe (fork and join by events).  Longer remainders stay on the main stream (launch_tail).
int tail_fork(const mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
              uint64_t out_begin, uint32_t* launches)
{
    d.tail_forked = false;
    if (out_begin >= frames) return MAVG_OK;
    if (frames - out_begin > 128) return MAVG_OK;        // tail_join launches it behind the kernel
    // A frame that straddles the last whole row is written by both kernels: the same bits for int16 (exact), but a
    // float32 tail rounds differently from the streaming kernel, so there the tail keeps running behind it (its value wins)
    if (p->desc.dtype == MAVG_F32 && !planar_batch(p) && (frames * p->desc.channels / 32 * 32) % p->desc.channels != 0)
        return MAVG_OK;
    if (!d.s_tail) {
        MAVG_CUDA(cudaStreamCreateWithFlags(&d.s_tail, cudaStreamNonBlocking));
        MAVG_CUDA(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
        MAVG_CUDA(cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming));
    }
    MAVG_CUDA(cudaEventRecord(d.ev_fork, d.stream));
    MAVG_CUDA(cudaStreamWaitEvent(d.s_tail, d.ev_fork, 0));
    cudaStream_t main_stream = d.stream;
    d.stream = d.s_tail;
    const int rc = launch_tail(p, d, in, out, halo, frames, out_begin, frames, launches);
    d.stream = main_stream;
    MAVG_TRY(rc);
    MAVG_CUDA(cudaEventRecord(d.ev_join, d.s_tail));
    d.tail_forked = true;
    return MAVG_OK;
}

int tail_join(const mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
              uint64_t out_begin, uint32_t* launches)
{
    if (d.tail_forked) {
        d.tail_forked = false;
        MAVG_CUDA(cudaStreamWaitEvent(d.stream, d.ev_join, 0));
        return MAVG_OK;
    }
    return launch_tail(p, d, in, out, halo, frames, out_begin, frames, launches);
}

bool stream_eligible(const mavg_plan* p, const DevCtx& d, const void* in, const void* out, const void* halo,
                     uint64_t frames)
{
    if (p->path != MAVG_PATH_STREAM || !p->geom.ok) return false;
    if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)halo) & 15u) return false;
    uint64_t n, signals, stride;
    shard_signals(p, d, frames, &n, &signals, &stride);
    const uint64_t row = 128 / p->geom.elem;
    if (n < row) return false;
    if (signals > 1 && (stride * p->geom.elem % 16) != 0) return false;  // tensor-map strides: multiples of 16 bytes
    if (n / row > 0x7fffffffull - 65536 || signals > 0x7fffffffull) return false;
    return true;
}

int make_map_2d(CUtensorMap* map, const void* base, uint64_t channels, uint64_t frames, uint32_t tile_frames_,
                uint32_t tile_channels)
{
    EncodeTiledFn enc = nullptr;
    MAVG_TRY(get_encoder(&enc));
    cuuint64_t dims[2] = {channels, frames};
    cuuint64_t strides[1] = {channels * 4};
    cuuint32_t box[2] = {tile_channels, tile_frames_};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(MAVG_ERR_DRIVER, "cuTensorMapEncodeTiled (2-D) failed (%d) channels=%llu frames=%llu", (int)r,
                    (unsigned long long)channels, (unsigned long long)frames);
    return MAVG_OK;
}

// many-channel interleaved float32 through the column kernel
int launch_cols(mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                uint32_t* launches)
{
    const StreamGeom& g = p->geom;
    const uint32_t C = g.pair ? g.C : p->desc.channels;  // columns as the kernel sees them (int16: channel-pair words)
    const uint32_t FT = (kColsNW / g.cww) * kColsRF;
    const uint32_t CW = 32u * g.cww;
    CUtensorMap in_map, halo_map;
    MAVG_TRY(make_map_2d(&in_map, in, C, frames, FT, CW));
    if (halo) MAVG_TRY(make_map_2d(&halo_map, halo, C, (uint64_t)g.H * FT, FT, CW));
    else halo_map = in_map;
    mavg::ColsParams cp;
    cp.inv_k = 1.0f / (float)p->desc.window;
    cp.k = p->desc.window;
    cp.n_full = g.n_full;
    cp.m_part = g.m_part;
    cp.channels = C;
    cp.frames = frames;
    cp.col_blocks = (int32_t)((C + CW - 1) / CW);
    const uint64_t tiles = (frames + FT - 1) / FT;
    cp.tiles_per_col = (int32_t)tiles;
    const uint64_t ctas = (uint64_t)d.sm_count;
    // tile ranges per CTA: eight on long signals (the tail of the grid stays short), fewer when that would make a range
    // shorter than 64 tiles -- every range replays H history tiles and restarts the load pipeline (measured on 2^27
    // int16 samples: ranges of 7 tiles cost 11 % extra reads and a fifth of the time waiting for the first tiles)
    const uint64_t tiles_all = (frames + FT - 1) / FT * ((C + CW - 1) / CW);
    const uint64_t per_cta = p->desc.tuning.chunks_per_cta
                                 ? p->desc.tuning.chunks_per_cta
                                 : std::max<uint64_t>(1, std::min<uint64_t>(8, tiles_all / ((uint64_t)d.sm_count * 64)));
    uint64_t cps = std::min<uint64_t>(tiles, (ctas * per_cta + cp.col_blocks - 1) / cp.col_blocks);
    uint64_t chunk_tiles = (tiles + cps - 1) / cps;
    cps = (tiles + chunk_tiles - 1) / chunk_tiles;
    if (cps * cp.col_blocks > 0x7fffffffull || tiles > 0x7fffffffull / FT)
        return fail(MAVG_ERR_UNSUPPORTED, "signal too long for the column kernel");
    cp.chunk_tiles = (int32_t)chunk_tiles;
    cp.chunks_per_col = (int32_t)cps;
    cp.total_chunks = (int32_t)(cps * cp.col_blocks);
    cp.hist_tiles = g.H;
    cp.stages = g.S;
    cp.prefetch = g.P;
    cp.has_halo = halo ? 1 : 0;
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, (uint64_t)cp.total_chunks);
    if (g.pair) {
        typedef void (*Kern16)(const CUtensorMap, const CUtensorMap, const CUtensorMap, uint32_t*, const mavg::ColsParams,
                               const uint32_t, const uint32_t, const uint32_t);
#define MAVG_COLS16(NW_, RF_, TM_)                                                 \
    (g.cww == 8   ? (Kern16)mavg::stream_cols_i16x2_kernel<NW_, RF_, 8, TM_>       \
     : g.cww == 4 ? (Kern16)mavg::stream_cols_i16x2_kernel<NW_, RF_, 4, TM_>       \
     : g.cww == 2 ? (Kern16)mavg::stream_cols_i16x2_kernel<NW_, RF_, 2, TM_>       \
                  : (Kern16)mavg::stream_cols_i16x2_kernel<NW_, RF_, 1, TM_>)
        Kern16 kern16 = g.NT == 256 ? (g.tmast ? MAVG_COLS16(8, 32, true) : MAVG_COLS16(8, 32, false))
                                    : (g.tmast ? MAVG_COLS16(kColsNW, kColsRF, true) : MAVG_COLS16(kColsNW, kColsRF, false));
#undef MAVG_COLS16
        CUtensorMap out_map = in_map;
        if (g.tmast) MAVG_TRY(make_map_2d(&out_map, out, C, frames, FT, CW));
        MAVG_CUDA(cudaFuncSetAttribute(kern16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        kern16<<<grid, g.NT, g.smem, d.stream>>>(in_map, halo_map, out_map, (uint32_t*)out, cp, g.div_mul, g.div_shift, g.wscale);
        MAVG_CUDA(cudaGetLastError());
        ++*launches;
        return MAVG_OK;
    }
    void (*kern)(const CUtensorMap, const CUtensorMap, float*, const mavg::ColsParams) =
        g.cww == 8   ? mavg::stream_cols_f32_kernel<kColsNW, kColsRF, 8>
        : g.cww == 4 ? mavg::stream_cols_f32_kernel<kColsNW, kColsRF, 4>
        : g.cww == 2 ? mavg::stream_cols_f32_kernel<kColsNW, kColsRF, 2>
                     : mavg::stream_cols_f32_kernel<kColsNW, kColsRF, 1>;
    MAVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    kern<<<grid, kColsNW * 32, g.smem, d.stream>>>(in_map, halo_map, (float*)out, cp);
    MAVG_CUDA(cudaGetLastError());
    ++*launches;
    return MAVG_OK;
}

// 3..31 interleaved float32 channels through the few-channel kernel
int launch_fewc(mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                uint32_t* launches)
{
    const StreamGeom& g = p->geom;
    const uint32_t C = g.C;                              // channels as the kernel sees them (pairs for g.pair)
    const uint64_t n = frames * C;                       // flat kernel elements (samples, or words for g.pair)
    const uint64_t row = 128 / g.elem;                   // elements per 128-byte row
    const uint64_t rows = n / row;
    const uint64_t tile_floats = (uint64_t)g.runs * C * g.R;   // samples per tile
    const uint32_t tile_rows = (uint32_t)(tile_floats / row);
    CUtensorMap in_map, out_map, halo_map;
    MAVG_TRY(make_map(&in_map, in, rows, 1, rows * 128, tile_rows, g.elem));
    MAVG_TRY(make_map(&out_map, out, rows, 1, rows * 128, tile_rows, g.elem));
    if (halo) MAVG_TRY(make_map(&halo_map, halo, (uint64_t)g.H * tile_rows, 1, (uint64_t)g.H * tile_rows * 128, tile_rows, g.elem));
    else halo_map = in_map;
    mavg::FewcParams fp;
    memset(&fp, 0, sizeof fp);
    mavg::StreamParams& sp = fp.sp;
    sp.inv_k = 1.0f / (float)p->desc.window;
    sp.k = p->desc.window;
    sp.n_full = g.n_full;
    sp.m_part = g.m_part;
    sp.div_mul = g.div_mul;
    sp.div_shift = g.div_shift;
    sp.wscale = g.wscale;
    const uint64_t tiles = (rows * row + tile_floats - 1) / tile_floats;
    sp.tiles_per_signal = (int32_t)tiles;
    const uint64_t ctas = (uint64_t)d.sm_count;
    uint64_t cps = std::min<uint64_t>(tiles, ctas * std::max<uint32_t>(1u, p->desc.tuning.chunks_per_cta));
    uint64_t chunk_tiles = (tiles + cps - 1) / cps;
    cps = (tiles + chunk_tiles - 1) / chunk_tiles;
    sp.chunk_tiles = (int32_t)chunk_tiles;
    sp.chunks_per_signal = (int32_t)cps;
    sp.total_chunks = (int32_t)cps;
    sp.hist_tiles = g.H;
    sp.stages = g.S;
    sp.prefetch = g.P;
    sp.has_halo = halo ? 1 : 0;
    sp.pdl = pdl_mode(p->desc.tuning);
    fp.channels = C;
    fp.runs = (uint32_t)g.runs;
    fp.long_mode = g.long_mode ? 1u : 0u;
    void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const mavg::FewcParams) =
        g.pair ? mavg::stream_fewc_i16x2_kernel<16>
               : g.elem == 4 ? mavg::stream_fewc_f32_kernel<16> : mavg::stream_fewc_i16_kernel<32>;
    MAVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, cps);
    // flat samples past the last whole 128-byte row: the frames that touch them go to the tail kernel (side stream)
    if (rows * row < n) MAVG_TRY(tail_fork(p, d, in, out, halo, frames, rows * row / C, launches));
    MAVG_CUDA(launch_ring(kern, grid, 512, g.smem, d.stream, sp.pdl, in_map, out_map, halo_map, fp));
    ++*launches;
    if (rows * row < n) MAVG_TRY(tail_join(p, d, in, out, halo, frames, rows * row / C, launches));
    return MAVG_OK;
}

// float32 mono / planar, window longer than the ring: far-lag kernel.  `halo`, when given, lies directly in front of `in`.
int launch_far(mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames, uint32_t* launches)
{
    const StreamGeom& g = p->geom;
    uint64_t n, signals, stride;
    shard_signals(p, d, frames, &n, &signals, &stride);
    const uint32_t es = g.elem;                         // 4: float32 kernel, 2: int16 twin
    const uint32_t row = 128 / es;                      // samples per 128-byte row
    const uint64_t rows = n / row;
    const uint64_t T = (uint64_t)g.NT * g.R;
    const uint32_t tile_rows = (uint32_t)(T / row);
    const uint64_t row_base = halo ? (uint64_t)g.H * tile_rows : 0;
    const void* base = halo ? halo : in;
    const int swz = es == 2 ? mavg::i16_swizzle_bytes(g.R) : 128;
    CUtensorMap in_map, out_map, lag_map;
    MAVG_TRY(make_map(&in_map, base, rows + row_base, signals, stride * es, tile_rows, es, swz));
    MAVG_TRY(make_map(&lag_map, base, rows + row_base, signals, stride * es,
                      tile_rows / mavg::far_lag_boxes((int)tile_rows) + 1, es, swz));
    MAVG_TRY(make_map(&out_map, out, rows, signals, stride * es, tile_rows, es, swz));
    mavg::FarParams fp;
    memset(&fp, 0, sizeof fp);
    mavg::StreamParams& sp = fp.sp;
    const uint32_t k = (uint32_t)((uint64_t)p->desc.window * g.C);   // lag distance in flat samples
    sp.inv_k = 1.0f / (float)p->desc.window;
    sp.k = es == 2 ? p->desc.window : k;                // the int16 kernel multiplies by its channel count itself
    sp.div_mul = g.div_mul;
    sp.div_shift = g.div_shift;
    sp.wscale = g.wscale;
    const uint64_t tiles = (rows * row + T - 1) / T;
    sp.tiles_per_signal = (int32_t)tiles;
    const uint64_t ctas = (uint64_t)d.sm_count;
    uint64_t cps = std::max<uint64_t>(1, ctas * std::max<uint32_t>(1u, p->desc.tuning.chunks_per_cta) / signals);
    cps = std::min<uint64_t>(cps, tiles);
    uint64_t chunk_tiles = (tiles + cps - 1) / cps;
    cps = (tiles + chunk_tiles - 1) / chunk_tiles;
    if (cps * signals > 0x7fffffffull) return fail(MAVG_ERR_UNSUPPORTED, "too many tile ranges");
    sp.chunk_tiles = (int32_t)chunk_tiles;
    sp.chunks_per_signal = (int32_t)cps;
    sp.total_chunks = (int32_t)(cps * signals);
    sp.hist_tiles = 0;
    sp.stages = g.S;
    sp.prefetch = g.P;
    sp.has_halo = 0;
    fp.warm_tiles = g.H;
    fp.row_base = (int32_t)row_base;
    fp.koff = (row - k % row) % row;
    fp.lag_rows = (int32_t)((k + fp.koff) / row);
    fp.lag_stages = g.lag_stages;
    fp.lag_prefetch = std::max(1, g.lag_stages - 1);
    {
        static const int hints = [] {
            const char* e = getenv("MAVG_FAR_HINTS");
            const int v = e ? atoi(e) : 2;
            return (v >= 0 && v <= 2) ? v : 2;
        }();
        fp.hints = hints;   // measured at k = 60 000 (2^27 samples): DRAM reads 899 -> 645 -> 643 MB for 537 MB of input
    }
    typedef void (*FarKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const mavg::FarParams);
#define MAVG_FAR(NT_)                                                                                                     \
    (g.C == 2 ? (g.MIS == 0 ? (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 0, 2> : (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 2, 2>) \
     : g.MIS == 0 ? (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 0>                                                    \
     : g.MIS == 1 ? (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 1>                                                    \
     : g.MIS == 2 ? (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 2>                                                    \
                  : (FarKernel)mavg::stream_far_f32_kernel<NT_, 16, 3>)
    FarKernel kern = g.NT == 512 ? MAVG_FAR(512) : MAVG_FAR(384);
#undef MAVG_FAR
    if (es == 2) {
        kern = nullptr;
#define MAVG_FAR16(NT_, R_, C_, M_) \
    if (g.C == C_ && g.MIS == M_ && g.NT == NT_) kern = (FarKernel)mavg::stream_far_i16_kernel<NT_, R_, C_, M_>;
        MAVG_FAR16(384, 32, 1, 0) MAVG_FAR16(384, 32, 1, 1) MAVG_FAR16(384, 32, 1, 2) MAVG_FAR16(384, 32, 1, 3)
        MAVG_FAR16(384, 32, 1, 4) MAVG_FAR16(384, 32, 1, 5) MAVG_FAR16(384, 32, 1, 6) MAVG_FAR16(384, 32, 1, 7)
        MAVG_FAR16(384, 32, 2, 0) MAVG_FAR16(384, 32, 2, 2) MAVG_FAR16(384, 32, 2, 4) MAVG_FAR16(384, 32, 2, 6)
        MAVG_FAR16(192, 64, 4, 0) MAVG_FAR16(192, 64, 4, 4)
        MAVG_FAR16(224, 72, 6, 0) MAVG_FAR16(224, 72, 6, 2) MAVG_FAR16(224, 72, 6, 4) MAVG_FAR16(224, 72, 6, 6)
        MAVG_FAR16(192, 64, 8, 0)
        MAVG_FAR16(224, 72, 12, 0) MAVG_FAR16(224, 72, 12, 4) MAVG_FAR16(192, 64, 16, 0)
#undef MAVG_FAR16
        if (!kern) return fail(MAVG_ERR_UNSUPPORTED, "no far-lag int16 kernel for %u channels, misalignment %d", g.C, g.MIS);
    }
    MAVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, (uint64_t)sp.total_chunks);
    if (rows * row < n) MAVG_TRY(tail_fork(p, d, in, out, halo, frames, rows * row / g.C, launches));
    kern<<<grid, g.NT, g.smem, d.stream>>>(in_map, out_map, lag_map, fp);
    MAVG_CUDA(cudaGetLastError());
    ++*launches;
    if (rows * row < n) MAVG_TRY(tail_join(p, d, in, out, halo, frames, rows * row / g.C, launches));
    return MAVG_OK;
}

// Far windows no streaming kernel takes (int16 beyond the ring or beyond 32768 frames, 3..8 channels beyond their
// history): single-pass prefix sum of the shard (mavg_prefix_sum's kernel) and of its left context, then one
// difference pass -- the reference's scan binaries, exact, at a cost that does not depend on k (the generic
// kernel's run starts grow with the window: 55 Gsamples/s for stereo int16 at k = 40 000).
int launch_prefix_diff(mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                       uint32_t* launches)
{
    const uint32_t C = p->desc.channels;
    const uint64_t hf = halo ? p->halo_frames : 0;
    const size_t need = (size_t)(frames + hf) * C * 8;
    if (d.prefix_bytes < need) {
        if (d.d_prefix) MAVG_CUDA(cudaFree(d.d_prefix));
        d.d_prefix = nullptr;
        d.prefix_bytes = 0;
        if (cudaMalloc(&d.d_prefix, need) != cudaSuccess) {
            cudaGetLastError();
            return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu prefix bytes failed on device %d", need, d.device);
        }
        d.prefix_bytes = need;
    }
    char* P = (char*)d.d_prefix;
    char* HP = P + (size_t)frames * C * 8;
    MAVG_TRY(mavg_prefix_sum((int)p->desc.dtype, in, P, frames, C, d.stream));
    ++*launches;
    if (hf) {
        MAVG_TRY(mavg_prefix_sum((int)p->desc.dtype, halo, HP, hf, C, d.stream));
        ++*launches;
    }
    const unsigned blocks = (unsigned)std::min<uint64_t>((frames + 255) / 256, (uint64_t)d.sm_count * 16);
    if (p->desc.dtype == MAVG_F32)
        mavg::prefix_diff_kernel<float, double><<<blocks, 256, 0, d.stream>>>((const double*)P, hf ? (const double*)HP : nullptr,
                                                                             (float*)out, frames, C, p->desc.window, hf);
    else
        mavg::prefix_diff_kernel<int16_t, long long><<<blocks, 256, 0, d.stream>>>(
            (const long long*)P, hf ? (const long long*)HP : nullptr, (int16_t*)out, frames, C, p->desc.window, hf);
    MAVG_CUDA(cudaGetLastError());
    ++*launches;
    return MAVG_OK;
}

// Enqueue the kernels for `frames` frames (a whole shard, or one slice of it whose left context
// is `halo`) on the device stream.  Planar batches always run whole (frames = desc.frames).
int launch_shard(mavg_plan* p, DevCtx& d, const void* in, void* out, const void* halo, uint64_t frames,
                 uint32_t* launches)
{
    MAVG_CUDA(cudaSetDevice(d.device));
    if (planar_batch(p)) frames = p->desc.frames;
    if (frames == 0 || (planar_batch(p) && d.channels == 0)) return MAVG_OK;
    if (p->path == MAVG_PATH_STREAM && p->geom.ok && p->geom.mode == 4) {
        const uint64_t nflat = frames * p->desc.channels;
        if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)halo) & 15u) == 0 && nflat >= 64 && nflat / 32 < (1ull << 31) - 65536)
            return launch_fewc(p, d, in, out, halo, frames, launches);
        return launch_generic(p, d, in, out, halo, frames, 0, frames, launches);
    }
    if (p->path == MAVG_PATH_STREAM && p->geom.ok && p->geom.mode == 5) {
        // the left context has to sit directly in front of the shard (run_host slices, contiguous callers);
        // a context somewhere else (a peer's tail) is served by the generic kernel
        const size_t fes = p->geom.elem;                  // 4: float32, 2: int16 twin
        const size_t hb = (size_t)p->halo_frames * (planar_batch(p) ? 1 : p->desc.channels) * fes;
        const bool contiguous = halo == nullptr || (const char*)halo + hb == (const char*)in;
        const uint64_t frow = 128 / fes;
        const uint64_t rows_all = (planar_batch(p) ? p->desc.frames : frames * p->desc.channels) / frow +
                                  (uint64_t)p->geom.H * (p->geom.NT * p->geom.R / frow);
        if (contiguous && stream_eligible(p, d, in, out, halo, frames) && rows_all < 0x7fffffffull - 65536)
            return launch_far(p, d, in, out, halo, frames, launches);
        if (!contiguous && !planar_batch(p) && stream_eligible(p, d, in, out, halo, frames) &&
            rows_all < 0x7fffffffull - 65536) {
            // The left context lives somewhere else (a peer's tail over NVLink, a staged halo).  Only the first
            // halo_frames frames of the shard can see it: [context | those frames] is copied into a small plan-owned
            // buffer (two times the window, a few MB) where the context IS contiguous, and filtered from there straight
            // into the output; the rest of the shard finds its context inside the shard itself.
            const size_t fb = (size_t)p->desc.channels * fes;
            const uint64_t fa = std::min<uint64_t>(frames, p->halo_frames);
            const size_t need = hb + fa * fb;
            if (d.far_stage_bytes < need) {
                if (d.d_far_stage) MAVG_CUDA(cudaFree(d.d_far_stage));
                d.d_far_stage = nullptr;
                d.far_stage_bytes = 0;
                if (cudaMalloc(&d.d_far_stage, need) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu far-lag staging bytes failed", need);
                }
                d.far_stage_bytes = need;
            }
            char* stage = (char*)d.d_far_stage;
            MAVG_CUDA(cudaMemcpyAsync(stage, halo, hb, cudaMemcpyDefault, d.stream));
            MAVG_CUDA(cudaMemcpyAsync(stage + hb, in, fa * fb, cudaMemcpyDeviceToDevice, d.stream));
            MAVG_TRY(launch_far(p, d, stage + hb, out, stage, fa, launches));
            if (frames > fa)
                MAVG_TRY(launch_far(p, d, (const char*)in + fa * fb, (char*)out + fa * fb, (const char*)in + fa * fb - hb,
                                    frames - fa, launches));
            return MAVG_OK;
        }
        return launch_generic(p, d, in, out, halo, frames, 0, frames, launches);
    }
    if (p->path == MAVG_PATH_STREAM && p->geom.ok && p->geom.mode == 3) {
        if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)halo) & 15u) == 0 && frames < (1ull << 31))
            return launch_cols(p, d, in, out, halo, frames, launches);
        return launch_generic(p, d, in, out, halo, frames, 0, frames, launches);
    }
    if (p->path == MAVG_PATH_GENERIC && p->prefix_diff && frames >= 64)
        return launch_prefix_diff(p, d, in, out, halo, frames, launches);
    if (!stream_eligible(p, d, in, out, halo, frames))
        return launch_generic(p, d, in, out, halo, frames, 0, frames, launches);
    const StreamGeom& g = p->geom;
    uint64_t n, signals, stride;
    shard_signals(p, d, frames, &n, &signals, &stride);
    const uint64_t row = 128 / g.elem;                 // samples per 128-byte row
    const uint64_t rows = n / row;
    const uint32_t tile_rows = (uint32_t)(g.NT * g.R / row);
    const uint64_t T = (uint64_t)g.NT * g.R;

    CUtensorMap in_map, out_map, halo_map;
    // int16: the swizzle under which a warp's 128-bit accesses to consecutive runs of R samples do not collide
    const int swz = g.elem == 2 ? mavg::i16_swizzle_bytes(g.R) : 128;
    MAVG_TRY(make_map(&in_map, in, rows, signals, stride * g.elem, tile_rows, g.elem, swz));
    MAVG_TRY(make_map(&out_map, out, rows, signals, stride * g.elem, tile_rows, g.elem, swz));
    const uint64_t halo_rows = (uint64_t)g.H * tile_rows;
    if (halo) MAVG_TRY(make_map(&halo_map, halo, halo_rows, 1, halo_rows * 128, tile_rows, g.elem, swz));
    else halo_map = in_map;

    mavg::StreamParams sp;
    sp.inv_k = 1.0f / (float)p->desc.window;
    sp.k = p->desc.window;
    sp.n_full = g.n_full;
    sp.m_part = g.m_part;
    sp.lag_chunks = g.lag_chunks;
    sp.div_mul = g.div_mul;
    sp.div_shift = g.div_shift;
    sp.wscale = g.wscale;
    memset(sp.wtab, 0, sizeof sp.wtab);
    const uint64_t tiles = (rows * row + T - 1) / T;
    sp.tiles_per_signal = (int32_t)tiles;
    const uint64_t ctas = (uint64_t)d.sm_count * g.ctas_per_sm;
    const uint64_t want_chunks = ctas * std::max<uint32_t>(1u, p->desc.tuning.chunks_per_cta);
    uint64_t cps = std::max<uint64_t>(1, want_chunks / signals);
    cps = std::min<uint64_t>(cps, tiles);
    uint64_t chunk_tiles = (tiles + cps - 1) / cps;
    cps = (tiles + chunk_tiles - 1) / chunk_tiles;
    sp.chunk_tiles = (int32_t)chunk_tiles;
    sp.chunks_per_signal = (int32_t)cps;
    if (cps * signals > 0x7fffffffull) return fail(MAVG_ERR_UNSUPPORTED, "too many tile ranges");
    sp.total_chunks = (int32_t)(cps * signals);
    sp.hist_tiles = g.H;
    sp.stages = g.S;
    sp.prefetch = g.P;
    sp.has_halo = halo ? 1 : 0;
    sp.pdl = pdl_mode(p->desc.tuning);

    StreamKernel kern = g.elem == 4 ? pick_kernel(g, p->desc.window, p->desc.op == MAVG_OP_RMS) : pick_i16_kernel(g);
    if (!kern) return fail(MAVG_ERR_UNSUPPORTED, "no stream kernel variant for this window");
    MAVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(ctas, (uint64_t)sp.total_chunks);
    // samples past the last whole 128-byte row (per signal): tail kernel, forked onto the side stream
    if (rows * row < n) MAVG_TRY(tail_fork(p, d, in, out, halo, frames, rows * row / g.C, launches));
    MAVG_CUDA(launch_ring(kern, grid, (unsigned)g.NT, g.smem, d.stream, sp.pdl, in_map, out_map, halo_map, sp));
    ++*launches;
    if (rows * row < n) MAVG_TRY(tail_join(p, d, in, out, halo, frames, rows * row / g.C, launches));
    return MAVG_OK;
}

int validate(const mavg_desc* d)
{
    if (!d) return fail(MAVG_ERR_INVALID_ARG, "desc is null");
    if (d->struct_size != sizeof(mavg_desc))
        return fail(MAVG_ERR_INVALID_ARG, "desc.struct_size %u != %zu", d->struct_size, sizeof(mavg_desc));
    if (d->dtype > MAVG_I16) return fail(MAVG_ERR_INVALID_ARG, "unknown dtype %u", d->dtype);
    if (d->layout > MAVG_PLANAR) return fail(MAVG_ERR_INVALID_ARG, "unknown layout %u", d->layout);
    if (d->path > MAVG_PATH_GENERIC) return fail(MAVG_ERR_INVALID_ARG, "unknown path %u", d->path);
    if (d->op > MAVG_OP_RMS) return fail(MAVG_ERR_INVALID_ARG, "unknown op %u", d->op);
    if (d->channels == 0) return fail(MAVG_ERR_INVALID_ARG, "channels must be >= 1");
    if (d->window == 0) return fail(MAVG_ERR_INVALID_ARG, "window must be >= 1");
    if (d->num_devices > MAVG_MAX_DEVICES) return fail(MAVG_ERR_INVALID_ARG, "too many devices");
    // basics/profilable_sm_vload4.cu:231-234
    if (d->block_size != 0 && (d->block_size < 32 || d->block_size > 1024 || d->block_size % 32 != 0))
        return fail(MAVG_ERR_BLOCK_SIZE, "Block size must be multiple of 32 in 32..1024 (got %u)", d->block_size);
    if (d->first_frame != 0 && d->num_devices > 1)
        return fail(MAVG_ERR_INVALID_ARG, "first_frame > 0 is for single-device shard plans");
    return MAVG_OK;
}

void record(const mavg_plan* p, DevCtx& d, int which)
{
    if (p->timing_on) cudaEventRecord(d.ev[which], d.stream);
}

int gather_timing(mavg_plan* p)
{
    mavg_timing t = {0, 0, 0, 0};
    for (DevCtx& d : p->dev) {
        if (!d.timed) continue;
        float a = 0, b = 0, c = 0;
        if (cudaEventElapsedTime(&a, d.ev[0], d.ev[1]) != cudaSuccess) a = 0;
        if (cudaEventElapsedTime(&b, d.ev[1], d.ev[2]) != cudaSuccess) b = 0;
        if (cudaEventElapsedTime(&c, d.ev[2], d.ev[3]) != cudaSuccess) c = 0;
        cudaGetLastError();
        t.h2d_ms = std::max(t.h2d_ms, a);
        t.compute_ms = std::max(t.compute_ms, b);
        t.d2h_ms = std::max(t.d2h_ms, c);
        t.total_ms = std::max(t.total_ms, a + b + c);
    }
    p->timing = t;
    return MAVG_OK;
}

}  // namespace

namespace {
// Scratch of mavg_prefix_sum comes from a library-owned memory pool (one per device, created on first use) whose
// release threshold keeps freed blocks cached: the device's DEFAULT pool is left alone -- changing its threshold
// would change the behaviour of every other cudaMallocAsync user in the host process.
std::mutex g_pool_mutex;
cudaMemPool_t g_scan_pool[64] = {nullptr};

int scan_pool(int dev, cudaMemPool_t* out)
{
    if (dev < 0 || dev >= 64) return fail(MAVG_ERR_INVALID_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_scan_pool[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        MAVG_CUDA(cudaMemPoolCreate(&pool, &props));
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaGetLastError();
        g_scan_pool[dev] = pool;
    }
    *out = g_scan_pool[dev];
    return MAVG_OK;
}

template <typename TIn, typename TLoc, typename TAcc, int C, int CB>
int launch_scan_cb(const void* d_in, void* d_out, uint64_t n, cudaStream_t st)
{
    const uint64_t chunk = (uint64_t)mavg::scan_chunk_elems<TLoc, C, CB>();
    const uint64_t tiles = (n + chunk - 1) / chunk;
    if (tiles > 0x7fffffffull) return fail(MAVG_ERR_UNSUPPORTED, "signal too long for mavg_prefix_sum");
    auto kern = mavg::scan_lookback_kernel<TIn, TLoc, TAcc, C, CB>;
    const uint32_t smem = mavg::scan_smem_bytes<TLoc, C, CB>();
    MAVG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // scratch: ticket (padded to 256 bytes), 16-byte descriptors [chunks][C]; all zero = "not published"
    const size_t total = 256 + (size_t)tiles * C * sizeof(ulonglong2);
    int dev = 0;
    MAVG_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool = nullptr;
    MAVG_TRY(scan_pool(dev, &pool));
    char* scratch = nullptr;
    MAVG_CUDA(cudaMallocFromPoolAsync((void**)&scratch, total, pool, st));
    cudaError_t e = cudaMemsetAsync(scratch, 0, total, st);
    if (e == cudaSuccess) {
        uint32_t first_tile = 0;
        // int16 with 1, 2, 4 or 8 channels, float32 with 1, 2 or 4: the whole chunks go to the vectorised kernel, the
        // ragged last chunk of the same chain to the general one (stream-ordered behind it)
        constexpr int RL = CB / (int)sizeof(TLoc) / mavg::kScanThreads;   // run length
        constexpr bool kFast = (RL == 32 || RL == 16) &&
                               ((std::is_same<TIn, int16_t>::value && (C == 1 || C == 2 || C == 4 || C == 8)) ||
                                (std::is_same<TIn, float>::value && (C == 1 || C == 2 || C == 4)));
        if constexpr (kFast) {
            const uint64_t whole = n / chunk;
            const bool aligned = (((uintptr_t)d_in | (uintptr_t)d_out) & 15u) == 0;
            if (whole > 0 && aligned && !getenv("MAVG_SCAN_GENERAL")) {
                // int16: two chunks per CTA (MAVG_SCAN_NCH=1: one).  The second chunk's loads fly under the first chunk's
                // phases and it needs no look-back; an odd whole chunk goes to the general kernel with the ragged end.
                // 2^28 samples: mono 0.510 -> 0.463 ms, stereo 0.520 -> 0.451 ms.  float32 -> float64 stays at one chunk
                // (eight more 16-byte registers per thread cost it a resident CTA: 0.715 -> 0.724 ms).
                static const int nch_env = [] {
                    const char* ev = getenv("MAVG_SCAN_NCH");
                    return ev ? atoi(ev) : 0;
                }();
                // (4 / 8 channels keep one chunk per CTA: the second chunk's load registers would cost them resident CTAs)
                const int nch = nch_env == 1 ? 1 : nch_env == 2 ? 2 : (std::is_same<TIn, int16_t>::value && C <= 2 ? 2 : 1);
                const uint32_t fsmem = mavg::scan_fast_smem_bytes<TIn, C, RL>();
                if (nch == 2 && whole >= 2) {
                    auto fast = mavg::scan_lookback_fast_kernel<TIn, C, RL, 2>;
                    e = cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
                    if (e == cudaSuccess) {
                        fast<<<(unsigned)(whole / 2), mavg::kScanThreads, fsmem, st>>>((const TIn*)d_in, (TAcc*)d_out,
                                                                                      (ulonglong2*)(scratch + 256));
                        e = cudaGetLastError();
                        first_tile = (uint32_t)(whole / 2 * 2);
                    }
                } else {
                    auto fast = mavg::scan_lookback_fast_kernel<TIn, C, RL, 1>;
                    e = cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
                    if (e == cudaSuccess) {
                        fast<<<(unsigned)whole, mavg::kScanThreads, fsmem, st>>>((const TIn*)d_in, (TAcc*)d_out,
                                                                                (ulonglong2*)(scratch + 256));
                        e = cudaGetLastError();
                        first_tile = (uint32_t)whole;
                    }
                }
            }
        }
        if (e == cudaSuccess && first_tile < tiles) {
            kern<<<(unsigned)(tiles - first_tile), mavg::kScanThreads, smem, st>>>((const TIn*)d_in, (TAcc*)d_out, n, first_tile,
                                                                                  (ulonglong2*)(scratch + 256));
            e = cudaGetLastError();
        }
    }
    cudaFreeAsync(scratch, st);     // on every path: stream-ordered, after the kernel
    if (e != cudaSuccess) return fail(MAVG_ERR_CUDA, "prefix-sum launch failed: %s", cudaGetErrorString(e));
    return MAVG_OK;
}

// shared-memory bytes of chunk-local prefixes per CTA: MAVG_SCAN_CHUNK_KB in the environment (16, 32 or 64) overrides
// the default for measurements
int scan_chunk_kb()
{
    static const int kb = [] {
        const char* e = getenv("MAVG_SCAN_CHUNK_KB");
        const int v = e ? atoi(e) : 0;
        return (v == 16 || v == 32 || v == 64) ? v : 0;
    }();
    return kb;
}

template <typename TIn, typename TLoc, typename TAcc, int C>
int launch_scan_c(const void* d_in, void* d_out, uint64_t n, cudaStream_t st)
{
    // Measured on 2^28 samples (profiles/r02): int16 runs best with 32 KB of chunk-local prefixes per CTA (8192 elements,
    // six CTAs per SM in different phases), float32 -> float64 with 64 KB (its prefixes are 8 bytes: the same 8192 elements)
    if constexpr (C == 1 || C == 2) {   // the chunk-size variants exist for the shapes that are measured
        int kb = scan_chunk_kb();
        if (kb == 0) kb = sizeof(TLoc) == 4 ? 32 : 64;
        if (kb == 16) return launch_scan_cb<TIn, TLoc, TAcc, C, 16384>(d_in, d_out, n, st);
        if (kb == 32) return launch_scan_cb<TIn, TLoc, TAcc, C, 32768>(d_in, d_out, n, st);
        return launch_scan_cb<TIn, TLoc, TAcc, C, 65536>(d_in, d_out, n, st);
    }
    return launch_scan_cb<TIn, TLoc, TAcc, C, (sizeof(TLoc) == 4 ? 32768 : 65536)>(d_in, d_out, n, st);
}

template <typename TIn, typename TLoc, typename TAcc>
int launch_scan(const void* d_in, void* d_out, uint64_t n, uint32_t C, cudaStream_t st)
{
    switch (C) {
    case 1: return launch_scan_c<TIn, TLoc, TAcc, 1>(d_in, d_out, n, st);
    case 2: return launch_scan_c<TIn, TLoc, TAcc, 2>(d_in, d_out, n, st);
    case 3: return launch_scan_c<TIn, TLoc, TAcc, 3>(d_in, d_out, n, st);
    case 4: return launch_scan_c<TIn, TLoc, TAcc, 4>(d_in, d_out, n, st);
    case 5: return launch_scan_c<TIn, TLoc, TAcc, 5>(d_in, d_out, n, st);
    case 6: return launch_scan_c<TIn, TLoc, TAcc, 6>(d_in, d_out, n, st);
    case 7: return launch_scan_c<TIn, TLoc, TAcc, 7>(d_in, d_out, n, st);
    case 8: return launch_scan_c<TIn, TLoc, TAcc, 8>(d_in, d_out, n, st);
    default: return fail(MAVG_ERR_UNSUPPORTED, "mavg_prefix_sum takes 1 to 8 interleaved channels (got %u)", C);
    }
}
}  // namespace


// =================================================================================
// C ABI
// =================================================================================
extern "C" {

int mavg_version(void) { return MAVG_VERSION_MAJOR * 10000 + MAVG_VERSION_MINOR * 100 + MAVG_VERSION_PATCH; }

const char* mavg_strerror(int status)
{
    switch (status) {
    case MAVG_OK: return "ok";
    case MAVG_ERR_INVALID_ARG: return "invalid argument";
    case MAVG_ERR_UNSUPPORTED: return "unsupported configuration";
    case MAVG_ERR_CUDA: return "CUDA runtime error";
    case MAVG_ERR_NO_DEVICE: return "no CUDA device";
    case MAVG_ERR_ALLOC: return "allocation failed";
    case MAVG_ERR_BLOCK_SIZE: return "block size must be a multiple of 32 in 32..1024";
    case MAVG_ERR_DRIVER: return "CUDA driver entry point unavailable";
    default: return "unknown status";
    }
}

const char* mavg_last_error(void) { return g_last_error.c_str(); }

int mavg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mavg_plan_create(const mavg_desc* desc, mavg_plan** out)
{
    if (!out) return fail(MAVG_ERR_INVALID_ARG, "plan out-pointer is null");
    *out = nullptr;
    MAVG_TRY(validate(desc));
    int ndev_avail = mavg_device_count();
    if (ndev_avail <= 0) return fail(MAVG_ERR_NO_DEVICE, "no CUDA device available: libmavg has no CPU fallback");

    mavg_plan* p = new (std::nothrow) mavg_plan();
    if (!p) return fail(MAVG_ERR_ALLOC, "out of host memory");
    p->desc = *desc;
    DeviceGuard guard;

    // ---- kernel family
    const bool planar = desc->layout == MAVG_PLANAR && desc->channels > 1;
    bool stream_shape;
    if (desc->op == MAVG_OP_RMS) {
        // moving RMS: the float32 TMA streaming kernel (mono / stereo / planar, default shape) squares on load and
        // takes the root on store; every other shape is served by the generic kernel (fp64 / exact int64 sums of squares)
        stream_shape = desc->dtype == MAVG_F32 && (desc->channels <= 2 || planar);
        if (stream_shape) {
            mavg_tuning tu = desc->tuning;
            tu.threads = 0;
            tu.run = 0;
            p->geom = plan_stream(desc->window, tu, planar ? 1u : desc->channels);
        }
    } else if (desc->dtype == MAVG_F32) {
        stream_shape = desc->channels <= 2 || planar;
        p->geom = plan_stream(desc->window, desc->tuning, planar ? 1u : desc->channels);
        if (!p->geom.ok && (planar || desc->channels <= 2)) p->geom = plan_far(desc->window, planar ? 1u : desc->channels, desc->tuning);
        if (!planar && desc->channels >= 32) {
            p->geom = plan_cols(desc->window, desc->channels, desc->tuning);
            stream_shape = p->geom.ok;
        } else if (!planar && desc->channels >= 3) {
            p->geom = plan_fewc(desc->window, desc->channels, desc->tuning);
            stream_shape = p->geom.ok;
        }
    } else {
        stream_shape = desc->channels <= 2 || planar;
        p->geom = plan_stream_i16(desc->window, planar ? 1u : desc->channels, desc->tuning);
        if (!planar && desc->channels >= 32) {
            p->geom = plan_cols(desc->window, desc->channels, desc->tuning, true);
            stream_shape = p->geom.ok;
        } else {
            // 3+ channels: the flat-stream kernel while the window fits its ring (stream_shape follows the plan)
            if (!planar && desc->channels >= 3) stream_shape = p->geom.ok;
            // beyond the ring: the far-lag int16 kernel (mono / stereo / planar, 4 / 6 / 8 / 12 / 16 channels, k <= 46 340) ...
            if (!p->geom.ok) {
                const StreamGeom gf = plan_far_i16(desc->window, planar ? 1u : desc->channels, desc->tuning);
                if (gf.ok) {
                    p->geom = gf;
                    stream_shape = true;
                }
            }
            // ... then the few-channel kernels
            if (!p->geom.ok && !planar && desc->channels >= 3) {
                p->geom = plan_fewc(desc->window, desc->channels, desc->tuning, 2);
                stream_shape = p->geom.ok;
            }
        }
    }
    if (desc->path == MAVG_PATH_STREAM && !(stream_shape && p->geom.ok)) {
        delete p;
        return fail(MAVG_ERR_UNSUPPORTED,
                    "stream path needs mono, stereo or planar input and a window that fits the shared-memory history");
    }
    p->path = (desc->path != MAVG_PATH_GENERIC && stream_shape && p->geom.ok) ? MAVG_PATH_STREAM : MAVG_PATH_GENERIC;

    p->prefix_diff = p->path == MAVG_PATH_GENERIC && desc->path != MAVG_PATH_GENERIC && desc->op == MAVG_OP_MEAN &&
                     !planar && desc->channels <= 8 && (uint64_t)desc->window * desc->channels >= 8192;
    // ---- left context a frame shard needs
    if (p->path == MAVG_PATH_STREAM) {
        p->halo_frames = (uint64_t)p->geom.H * tile_frames(p);  // whole history tiles: sharding keeps bit-identical sums
    } else {
        p->halo_frames = desc->window;
    }

    // ---- devices and shards
    const uint32_t nd = std::max<uint32_t>(1u, desc->num_devices);
    int cur = 0;
    cudaGetDevice(&cur);
    p->dev.resize(nd);
    // frame shards start on tile boundaries (stream) so every device runs the same tile grid; the generic kernel's
    // 64-frame runs are anchored at the shard start, so its shards start on multiples of 64 frames
    const uint64_t align = !frame_sharded(p) ? 1 : (p->path == MAVG_PATH_STREAM ? tile_frames(p) : 64);
    for (uint32_t r = 0; r < nd; ++r) {
        DevCtx& d = p->dev[r];
        d.device = desc->num_devices >= 1 ? desc->devices[r] : cur;
        if (desc->num_devices == 0) d.device = cur;
        if (d.device < 0 || d.device >= ndev_avail) {
            mavg_plan_destroy(p);
            return fail(MAVG_ERR_INVALID_ARG, "device %d out of range (have %d)", d.device, ndev_avail);
        }
        if (planar_batch(p)) {
            const uint32_t c0 = (uint32_t)((uint64_t)desc->channels * r / nd);
            const uint32_t c1 = (uint32_t)((uint64_t)desc->channels * (r + 1) / nd);
            d.first_channel = c0;
            d.channels = c1 - c0;
            d.frames = desc->frames;
        } else {
            auto cut = [&](uint32_t i) -> uint64_t {
                if (i == 0) return 0;
                if (i == nd) return desc->frames;
                uint64_t f = desc->frames / nd * i;
                f = (f + align - 1) / align * align;
                return std::min<uint64_t>(f, desc->frames);
            };
            d.first_frame = cut(r);
            d.frames = cut(r + 1) - cut(r);
            d.channels = desc->channels;
            if (r > 0 && d.frames > 0 && p->dev[r - 1].frames < p->halo_frames) {
                mavg_plan_destroy(p);
                return fail(MAVG_ERR_UNSUPPORTED, "signal too short to shard over %u devices with window %u", nd,
                            desc->window);
            }
        }
        cudaError_t e = cudaSetDevice(d.device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&d.ev[i]);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device);
        if (e != cudaSuccess) {
            mavg_plan_destroy(p);
            return fail(MAVG_ERR_CUDA, "device %d setup failed: %s", d.device, cudaGetErrorString(e));
        }
    }
    // peer access between neighbours (halo is read in place from the left neighbour's shard)
    p->peer_ok = nd > 1;
    for (uint32_t r = 1; r < nd && frame_sharded(p); ++r) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, p->dev[r].device, p->dev[r - 1].device);
        if (!can) { p->peer_ok = false; continue; }
        cudaSetDevice(p->dev[r].device);
        cudaError_t e = cudaDeviceEnablePeerAccess(p->dev[r - 1].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) p->peer_ok = false;
        cudaGetLastError();
    }
    *out = p;
    return MAVG_OK;
}

int mavg_plan_destroy(mavg_plan* p)
{
    if (!p) return MAVG_OK;
    DeviceGuard guard;
    for (DevCtx& d : p->dev) {
        if (cudaSetDevice(d.device) != cudaSuccess) { cudaGetLastError(); continue; }
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.d_in) cudaFree(d.d_in);
        if (d.d_out) cudaFree(d.d_out);
        if (d.d_halo) cudaFree(d.d_halo);
        if (d.d_bsum) cudaFree(d.d_bsum);
        if (d.d_scratch) cudaFree(d.d_scratch);
        if (d.d_far_stage) cudaFree(d.d_far_stage);
        if (d.d_sweep_in) cudaFree(d.d_sweep_in);
        if (d.d_prefix) cudaFree(d.d_prefix);
        if (d.ev_pass) cudaEventDestroy(d.ev_pass);
        if (d.s_tail) { cudaStreamSynchronize(d.s_tail); cudaStreamDestroy(d.s_tail); }
        if (d.ev_fork) cudaEventDestroy(d.ev_fork);
        if (d.ev_join) cudaEventDestroy(d.ev_join);
        if (d.s_h2d) { cudaStreamSynchronize(d.s_h2d); cudaStreamDestroy(d.s_h2d); }
        if (d.s_d2h) { cudaStreamSynchronize(d.s_d2h); cudaStreamDestroy(d.s_d2h); }
        for (cudaEvent_t e : d.pool) cudaEventDestroy(e);
        for (int i = 0; i < 4; ++i)
            if (d.ev[i]) cudaEventDestroy(d.ev[i]);
        if (d.stream && d.own_stream) cudaStreamDestroy(d.stream);
    }
    cudaGetLastError();
    delete p;
    return MAVG_OK;
}

int mavg_plan_info(const mavg_plan* p, mavg_info* info)
{
    if (!p || !info) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    memset(info, 0, sizeof *info);
    info->path = p->path;
    info->mode = p->prefix_diff ? 7u : (uint32_t)p->geom.mode;
    info->threads = p->geom.NT;
    info->run = p->geom.R;
    info->tile_samples = p->geom.NT * p->geom.R;
    info->history_tiles = p->geom.H;
    info->stages = p->geom.S;
    info->grid = p->dev.empty() ? 0 : p->dev[0].sm_count * p->geom.ctas_per_sm;
    info->smem_bytes = p->geom.smem;
    info->launches_per_run = p->launches_last_run;
    info->num_devices = (uint32_t)p->dev.size();
    info->halo_frames = p->halo_frames;
    for (size_t r = 0; r < p->dev.size(); ++r)
        info->shard_frames[r] = planar_batch(p) ? p->dev[r].channels : p->dev[r].frames;
    return MAVG_OK;
}

int mavg_set_stream(mavg_plan* p, void* cuda_stream)
{
    if (!p) return fail(MAVG_ERR_INVALID_ARG, "plan is null");
    if (p->dev.size() != 1) return fail(MAVG_ERR_UNSUPPORTED, "mavg_set_stream is for single-device plans");
    DevCtx& d = p->dev[0];
    DeviceGuard guard;
    MAVG_CUDA(cudaSetDevice(d.device));
    if (d.stream && d.own_stream) {
        cudaStreamSynchronize(d.stream);
        cudaStreamDestroy(d.stream);
    }
    d.stream = (cudaStream_t)cuda_stream;
    d.own_stream = false;
    return MAVG_OK;
}

int mavg_enable_timing(mavg_plan* p, int enable)
{
    if (!p) return fail(MAVG_ERR_INVALID_ARG, "plan is null");
    p->timing_on = enable != 0;
    if (!p->timing_on)
        for (DevCtx& d : p->dev) d.timed = false;
    return MAVG_OK;
}

int mavg_plan_buffers(mavg_plan* p, uint32_t rank, void** d_in, void** d_out)
{
    if (!p || rank >= p->dev.size()) return fail(MAVG_ERR_INVALID_ARG, "bad plan or rank");
    DeviceGuard guard;
    MAVG_TRY(alloc_owned(p, p->dev[rank]));
    if (d_in) *d_in = p->dev[rank].d_in;
    if (d_out) *d_out = p->dev[rank].d_out;
    return MAVG_OK;
}

int mavg_run_device_halo(mavg_plan* p, const void* d_in, void* d_out, const void* d_halo)
{
    if (!p || !d_in || !d_out) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (d_in == d_out) return fail(MAVG_ERR_INVALID_ARG, "output must not alias input (windows re-read samples behind the write front)");
    if (p->dev.size() != 1) return fail(MAVG_ERR_UNSUPPORTED, "mavg_run_device_halo is for single-device plans");
    if (d_halo && !frame_sharded(p)) return fail(MAVG_ERR_UNSUPPORTED, "planar batches take no halo");
    DeviceGuard guard;
    DevCtx& d = p->dev[0];
    MAVG_CUDA(cudaSetDevice(d.device));
    uint32_t launches = 0;
    record(p, d, 0);
    record(p, d, 1);
    int s = launch_shard(p, d, d_in, d_out, d_halo, d.frames, &launches);
    record(p, d, 2);
    record(p, d, 3);
    d.timed = p->timing_on;
    p->launches_last_run = launches;
    return s;
}

int mavg_run_device(mavg_plan* p, const void* const* d_in, void* const* d_out)
{
    if (!p || !d_in || !d_out) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    DeviceGuard guard;
    uint32_t launches = 0;
    const size_t es = elem_size(p->desc.dtype);
    for (size_t r = 0; r < p->dev.size(); ++r) {
        DevCtx& d = p->dev[r];
        if (!d_in[r] || !d_out[r]) return fail(MAVG_ERR_INVALID_ARG, "null shard pointer for device index %zu", r);
        if (d_in[r] == d_out[r]) return fail(MAVG_ERR_INVALID_ARG, "output must not alias input");
        MAVG_CUDA(cudaSetDevice(d.device));
        const void* halo = nullptr;
        record(p, d, 0);
        if (r > 0 && frame_sharded(p) && d.frames > 0) {
            // left context = tail of the left neighbour's shard
            const DevCtx& l = p->dev[r - 1];
            const char* tail = (const char*)d_in[r - 1] + (l.frames - p->halo_frames) * p->desc.channels * es;
            if (p->peer_ok) {
                halo = tail;  // read in place over NVLink by the kernel's TMA / global loads
            } else {
                MAVG_TRY(alloc_halo(p, d));
                MAVG_CUDA(cudaMemcpyPeerAsync(d.d_halo, d.device, tail, l.device,
                                              p->halo_frames * p->desc.channels * es, d.stream));
                halo = d.d_halo;
            }
        }
        record(p, d, 1);
        MAVG_TRY(launch_shard(p, d, d_in[r], d_out[r], halo, d.frames, &launches));
        record(p, d, 2);
        record(p, d, 3);
        d.timed = p->timing_on;
    }
    p->launches_last_run = launches;
    return MAVG_OK;
}

int mavg_run_cascade(mavg_plan* p, const void* const* d_in, void* const* d_out, void* const* d_scratch, uint32_t passes)
{
    if (!p || !d_in || !d_out) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (passes == 0) return fail(MAVG_ERR_INVALID_ARG, "passes must be >= 1");
    if (passes == 1) return mavg_run_device(p, d_in, d_out);
    if (p->desc.first_frame != 0) return fail(MAVG_ERR_UNSUPPORTED, "cascades need whole-signal plans (first_frame == 0)");
    DeviceGuard guard;
    const size_t nd = p->dev.size();
    const size_t es = elem_size(p->desc.dtype);
    void* scratch[MAVG_MAX_DEVICES];
    for (size_t r = 0; r < nd; ++r) {
        DevCtx& d = p->dev[r];
        if (!d_in[r] || !d_out[r]) return fail(MAVG_ERR_INVALID_ARG, "null shard pointer for device index %zu", r);
        if (d_in[r] == d_out[r]) return fail(MAVG_ERR_INVALID_ARG, "output must not alias input");
        scratch[r] = d_scratch ? d_scratch[r] : nullptr;
        if (!scratch[r]) {
            if (!d.d_scratch) {
                const size_t bytes = std::max<size_t>(shard_elems(p, d) * es, 256);
                MAVG_CUDA(cudaSetDevice(d.device));
                if (cudaMalloc(&d.d_scratch, bytes) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu scratch bytes failed on device %d", bytes, d.device);
                }
            }
            scratch[r] = d.d_scratch;
        }
        if (scratch[r] == d_in[r] || scratch[r] == d_out[r]) return fail(MAVG_ERR_INVALID_ARG, "scratch must not alias input or output");
        if (!d.ev_pass) {
            MAVG_CUDA(cudaSetDevice(d.device));
            MAVG_CUDA(cudaEventCreateWithFlags(&d.ev_pass, cudaEventDisableTiming));
        }
    }
    // one timed region around all passes: start events here, per-pass recording off, end events after the last pass
    const bool was_timed = p->timing_on;
    // pass i + 1 reads what pass i wrote: tile loads must not start before the previous kernel has completed
    const uint32_t overlap_saved = p->desc.tuning.overlap;
    if (pdl_mode(p->desc.tuning) == 2) p->desc.tuning.overlap = 1;
    if (was_timed)
        for (size_t r = 0; r < nd; ++r) {
            MAVG_CUDA(cudaSetDevice(p->dev[r].device));
            record(p, p->dev[r], 0);
            record(p, p->dev[r], 1);
        }
    uint32_t launches = 0;
    const void* src[MAVG_MAX_DEVICES];
    void* dst[MAVG_MAX_DEVICES];
    int rc = MAVG_OK;
    p->timing_on = false;
    for (uint32_t i = 0; i < passes && rc == MAVG_OK; ++i) {
        // ping-pong so that the last pass writes d_out: pass i writes d_out when (passes - 1 - i) is even
        for (size_t r = 0; r < nd; ++r) {
            src[r] = (i == 0) ? d_in[r] : dst[r];
            dst[r] = ((passes - 1 - i) % 2 == 0) ? d_out[r] : scratch[r];
        }
        // a pass reads its left neighbour's previous output (halo) and overwrites the buffer its right neighbour
        // read one pass earlier: every device waits for every device's previous pass
        if (i > 0 && nd > 1) {
            for (size_t r = 0; r < nd && rc == MAVG_OK; ++r) {
                if (cudaSetDevice(p->dev[r].device) != cudaSuccess) rc = fail(MAVG_ERR_CUDA, "cudaSetDevice failed");
                for (size_t q = 0; q < nd && rc == MAVG_OK; ++q)
                    if (q != r && cudaStreamWaitEvent(p->dev[r].stream, p->dev[q].ev_pass, 0) != cudaSuccess)
                        rc = fail(MAVG_ERR_CUDA, "cudaStreamWaitEvent failed");
            }
            if (rc != MAVG_OK) break;
        }
        rc = mavg_run_device(p, src, dst);
        launches += p->launches_last_run;
        if (rc == MAVG_OK && nd > 1)
            for (size_t r = 0; r < nd; ++r) {
                cudaSetDevice(p->dev[r].device);
                if (cudaEventRecord(p->dev[r].ev_pass, p->dev[r].stream) != cudaSuccess)
                    rc = fail(MAVG_ERR_CUDA, "cudaEventRecord failed");
            }
    }
    p->timing_on = was_timed;
    p->desc.tuning.overlap = overlap_saved;
    if (rc != MAVG_OK) return rc;
    if (was_timed)
        for (size_t r = 0; r < nd; ++r) {
            MAVG_CUDA(cudaSetDevice(p->dev[r].device));
            record(p, p->dev[r], 2);
            record(p, p->dev[r], 3);
            p->dev[r].timed = true;
        }
    p->launches_last_run = launches;
    return MAVG_OK;
}

int mavg_run_owned(mavg_plan* p)
{
    if (!p) return fail(MAVG_ERR_INVALID_ARG, "plan is null");
    const void* in[MAVG_MAX_DEVICES];
    void* out[MAVG_MAX_DEVICES];
    DeviceGuard guard;
    for (size_t r = 0; r < p->dev.size(); ++r) {
        MAVG_TRY(alloc_owned(p, p->dev[r]));
        in[r] = p->dev[r].d_in;
        out[r] = p->dev[r].d_out;
    }
    return mavg_run_device(p, in, out);
}

int mavg_synchronize(mavg_plan* p)
{
    if (!p) return fail(MAVG_ERR_INVALID_ARG, "plan is null");
    DeviceGuard guard;
    for (DevCtx& d : p->dev) {
        MAVG_CUDA(cudaSetDevice(d.device));
        if (d.s_h2d) MAVG_CUDA(cudaStreamSynchronize(d.s_h2d));
        MAVG_CUDA(cudaStreamSynchronize(d.stream));
        if (d.s_d2h) MAVG_CUDA(cudaStreamSynchronize(d.s_d2h));
    }
    return gather_timing(p);
}

int mavg_get_timing(mavg_plan* p, mavg_timing* t)
{
    if (!p || !t) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    *t = p->timing;
    return MAVG_OK;
}

namespace {

// cudaMemcpyAsync into PAGEABLE host memory blocks the calling thread for the whole copy, so one thread issuing the
// H2D and the D2H copies of mavg_run_host runs the two directions one after the other (measured: 1.75 Gsamples/s
// on 2^28 float32 samples).  With a pageable destination a helper thread issues the D2H copies, so that they overlap
// the caller thread's H2D copies and kernel launches.  Page-locked buffers never come here.
class D2HWorker {
public:
    struct Job {
        int device;
        cudaStream_t stream;
        cudaEvent_t after;   // the slice's kernels
        void* dst;
        const void* src;
        size_t bytes;
    };
    ~D2HWorker() { finish(); }
    bool active() const { return started_; }
    void start()
    {
        started_ = true;
        th_ = std::thread([this] { loop(); });
    }
    void push(const Job& j)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(j);
        }
        cv_.notify_one();
    }
    // no more jobs: wait for the queued copies to be issued (and therefore finished: pageable copies are synchronous)
    cudaError_t finish()
    {
        if (started_ && th_.joinable()) {
            {
                std::lock_guard<std::mutex> g(m_);
                done_ = true;
            }
            cv_.notify_one();
            th_.join();
        }
        return err_;
    }

private:
    void loop()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this] { return done_ || !q_.empty(); });
                if (q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            cudaError_t e = cudaSetDevice(j.device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(j.stream, j.after, 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyDeviceToHost, j.stream);
            if (e != cudaSuccess && err_ == cudaSuccess) err_ = e;
        }
    }
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Job> q_;
    bool done_ = false, started_ = false;
    cudaError_t err_ = cudaSuccess;
};

bool is_pageable(const void* ptr)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

namespace {
int run_host_impl(mavg_plan* p, const void* h_in, void* h_out);
}

int mavg_run_host(mavg_plan* p, const void* h_in, void* h_out)
{
    if (!p || !h_in || !h_out) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (h_in == h_out) return fail(MAVG_ERR_INVALID_ARG, "output must not alias input");
    DeviceGuard guard;
    const int rc = run_host_impl(p, h_in, h_out);
    if (rc != MAVG_OK) {
        // a failure in the middle of the pipeline leaves copies queued on the caller's buffers: drain them before the
        // caller gets its buffers back (it may free them at once); the first error stays the one reported
        const std::string first = g_last_error;
        for (DevCtx& d : p->dev) {
            if (cudaSetDevice(d.device) != cudaSuccess) continue;
            if (d.s_h2d) cudaStreamSynchronize(d.s_h2d);
            if (d.stream) cudaStreamSynchronize(d.stream);
            if (d.s_d2h) cudaStreamSynchronize(d.s_d2h);
        }
        cudaGetLastError();
        g_last_error = first;
    }
    return rc;
}

namespace {
int run_host_impl(mavg_plan* p, const void* h_in, void* h_out)
{
    const size_t es = elem_size(p->desc.dtype);
    const uint64_t C = p->desc.channels;
    uint32_t launches = 0;
    D2HWorker worker;   // joins in its destructor on every return path
    if ((uint64_t)p->desc.frames * C * es >= (8ull << 20) && is_pageable(h_out)) worker.start();
    for (DevCtx& d : p->dev) {
        MAVG_TRY(alloc_owned(p, d));
        MAVG_CUDA(cudaSetDevice(d.device));
        if (!d.s_h2d) MAVG_CUDA(cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking));
        if (!d.s_d2h) MAVG_CUDA(cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking));
        const uint64_t elems = shard_elems(p, d);
        const char* src;
        const void* halo0 = nullptr;
        // order the copy streams behind whatever the compute stream was doing before this call
        if (d.pool.empty()) {
            cudaEvent_t e;
            MAVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            d.pool.push_back(e);
        }
        MAVG_CUDA(cudaEventRecord(d.pool[0], d.stream));
        MAVG_CUDA(cudaStreamWaitEvent(d.s_h2d, d.pool[0], 0));
        MAVG_CUDA(cudaStreamWaitEvent(d.s_d2h, d.pool[0], 0));
        if (p->timing_on) cudaEventRecord(d.ev[0], d.s_h2d);
        if (planar_batch(p)) {
            src = (const char*)h_in + (uint64_t)d.first_channel * p->desc.frames * es;
        } else {
            src = (const char*)h_in + d.first_frame * C * es;
            // a later shard's left context sits right before it in host memory: inside the caller's
            // whole-signal buffer (multi-device plans) or, for a shard plan (desc.first_frame > 0), in
            // the halo_frames frames the caller keeps in front of h_in
            if ((d.first_frame > 0 || p->desc.first_frame > 0) && d.frames > 0) {
                MAVG_TRY(alloc_halo(p, d));
                const uint64_t hb = p->halo_frames * C * es;
                MAVG_CUDA(cudaMemcpyAsync(d.d_halo, src - hb, hb, cudaMemcpyHostToDevice, d.s_h2d));
                halo0 = d.d_halo;
            }
        }
        char* dst = (char*)h_out + (src - (const char*)h_in);

        // slices: whole tiles, at least the left context long; planar batches and short shards go in one piece.
        // The H2D stream is the critical path: N slices of s bytes take N (s / B + o) + s / B with a fixed cost
        // o per slice (measured 21 us at B = 49.5 GB/s each way, PCIe 5 x16), which is least at s = sqrt(total o B)
        // ~ sqrt(total * 1 MiB): 32 MiB slices for a 1 GiB shard, 8 MiB for 64 MiB (profiles/r01/pcie_probe.json).
        // Reading and writing the pinned host buffers straight from the kernel (no copies, one launch) was
        // measured at 38.5 GB/s each way against 47 for the sliced copies, so it is not used.
        uint64_t slice_frames = d.frames;
        if (!planar_batch(p) && d.frames > 0) {
            const uint64_t unit = (p->path == MAVG_PATH_STREAM) ? tile_frames(p) : 1024;
            uint64_t slice_bytes = p->desc.tuning.slice_bytes;
            if (!slice_bytes) {
                slice_bytes = 4ull << 20;
                while (slice_bytes < (64ull << 20) && slice_bytes * slice_bytes < ((d.frames * C * es) << 20)) slice_bytes <<= 1;
            }
            uint64_t want = std::max<uint64_t>(slice_bytes / (C * es), p->halo_frames);
            want = (want + unit - 1) / unit * unit;
            if (want * 2 <= d.frames) slice_frames = want;
        }
        const uint64_t nslices = (planar_batch(p) || d.frames == 0) ? 1 : (d.frames + slice_frames - 1) / slice_frames;
        while (d.pool.size() < 1 + 2 * nslices) {
            cudaEvent_t e;
            MAVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            d.pool.push_back(e);
        }
        for (uint64_t i = 0; i < nslices && elems > 0; ++i) {
            uint64_t f0, fcount, off_e, cnt_e;
            if (planar_batch(p)) {
                f0 = 0; fcount = p->desc.frames; off_e = 0; cnt_e = elems;
            } else {
                f0 = i * slice_frames;
                fcount = std::min<uint64_t>(slice_frames, d.frames - f0);
                off_e = f0 * C; cnt_e = fcount * C;
            }
            cudaEvent_t e_in = d.pool[1 + 2 * i], e_k = d.pool[2 + 2 * i];
            MAVG_CUDA(cudaMemcpyAsync((char*)d.d_in + off_e * es, src + off_e * es, cnt_e * es, cudaMemcpyHostToDevice,
                                      d.s_h2d));
            MAVG_CUDA(cudaEventRecord(e_in, d.s_h2d));
            if (i + 1 == nslices && p->timing_on) cudaEventRecord(d.ev[1], d.s_h2d);
            MAVG_CUDA(cudaStreamWaitEvent(d.stream, e_in, 0));
            const void* halo = (i == 0) ? halo0 : (const char*)d.d_in + (off_e - p->halo_frames * C) * es;
            MAVG_TRY(launch_shard(p, d, (const char*)d.d_in + off_e * es, (char*)d.d_out + off_e * es, halo, fcount,
                                  &launches));
            MAVG_CUDA(cudaEventRecord(e_k, d.stream));
            if (i + 1 == nslices && p->timing_on) cudaEventRecord(d.ev[2], d.stream);
            if (worker.active()) {
                worker.push({d.device, d.s_d2h, e_k, dst + off_e * es, (char*)d.d_out + off_e * es, (size_t)(cnt_e * es)});
            } else {
                MAVG_CUDA(cudaStreamWaitEvent(d.s_d2h, e_k, 0));
                MAVG_CUDA(cudaMemcpyAsync(dst + off_e * es, (char*)d.d_out + off_e * es, cnt_e * es, cudaMemcpyDeviceToHost,
                                          d.s_d2h));
            }
        }
        if (elems == 0 && p->timing_on) {
            cudaEventRecord(d.ev[1], d.s_h2d);
            cudaEventRecord(d.ev[2], d.stream);
        }
    }
    MAVG_CUDA(worker.finish());
    for (DevCtx& d : p->dev) {
        if (p->timing_on) {
            MAVG_CUDA(cudaSetDevice(d.device));
            cudaEventRecord(d.ev[3], d.s_d2h);
        }
        d.timed = p->timing_on;
    }
    p->launches_last_run = launches;
    return mavg_synchronize(p);
}
}  // namespace

namespace {
uint64_t gcd_u64(uint64_t a, uint64_t b)
{
    while (b) { const uint64_t t = a % b; a = b; b = t; }
    return a;
}

// true when the sweep pipeline can serve these plans; otherwise mavg_run_host is called once per plan
bool sweep_compatible(mavg_plan* const* plans, uint32_t count, void* const* h_out)
{
    const mavg_plan* p0 = plans[0];
    if (planar_batch(p0) || p0->dev.size() != 1 || p0->desc.frames == 0) return false;
    for (uint32_t i = 0; i < count; ++i) {
        const mavg_plan* p = plans[i];
        if (p->dev.size() != 1 || p->dev[0].device != p0->dev[0].device) return false;
        if (p->desc.dtype != p0->desc.dtype || p->desc.layout != p0->desc.layout || p->desc.channels != p0->desc.channels ||
            p->desc.frames != p0->desc.frames || p->desc.first_frame != p0->desc.first_frame)
            return false;
        if (is_pageable(h_out[i])) return false;
    }
    return true;
}

int run_host_sweep_impl(mavg_plan* const* plans, uint32_t count, const void* h_in, void* const* h_out)
{
    mavg_plan* p0 = plans[0];
    DevCtx& d0 = p0->dev[0];
    const size_t es = elem_size(p0->desc.dtype);
    const uint64_t C = p0->desc.channels;
    const uint64_t frames = d0.frames;
    const bool has_halo = p0->desc.first_frame > 0;
    MAVG_CUDA(cudaSetDevice(d0.device));
    // slices: whole tiles of every plan (least common multiple of the plans' units), at least the longest left context
    uint64_t unit = 1, max_halo = 0;
    for (uint32_t i = 0; i < count; ++i) {
        const uint64_t u = plans[i]->path == MAVG_PATH_STREAM ? tile_frames(plans[i]) : 1024;
        unit = unit / gcd_u64(unit, u) * u;
        max_halo = std::max<uint64_t>(max_halo, plans[i]->halo_frames);
        if (unit > (1ull << 26)) return fail(MAVG_ERR_UNSUPPORTED, "sweep: the plans' tile sizes have no common slice size");
    }
    const uint64_t halo0 = has_halo ? max_halo : 0;
    uint64_t slice_bytes = p0->desc.tuning.slice_bytes;
    if (!slice_bytes) {
        slice_bytes = 4ull << 20;
        while (slice_bytes < (64ull << 20) && slice_bytes * slice_bytes < ((frames * C * es) << 20)) slice_bytes <<= 1;
    }
    uint64_t slice_frames = std::max<uint64_t>(slice_bytes / (C * es), max_halo);
    slice_frames = (slice_frames + unit - 1) / unit * unit;
    if (slice_frames * 2 > frames) slice_frames = frames;
    const uint64_t nslices = (frames + slice_frames - 1) / slice_frames;

    // device buffers: the input once (with the longest halo in front of it), one output per plan
    const size_t halo_pad = ((size_t)halo0 * C * es + 255) / 256 * 256;   // the shard starts 256-byte aligned on the device
    const size_t in_bytes = halo_pad + (size_t)frames * C * es;
    if (d0.sweep_in_bytes < in_bytes) {
        if (d0.d_sweep_in) MAVG_CUDA(cudaFree(d0.d_sweep_in));
        d0.d_sweep_in = nullptr;
        d0.sweep_in_bytes = 0;
        if (cudaMalloc(&d0.d_sweep_in, std::max<size_t>(in_bytes, 256)) != cudaSuccess) {
            cudaGetLastError();
            return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu sweep input bytes failed on device %d", in_bytes, d0.device);
        }
        d0.sweep_in_bytes = in_bytes;
    }
    for (uint32_t i = 0; i < count; ++i) {
        DevCtx& d = plans[i]->dev[0];
        if (!d.d_out) {
            const size_t bytes = std::max<size_t>((size_t)frames * C * es, 256);
            if (cudaMalloc(&d.d_out, bytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(MAVG_ERR_ALLOC, "cudaMalloc of %zu output bytes failed on device %d", bytes, d.device);
            }
        }
    }
    if (!d0.s_h2d) MAVG_CUDA(cudaStreamCreateWithFlags(&d0.s_h2d, cudaStreamNonBlocking));
    if (!d0.s_d2h) MAVG_CUDA(cudaStreamCreateWithFlags(&d0.s_d2h, cudaStreamNonBlocking));
    const size_t nev = 1 + (size_t)nslices * (1 + count);
    while (d0.pool.size() < nev) {
        cudaEvent_t e;
        MAVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d0.pool.push_back(e);
    }
    // the copy streams start behind whatever the plans' compute streams were doing before this call
    for (uint32_t i = 0; i < count; ++i) {
        MAVG_CUDA(cudaEventRecord(d0.pool[0], plans[i]->dev[0].stream));
        MAVG_CUDA(cudaStreamWaitEvent(d0.s_h2d, d0.pool[0], 0));
        MAVG_CUDA(cudaStreamWaitEvent(d0.s_d2h, d0.pool[0], 0));
    }
    if (p0->timing_on) cudaEventRecord(d0.ev[0], d0.s_h2d);
    char* const dev_in = (char*)d0.d_sweep_in + halo_pad;             // first frame of the shard
    if (has_halo)
        MAVG_CUDA(cudaMemcpyAsync(dev_in - halo0 * C * es, (const char*)h_in - halo0 * C * es, halo0 * C * es,
                                  cudaMemcpyHostToDevice, d0.s_h2d));
    std::vector<uint32_t> launches(count, 0);
    for (uint64_t i = 0; i < nslices; ++i) {
        const uint64_t f0 = i * slice_frames;
        const uint64_t fcount = std::min<uint64_t>(slice_frames, frames - f0);
        const size_t off = (size_t)f0 * C * es, bytes = (size_t)fcount * C * es;
        cudaEvent_t e_in = d0.pool[1 + i * (1 + count)];
        MAVG_CUDA(cudaMemcpyAsync(dev_in + off, (const char*)h_in + off, bytes, cudaMemcpyHostToDevice, d0.s_h2d));
        MAVG_CUDA(cudaEventRecord(e_in, d0.s_h2d));
        if (i + 1 == nslices && p0->timing_on) cudaEventRecord(d0.ev[1], d0.s_h2d);
        for (uint32_t j = 0; j < count; ++j) {
            mavg_plan* p = plans[j];
            DevCtx& d = p->dev[0];
            cudaEvent_t e_k = d0.pool[2 + i * (1 + count) + j];
            MAVG_CUDA(cudaStreamWaitEvent(d.stream, e_in, 0));
            const void* halo = (i == 0) ? (has_halo ? dev_in - p->halo_frames * C * es : nullptr)
                                        : dev_in + off - p->halo_frames * C * es;
            MAVG_TRY(launch_shard(p, d, dev_in + off, (char*)d.d_out + off, halo, fcount, &launches[j]));
            MAVG_CUDA(cudaEventRecord(e_k, d.stream));
            if (i + 1 == nslices && j + 1 == count && p0->timing_on) cudaEventRecord(d0.ev[2], d.stream);
            MAVG_CUDA(cudaStreamWaitEvent(d0.s_d2h, e_k, 0));
            MAVG_CUDA(cudaMemcpyAsync((char*)h_out[j] + off, (char*)d.d_out + off, bytes, cudaMemcpyDeviceToHost, d0.s_d2h));
        }
    }
    if (p0->timing_on) cudaEventRecord(d0.ev[3], d0.s_d2h);
    d0.timed = p0->timing_on;
    for (uint32_t j = 0; j < count; ++j) {
        plans[j]->launches_last_run = launches[j];
        if (j > 0) plans[j]->dev[0].timed = false;
        MAVG_CUDA(cudaStreamSynchronize(plans[j]->dev[0].stream));
    }
    MAVG_CUDA(cudaStreamSynchronize(d0.s_h2d));
    MAVG_CUDA(cudaStreamSynchronize(d0.s_d2h));
    return gather_timing(p0);
}
}  // namespace

int mavg_run_host_sweep(mavg_plan* const* plans, uint32_t count, const void* h_in, void* const* h_out)
{
    if (!plans || !h_in || !h_out || count == 0) return fail(MAVG_ERR_INVALID_ARG, "null argument or empty sweep");
    for (uint32_t i = 0; i < count; ++i) {
        if (!plans[i] || !h_out[i]) return fail(MAVG_ERR_INVALID_ARG, "plan or output %u is null", i);
        if (h_out[i] == h_in) return fail(MAVG_ERR_INVALID_ARG, "output must not alias input");
        for (uint32_t j = 0; j < i; ++j)
            if (plans[j] == plans[i] || h_out[j] == h_out[i])
                return fail(MAVG_ERR_INVALID_ARG, "plans and outputs of a sweep must be distinct");
    }
    DeviceGuard guard;
    if (count == 1 || !sweep_compatible(plans, count, h_out)) {
        for (uint32_t i = 0; i < count; ++i) MAVG_TRY(mavg_run_host(plans[i], h_in, h_out[i]));
        return MAVG_OK;
    }
    const int rc = run_host_sweep_impl(plans, count, h_in, h_out);
    if (rc != MAVG_OK) {
        // drain what is queued on the caller's buffers before they are handed back; the first error stays the one reported
        const std::string first = g_last_error;
        DevCtx& d0 = plans[0]->dev[0];
        if (cudaSetDevice(d0.device) == cudaSuccess) {
            if (d0.s_h2d) cudaStreamSynchronize(d0.s_h2d);
            for (uint32_t i = 0; i < count; ++i) cudaStreamSynchronize(plans[i]->dev[0].stream);
            if (d0.s_d2h) cudaStreamSynchronize(d0.s_d2h);
        }
        cudaGetLastError();
        g_last_error = first;
    }
    return rc;
}

int mavg_fill_synthetic_device(void* d_dst, int dtype, uint64_t n, uint64_t first_index, uint64_t seed, int dist,
                               void* cuda_stream)
{
    if (!d_dst && n) return fail(MAVG_ERR_INVALID_ARG, "null destination");
    if (dtype != MAVG_F32 && dtype != MAVG_I16) return fail(MAVG_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    if (dist < 0 || dist > MAVG_DIST_DC1E4) return fail(MAVG_ERR_INVALID_ARG, "unknown distribution %d", dist);
    if (n == 0) return MAVG_OK;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == MAVG_F32)
        mavg::fill_f32_kernel<<<blocks, 256, 0, st>>>((float*)d_dst, n, first_index, seed, dist);
    else
        mavg::fill_i16_kernel<<<blocks, 256, 0, st>>>((int16_t*)d_dst, n, first_index, seed);
    MAVG_CUDA(cudaGetLastError());
    return MAVG_OK;
}

int mavg_fill_synthetic(mavg_plan* p, uint64_t seed, int dist)
{
    if (!p) return fail(MAVG_ERR_INVALID_ARG, "plan is null");
    DeviceGuard guard;
    for (DevCtx& d : p->dev) {
        MAVG_TRY(alloc_owned(p, d));
        MAVG_CUDA(cudaSetDevice(d.device));
        uint64_t first;
        if (planar_batch(p)) first = (uint64_t)d.first_channel * p->desc.frames;
        else first = (p->desc.first_frame + d.first_frame) * p->desc.channels;
        MAVG_TRY(mavg_fill_synthetic_device(d.d_in, p->desc.dtype, shard_elems(p, d), first, seed, dist, d.stream));
    }
    return MAVG_OK;
}

int mavg_ipc_export(const void* d_ptr, void* handle64)
{
    if (!d_ptr || !handle64) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    MAVG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle64, &h, 64);
    return MAVG_OK;
}

int mavg_ipc_open(const void* handle64, void** d_ptr)
{
    if (!d_ptr || !handle64) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    MAVG_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MAVG_OK;
}

int mavg_ipc_close(void* d_ptr)
{
    if (!d_ptr) return MAVG_OK;
    MAVG_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return MAVG_OK;
}

int mavg_device_alloc(uint64_t bytes, void** d_ptr)
{
    if (!d_ptr) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (cudaMalloc(d_ptr, bytes ? bytes : 256) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        return fail(e == cudaErrorNoDevice ? MAVG_ERR_NO_DEVICE : MAVG_ERR_ALLOC, "cudaMalloc(%llu) failed: %s",
                    (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return MAVG_OK;
}

int mavg_device_free(void* d_ptr)
{
    if (d_ptr) MAVG_CUDA(cudaFree(d_ptr));
    return MAVG_OK;
}

int mavg_prefix_sum(int dtype, const void* d_in, void* d_out, uint64_t frames, uint32_t channels, void* cuda_stream)
{
    if (dtype != MAVG_F32 && dtype != MAVG_I16) return fail(MAVG_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    if (channels < 1 || channels > 8)
        return fail(MAVG_ERR_UNSUPPORTED, "mavg_prefix_sum takes 1 to 8 interleaved channels (got %u)", channels);
    if (frames == 0) return MAVG_OK;
    if (!d_in || !d_out) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (mavg_device_count() <= 0) return fail(MAVG_ERR_NO_DEVICE, "no CUDA device available: libmavg has no CPU fallback");
    const uint64_t n = frames * channels;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == MAVG_I16) return launch_scan<int16_t, int, long long>(d_in, d_out, n, channels, st);
    return launch_scan<float, double, double>(d_in, d_out, n, channels, st);
}

int mavg_host_alloc(uint64_t bytes, void** h_ptr)
{
    if (!h_ptr) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (cudaHostAlloc(h_ptr, bytes ? bytes : 64, cudaHostAllocPortable) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MAVG_ERR_NO_DEVICE : MAVG_ERR_ALLOC,
                    "cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return MAVG_OK;
}

int mavg_host_free(void* h_ptr)
{
    if (h_ptr) MAVG_CUDA(cudaFreeHost(h_ptr));
    return MAVG_OK;
}

int mavg_host_register(void* h_ptr, uint64_t bytes)
{
    if (!h_ptr) return fail(MAVG_ERR_INVALID_ARG, "null argument");
    if (bytes == 0) return MAVG_OK;
    const cudaError_t e = cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return MAVG_OK;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MAVG_ERR_NO_DEVICE : MAVG_ERR_ALLOC,
                    "cudaHostRegister(%llu bytes) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return MAVG_OK;
}

int mavg_host_unregister(void* h_ptr)
{
    if (!h_ptr) return MAVG_OK;
    const cudaError_t e = cudaHostUnregister(h_ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return MAVG_OK;
    }
    MAVG_CUDA(e);
    return MAVG_OK;
}

}  // extern "C"
