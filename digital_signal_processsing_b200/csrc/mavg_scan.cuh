// mavg_scan.cuh -- device-wide per-channel inclusive prefix sum, one pass, decoupled look-back.
//
// This is the reusable primitive behind the reference's scan binaries: recursive_hillis_steele +
// hillis_steele + uniform_add (basics/hillis_steele_averager.cu:16-84) and recursive_blelloch +
// blelloch_scan_inclusive + blelloch_uniform_add (basics/blelloch_scan_averager.cu:16-167) compute
// exactly this -- the in-place int64 inclusive prefix of the interleaved signal -- with a multi-level
// recursion, three full passes over 8-byte data and an aux array per level.  Here: one kernel, each tile
// read once and written once; tiles publish {aggregate, inclusive prefix} descriptors and a tile resolves
// its exclusive prefix by looking back over its predecessors' descriptors (Merrill & Garland's scheme),
// with tile ids drawn from an atomic ticket so that a tile only ever waits on tiles that already run.
// The moving-average kernels do NOT use it: a window needs only the previous ceil(k/T) tiles, which they
// keep in shared memory, so no global prefix (and no fp32 cancellation across the signal) is needed.
//
// int16 -> int64 (exact), float32 -> float64.  Interleaved channels: element r of a thread's 16-element run
// belongs to channel r % C (C divides 16), so every thread carries C running sums.
// Traffic: sizeof(in) + 8 bytes per sample.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace mavg {

enum : uint32_t { kScanInvalid = 0u, kScanAggregate = 1u, kScanPrefix = 2u };

// 64-bit shuffles (long long or double accumulators) as two 32-bit halves
template <typename TAcc>
__device__ __forceinline__ long long acc_bits(TAcc v)
{
    if constexpr (std::is_floating_point<TAcc>::value) return __double_as_longlong((double)v);
    else return (long long)v;
}
template <typename TAcc>
__device__ __forceinline__ TAcc acc_from_bits(long long b)
{
    if constexpr (std::is_floating_point<TAcc>::value) return (TAcc)__longlong_as_double(b);
    else return (TAcc)b;
}
template <typename TAcc>
__device__ __forceinline__ TAcc shfl_up_acc(TAcc v, int d)
{
    const long long b = acc_bits<TAcc>(v);
    const int lo = __shfl_up_sync(0xffffffffu, (int)(b & 0xffffffffll), d);
    const int hi = __shfl_up_sync(0xffffffffu, (int)(b >> 32), d);
    return acc_from_bits<TAcc>(((long long)hi << 32) | (unsigned int)lo);
}
template <typename TAcc>
__device__ __forceinline__ TAcc shfl_idx_acc(TAcc v, int src)
{
    const long long b = acc_bits<TAcc>(v);
    const int lo = __shfl_sync(0xffffffffu, (int)(b & 0xffffffffll), src);
    const int hi = __shfl_sync(0xffffffffu, (int)(b >> 32), src);
    return acc_from_bits<TAcc>(((long long)hi << 32) | (unsigned int)lo);
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// scratch layout: [0] ticket counter, then per tile: status word, then aggregates [tile][C], then prefixes [tile][C]
template <typename TIn, typename TAcc, int C>
__global__ void __launch_bounds__(256)
    scan_lookback_kernel(const TIn* __restrict__ in, TAcc* __restrict__ out, uint64_t n, uint32_t* __restrict__ ticket,
                         uint32_t* __restrict__ status, TAcc* __restrict__ aggr, TAcc* __restrict__ pref)
{
    constexpr int NT = 256, R = 16, T = NT * R, NW = NT / 32;
    static_assert(R % C == 0, "a run holds whole frames");
    __shared__ uint32_t s_tile;
    __shared__ TAcc s_warp[NW][C];
    __shared__ TAcc s_excl[C];
    __shared__ __align__(16) TAcc s_stage[NT * (R + 2)];   // 16 results + 16 bytes of padding per thread row

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * T + (uint64_t)tid * R;

    // ---- load the run (zero past the end) and scan it per channel
    TAcc v[R];
    if (base + R <= n && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
        if constexpr (sizeof(TIn) == 2) {
            const uint4* p = reinterpret_cast<const uint4*>(in + base);
            const uint4 a = __ldg(p), b = __ldg(p + 1);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                v[2 * i] = (TAcc)((int)(w[i] << 16) >> 16);
                v[2 * i + 1] = (TAcc)((int)w[i] >> 16);
            }
        } else {
            const float4* p = reinterpret_cast<const float4*>(in + base);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 a = __ldg(p + i);
                v[4 * i] = (TAcc)a.x; v[4 * i + 1] = (TAcc)a.y; v[4 * i + 2] = (TAcc)a.z; v[4 * i + 3] = (TAcc)a.w;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = (base + i < n) ? (TAcc)in[base + i] : (TAcc)0;
    }
#pragma unroll
    for (int i = C; i < R; ++i) v[i] += v[i - C];        // inclusive per-channel scan inside the run
    TAcc tot[C], inc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) inc[c] = tot[c] = v[R - C + c];

    // ---- warp scan of run totals, then block scan through shared memory
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const TAcc up = shfl_up_acc<TAcc>(inc[c], d);
            if (lane >= d) inc[c] += up;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < C; ++c) s_warp[warp][c] = inc[c];
    }
    __syncthreads();
    TAcc woff[C], tile_tot[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        TAcc a = 0;
        for (int w2 = 0; w2 < warp; ++w2) a += s_warp[w2][c];
        woff[c] = a;
        TAcc t = a;
        for (int w2 = warp; w2 < NW; ++w2) t += s_warp[w2][c];
        tile_tot[c] = t;
    }

    // ---- publish the aggregate, look back for the exclusive prefix (warp 0), publish the inclusive prefix
    if (warp == 0) {
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) aggr[(uint64_t)tile * C + c] = tile_tot[c];
            if (tile == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) pref[c] = tile_tot[c];
            }
            __threadfence();
            st_release_u32(status + tile, tile == 0 ? kScanPrefix : kScanAggregate);
        }
        TAcc excl[C];
#pragma unroll
        for (int c = 0; c < C; ++c) excl[c] = 0;
        if (tile > 0) {
            // each lane inspects one predecessor; windows of 32 tiles move backwards until a prefix is found
            long long look = (long long)tile - 1 - lane;
            for (;;) {
                uint32_t st = kScanPrefix;   // lanes before tile 0 behave like a terminating zero prefix
                if (look >= 0) {
                    do { st = ld_acquire_u32(status + look); } while (st == kScanInvalid);
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, st == kScanPrefix);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest predecessor with a full prefix
                TAcc part[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    TAcc val = 0;
                    if (look >= 0 && lane <= first)
                        val = (st == kScanPrefix) ? __ldcg(pref + (uint64_t)look * C + c) : __ldcg(aggr + (uint64_t)look * C + c);
                    part[c] = val;
                }
                // sum over lanes 0..first in a fixed order (lane `first` outermost) so results are deterministic
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    TAcc s = 0;
                    for (int l = 31; l >= 0; --l) {
                        const TAcc t = shfl_idx_acc<TAcc>(part[c], l);
                        if (l <= first) s += t;
                    }
                    excl[c] += s;
                }
                if (has_prefix) break;
                look -= 32;
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) pref[(uint64_t)tile * C + c] = excl[c] + tile_tot[c];
                __threadfence();
                st_release_u32(status + tile, kScanPrefix);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) s_excl[c] = excl[c];
        }
    }
    __syncthreads();

    // ---- add the offsets, stage through padded shared memory, store coalesced
    TAcc off[C];
#pragma unroll
    for (int c = 0; c < C; ++c) off[c] = s_excl[c] + woff[c] + (inc[c] - tot[c]);
#pragma unroll
    for (int i = 0; i < R; ++i) s_stage[tid * (R + 2) + i] = v[i] + off[i % C];
    __syncthreads();
    const uint64_t tile_base = (uint64_t)tile * T;
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
        const int chunk = j * NT + tid;                   // 16-byte chunk = 2 results, consecutive across the CTA
        const int row = chunk >> 3, col = chunk & 7;
        const uint64_t e = tile_base + (uint64_t)chunk * 2;
        const TAcc a = s_stage[row * (R + 2) + col * 2], b = s_stage[row * (R + 2) + col * 2 + 1];
        if (e + 1 < n && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
            *reinterpret_cast<double2*>(out + e) = make_double2(*reinterpret_cast<const double*>(&a),
                                                                *reinterpret_cast<const double*>(&b));
        } else {
            if (e < n) out[e] = a;
            if (e + 1 < n) out[e + 1] = b;
        }
    }
}

}  // namespace mavg
