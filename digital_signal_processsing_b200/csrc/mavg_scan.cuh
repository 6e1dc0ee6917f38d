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

template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d)
{
    if constexpr (sizeof(T) == 4) return __shfl_up_sync(0xffffffffu, v, d);
    else return shfl_up_acc<T>(v, d);
}

constexpr int kScanChunkBytes = 65536;   // shared-memory bytes of chunk-local prefixes per CTA

// One CTA = one chunk of E = kScanChunkBytes / sizeof(TLoc) elements (8192 for int16 input, 4096 for float32).
//   1. coalesced 16-byte loads, converted to the chunk-local type TLoc (int32 is enough for an int16 chunk:
//      8192 * 32768 < 2^31; double for float32) into a padded shared-memory array;
//   2. every thread scans its run of R = E / 256 elements in place (per channel), then warp-shuffle + block
//      scan of the run totals gives per-thread offsets and the chunk aggregate;
//   3. the aggregate is published at once (long before any output is written), warp 0 resolves the chunk's
//      exclusive prefix by decoupled look-back over chunk descriptors while other CTAs of the SM keep working;
//   4. striped output pass: out[e] = prefix + thread offset + local prefix, 16-byte stores, fully coalesced.
// Each element is read from HBM once and written once; chunk ids come from an atomic ticket so a chunk only
// ever waits on chunks that are already running.  Scratch: ticket, status[chunks], aggr/pref[chunks][C].
template <typename TIn, typename TLoc, typename TAcc, int C>
__global__ void __launch_bounds__(256)
    scan_lookback_kernel(const TIn* __restrict__ in, TAcc* __restrict__ out, uint64_t n, uint32_t* __restrict__ ticket,
                         uint32_t* __restrict__ status, TAcc* __restrict__ aggr, TAcc* __restrict__ pref)
{
    constexpr int NT = 256, NW = NT / 32;
    constexpr int E = kScanChunkBytes / (int)sizeof(TLoc);      // elements per chunk
    constexpr int R = E / NT;                         // run per thread (32 or 16)
    constexpr int VE = 16 / (int)sizeof(TIn);         // input elements per 16-byte load
    static_assert(R % C == 0 && R % 16 == 0, "a run holds whole frames");
    extern __shared__ __align__(16) uint8_t scan_smem[];
    TLoc* loc = reinterpret_cast<TLoc*>(scan_smem);                    // [E + E/32], index e + (e >> 5)
    TLoc* toff = loc + (E + E / 32);                                   // [NT][C] exclusive offset of each run
    TLoc* s_warp = toff + NT * C;                                      // [NW][C]
    __shared__ TAcc s_excl[C];           // exclusive prefix of the chunk, broadcast from warp 0
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t cbase = (uint64_t)tile * E;
    auto pidx = [](int e) { return e + (e >> 5); };

    // ---- 1. load + convert (zero past the end)
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
#pragma unroll 2
    for (int q = tid; q < E / VE; q += NT) {
        const uint64_t e0 = cbase + (uint64_t)q * VE;
        const int l0 = q * VE;
        if (vec_ok && e0 + VE <= n) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + e0));
            if constexpr (sizeof(TIn) == 2) {
                const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    loc[pidx(l0 + 2 * i)] = (TLoc)((int)(w[i] << 16) >> 16);
                    loc[pidx(l0 + 2 * i + 1)] = (TLoc)((int)w[i] >> 16);
                }
            } else {
                loc[pidx(l0 + 0)] = (TLoc)__uint_as_float(raw.x);
                loc[pidx(l0 + 1)] = (TLoc)__uint_as_float(raw.y);
                loc[pidx(l0 + 2)] = (TLoc)__uint_as_float(raw.z);
                loc[pidx(l0 + 3)] = (TLoc)__uint_as_float(raw.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < VE; ++i) loc[pidx(l0 + i)] = (e0 + i < n) ? (TLoc)in[e0 + i] : (TLoc)0;
        }
    }
    __syncthreads();

    // ---- 2. in-place scan of the own run, 16 elements at a time, one carry per channel
    TLoc carry[C];
#pragma unroll
    for (int c = 0; c < C; ++c) carry[c] = 0;
    for (int b = 0; b < R; b += 16) {
        TLoc v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = loc[pidx(tid * R + b + i)];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            carry[i % C] += v[i];
            v[i] = carry[i % C];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) loc[pidx(tid * R + b + i)] = v[i];
    }
    TLoc inc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) inc[c] = carry[c];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const TLoc up = shfl_up_t<TLoc>(inc[c], d);
            if (lane >= d) inc[c] += up;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < C; ++c) s_warp[warp * C + c] = inc[c];
    }
    __syncthreads();
    TAcc chunk_tot[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        TLoc a = 0;
        for (int w2 = 0; w2 < warp; ++w2) a += s_warp[w2 * C + c];
        toff[tid * C + c] = a + (inc[c] - carry[c]);         // exclusive offset of this run inside the chunk
        TLoc t = a;
        for (int w2 = warp; w2 < NW; ++w2) t += s_warp[w2 * C + c];
        chunk_tot[c] = (TAcc)t;
    }

    // ---- 3. publish the aggregate, look back for the exclusive prefix (warp 0), publish the inclusive prefix.
    // (A CTA-wide look-back, 256 predecessors per round, was tried: it was slower -- 1.0 vs 0.87 ms on 2^28
    // int16 samples -- because the wait is for predecessors to publish at all, not for the number of rounds.)
    if (warp == 0) {
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) aggr[(uint64_t)tile * C + c] = chunk_tot[c];
            if (tile == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) pref[c] = chunk_tot[c];
            }
            __threadfence();
            st_release_u32(status + tile, tile == 0 ? kScanPrefix : kScanAggregate);
        }
        TAcc excl[C];
#pragma unroll
        for (int c = 0; c < C; ++c) excl[c] = 0;
        if (tile > 0) {
            // each lane inspects one predecessor; windows of 32 chunks move backwards until a prefix is found
            long long look = (long long)tile - 1 - lane;
            for (;;) {
                uint32_t st = kScanPrefix;   // lanes before chunk 0 behave like a terminating zero prefix
                if (look >= 0) {
                    do { st = ld_acquire_u32(status + look); } while (st == kScanInvalid);
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, st == kScanPrefix);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest predecessor with a full prefix
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    TAcc val = 0;
                    if (look >= 0 && lane <= first)
                        val = (st == kScanPrefix) ? __ldcg(pref + (uint64_t)look * C + c) : __ldcg(aggr + (uint64_t)look * C + c);
                    // fixed-shape butterfly: the same association for a given `first`
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        const long long ob = acc_bits<TAcc>(val);
                        const int lo = __shfl_xor_sync(0xffffffffu, (int)(ob & 0xffffffffll), d);
                        const int hi = __shfl_xor_sync(0xffffffffu, (int)(ob >> 32), d);
                        val += acc_from_bits<TAcc>(((long long)hi << 32) | (unsigned int)lo);
                    }
                    excl[c] += val;
                }
                if (has_prefix) break;
                look -= 32;
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) pref[(uint64_t)tile * C + c] = excl[c] + chunk_tot[c];
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

    // ---- 4. striped, coalesced output: two results per 16-byte store
    TAcc base[C];
#pragma unroll
    for (int c = 0; c < C; ++c) base[c] = s_excl[c];
    const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
#pragma unroll 4
    for (int q = tid; q < E / 2; q += NT) {
        const int e = 2 * q;
        const uint64_t g = cbase + (uint64_t)e;
        if (g >= n) break;
        const int run = e / R;                       // e and e + 1 lie in the same run (R is even)
        const int c0 = e % C, c1 = (e + 1) % C;
        const TAcc a = base[c0] + (TAcc)(toff[run * C + c0] + loc[pidx(e)]);
        const TAcc b = base[c1] + (TAcc)(toff[run * C + c1] + loc[pidx(e + 1)]);
        if (out_vec && g + 1 < n) {
            *reinterpret_cast<double2*>(out + g) = make_double2(*reinterpret_cast<const double*>(&a),
                                                                *reinterpret_cast<const double*>(&b));
        } else {
            out[g] = a;
            if (g + 1 < n) out[g + 1] = b;
        }
    }
}

// bytes of dynamic shared memory scan_lookback_kernel needs
template <typename TLoc, int C>
constexpr uint32_t scan_smem_bytes()
{
    constexpr uint32_t E = kScanChunkBytes / sizeof(TLoc);
    return (E + E / 32) * sizeof(TLoc) + 256 * C * sizeof(TLoc) + (8 * C + 1) * sizeof(TLoc) + C * 8 + 16;
}

}  // namespace mavg
