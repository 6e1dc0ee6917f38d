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
// int16 -> int64 (exact), float32 -> float64.  Interleaved channels, any count from 1 to 8: a thread's run is a whole
// number of frames (its length is a multiple of lcm(C, 2)), so element r of the run belongs to channel r % C at
// compile time and every thread carries C running sums.
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
__device__ __forceinline__ TAcc shfl_xor_acc(TAcc v, int d)
{
    const long long b = acc_bits<TAcc>(v);
    const int lo = __shfl_xor_sync(0xffffffffu, (int)(b & 0xffffffffll), d);
    const int hi = __shfl_xor_sync(0xffffffffu, (int)(b >> 32), d);
    return acc_from_bits<TAcc>(((long long)hi << 32) | (unsigned int)lo);
}
template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d)
{
    if constexpr (sizeof(T) == 4) return __shfl_up_sync(0xffffffffu, v, d);
    else return shfl_up_acc<T>(v, d);
}
template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int d)
{
    if constexpr (sizeof(T) == 4) return __shfl_xor_sync(0xffffffffu, v, d);
    else return shfl_xor_acc<T>(v, d);
}

// Chunk descriptor {status, value bits}: ONE 16-byte word, written and read with single 128-bit accesses (the scheme of
// CUB's ScanTileState for 8-byte values), so the value travels with its flag -- no fence between a value store and a
// flag store, no second dependent load on the reading side.  That halves the latency of one hop of the look-back
// chain (round 1 used separate status / aggregate / prefix arrays: flag acquire -> value load -> fence -> flag release,
// ~1.7 us per hop, which capped the primitive at 32 chunks per hop = 3.1 TB/s).
__device__ __forceinline__ ulonglong2 ld_desc(const ulonglong2* p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(ulonglong2* p, unsigned long long status, unsigned long long bits)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(status), "l"(bits) : "memory");
}

constexpr int kScanThreads = 256;

// Run length per thread: the largest multiple of lcm(C, 2) not above kScanChunkBytes / sizeof(TLoc) / 256, so that a
// run holds whole frames (the channel of run element i is i % C at compile time) and an even number of elements
// (two results per 16-byte store).  C = 1, 2, 4, 8: 64 (int16) / 32 (float32); C = 3, 5, 6: 60 / 30; C = 7: 56 / 28.
// CB = shared-memory bytes of chunk-local prefixes per CTA (before padding)
template <typename TLoc, int C, int CB>
__host__ __device__ constexpr int scan_run_len()
{
    constexpr int rmax = CB / (int)sizeof(TLoc) / kScanThreads;
    constexpr int m = (C % 2 == 0) ? C : 2 * C;
    return rmax / m * m;
}
template <typename TLoc, int C, int CB>
__host__ __device__ constexpr int scan_chunk_elems() { return kScanThreads * scan_run_len<TLoc, C, CB>(); }
// register block of the in-place run scan: the largest divisor of the run that is a multiple of C and at most 16
template <int R, int C>
__host__ __device__ constexpr int scan_block_len()
{
    int best = C;
    for (int b = C; b <= 16; b += C)
        if (R % b == 0) best = b;
    return best;
}

// Decoupled look-back of chunk `tile`, executed by ONE warp: each lane inspects one predecessor per round; rounds of
// 32 chunks move backwards until every channel has met a full prefix.  All C descriptors of a predecessor are
// requested together and the next round's are requested before this round's are examined, so a round costs one
// memory latency however many channels.  Writes the chunk's exclusive prefix to s_excl[C] (shared) and publishes
// {PREFIX, exclusive + s_tot[c]}.
template <typename TAcc, int C>
__device__ __forceinline__ void scan_look_back(ulonglong2* __restrict__ desc, uint32_t tile, int lane, TAcc* s_excl,
                                               const TAcc* s_tot)
{
    ulonglong2* my_desc = desc + (uint64_t)tile * C;
    if (tile == 0) {
        if (lane < C) s_excl[lane] = 0;
        return;
    }
    long long look = (long long)tile - 1 - lane;
    const ulonglong2 stop = make_ulonglong2(kScanPrefix, 0ull);   // lanes before chunk 0: a terminating zero prefix
    ulonglong2 cur[C], nxt[C];
    TAcc acc[C];
    bool open_[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        cur[c] = look >= 0 ? ld_desc(desc + (uint64_t)look * C + c) : stop;
        acc[c] = 0;
        open_[c] = true;
    }
    for (;;) {
#pragma unroll
        for (int c = 0; c < C; ++c) nxt[c] = look - 32 >= 0 ? ld_desc(desc + (uint64_t)(look - 32) * C + c) : stop;
        bool any_open = false;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (!open_[c]) continue;                                       // warp-uniform
            while (cur[c].x == kScanInvalid) cur[c] = ld_desc(desc + (uint64_t)look * C + c);
            const unsigned has_prefix = __ballot_sync(0xffffffffu, cur[c].x == kScanPrefix);
            const int first = has_prefix ? __ffs(has_prefix) - 1 : 32;   // nearest predecessor with a full prefix
            TAcc val = (lane <= first) ? acc_from_bits<TAcc>((long long)cur[c].y) : (TAcc)0;
            // fixed-shape butterfly: the same association for a given `first`
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) val += shfl_xor_acc<TAcc>(val, d);
            acc[c] += val;
            open_[c] = has_prefix == 0;
            any_open = any_open || open_[c];
        }
        if (!any_open) break;
        look -= 32;
#pragma unroll
        for (int c = 0; c < C; ++c) cur[c] = nxt[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (lane == 0) {
            s_excl[c] = acc[c];
            st_desc(my_desc + c, kScanPrefix, (unsigned long long)acc_bits<TAcc>(acc[c] + s_tot[c]));
        }
    }
}

// One CTA = one chunk of E = 256 * R elements (16384 for int16 input, 8192 for float32, a little less for odd C).
//   1. coalesced 16-byte loads, converted to the chunk-local type TLoc (int32 is enough for an int16 chunk:
//      16384 * 32768 < 2^31; double for float32) into a padded shared-memory array.  Where the channel of a loaded
//      element is known at compile time (C divides the vector width) the chunk AGGREGATE is reduced on the way and
//      published right behind the load barrier -- a whole run scan earlier than the look-back needs its
//      predecessors' aggregates, so the look-back no longer spins on unpublished descriptors;
//   2. every thread scans its run in place (per channel), then warp-shuffle + block scan of the run totals gives
//      per-thread offsets (and, for the other channel counts, the aggregate);
//   3. warp 0 resolves the chunk's exclusive prefix by decoupled look-back over the 16-byte chunk descriptors, 32
//      predecessors per round, the next round's descriptors already in flight; publishes the inclusive prefix;
//   4. striped output pass: out[e] = prefix + thread offset + local prefix, 16-byte stores, fully coalesced.
// Each element is read from HBM once and written once; chunk ids come from an atomic ticket so a chunk only
// ever waits on chunks that are already running.  Scratch: ticket, desc[chunks][C].
template <typename TIn, typename TLoc, typename TAcc, int C, int CB>
__global__ void __launch_bounds__(kScanThreads)
    scan_lookback_kernel(const TIn* __restrict__ in, TAcc* __restrict__ out, uint64_t n, uint32_t tile_base,
                         ulonglong2* __restrict__ desc)
{
    constexpr int NT = kScanThreads, NW = NT / 32;
    constexpr int R = scan_run_len<TLoc, C, CB>();    // run per thread
    constexpr int E = NT * R;                         // elements per chunk
    constexpr int BL = scan_block_len<R, C>();
    constexpr int VE = 16 / (int)sizeof(TIn);         // input elements per 16-byte load
    constexpr bool EARLY = (VE % C == 0);             // channel of a loaded element known at compile time
    static_assert(R % C == 0 && R % 2 == 0 && R % BL == 0 && BL % C == 0 && E % VE == 0, "a run holds whole frames");
    extern __shared__ __align__(16) uint8_t scan_smem[];
    TLoc* loc = reinterpret_cast<TLoc*>(scan_smem);                    // [E + E/32], index e + (e >> 5)
    TLoc* toff = loc + (E + E / 32);                                   // [NT][C] exclusive offset of each run
    TLoc* s_warp = toff + NT * C;                                      // [NW][C] scan: warp totals
    TLoc* s_part = s_warp + NW * C;                                    // [NW][C] early aggregate: warp partial sums
    __shared__ TAcc s_excl[C];           // exclusive prefix of the chunk, broadcast from warp 0
    __shared__ TAcc s_tot[C];            // chunk aggregate as published

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Chunk id = block index: the hardware hands out the blocks of a 1-D grid in increasing order, so a chunk only ever
    // waits on chunks that are already resident or done (what CUB's DeviceScan relies on as well).  Round 1 drew the id
    // from an atomic ticket: one more global round trip and a barrier in front of every chunk's loads.
    // tile_base: the int16 fast kernel below takes the whole chunks, this one the ragged last chunk of the same chain.
    const uint32_t tile = blockIdx.x + tile_base;
    const uint64_t cbase = (uint64_t)tile * E;
    // Where element e of the chunk lives in `loc`.  4-byte prefixes: an XOR swizzle of the bank bits with bits 5..10 of e
    // makes all three access patterns conflict-free -- the striped load phase (lanes 8 elements apart), the run
    // scan (lanes one run = 64 elements apart) and the striped output phase (lanes 2 apart); the padded layout
    // e + e/32 of round 1 was 2-way conflicted in the run scan (ncu: 35 % of shared wavefronts were conflicts).
    // 8-byte prefixes (float32 input) keep the padding.
    auto pidx = [](int e) {
        if constexpr (sizeof(TLoc) == 4) return e ^ ((e >> 5) & 31) ^ ((e >> 10) & 1);
        else return e + (e >> 5);
    };
    ulonglong2* my_desc = desc + (uint64_t)tile * C;

    // ---- 1. load + convert (zero past the end); early per-channel partial sums
    TLoc csum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) csum[c] = 0;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
#pragma unroll 4
    for (int q = tid; q < E / VE; q += NT) {
        const uint64_t e0 = cbase + (uint64_t)q * VE;
        const int l0 = q * VE;
        TLoc v[VE];
        if (vec_ok && e0 + VE <= n) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(in + e0));
            if constexpr (sizeof(TIn) == 2) {
                const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[2 * i] = (TLoc)((int)(w[i] << 16) >> 16);
                    v[2 * i + 1] = (TLoc)((int)w[i] >> 16);
                }
            } else {
                v[0] = (TLoc)__uint_as_float(raw.x);
                v[1] = (TLoc)__uint_as_float(raw.y);
                v[2] = (TLoc)__uint_as_float(raw.z);
                v[3] = (TLoc)__uint_as_float(raw.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < VE; ++i) v[i] = (e0 + i < n) ? (TLoc)in[e0 + i] : (TLoc)0;
        }
#pragma unroll
        for (int i = 0; i < VE; ++i) {
            loc[pidx(l0 + i)] = v[i];
            if constexpr (EARLY) csum[i % C] += v[i];
        }
    }
    if constexpr (EARLY) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) csum[c] += shfl_xor_t<TLoc>(csum[c], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) s_part[warp * C + c] = csum[c];
        }
    }
    __syncthreads();
    if constexpr (EARLY) {
        if (tid < C) {   // fixed order over the warps: the aggregate does not depend on scheduling
            TLoc t = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) t += s_part[w2 * C + tid];
            const TAcc tot = (TAcc)t;
            s_tot[tid] = tot;
            st_desc(my_desc + tid, tile == 0 ? kScanPrefix : kScanAggregate, (unsigned long long)acc_bits<TAcc>(tot));
        }
    }

    // ---- 2. in-place scan of the own run, BL elements at a time, one carry per channel
    TLoc carry[C];
#pragma unroll
    for (int c = 0; c < C; ++c) carry[c] = 0;
    for (int b = 0; b < R; b += BL) {
        TLoc v[BL];
#pragma unroll
        for (int i = 0; i < BL; ++i) v[i] = loc[pidx(tid * R + b + i)];
#pragma unroll
        for (int i = 0; i < BL; ++i) {
            carry[i % C] += v[i];
            v[i] = carry[i % C];
        }
#pragma unroll
        for (int i = 0; i < BL; ++i) loc[pidx(tid * R + b + i)] = v[i];
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
#pragma unroll
    for (int c = 0; c < C; ++c) {
        TLoc a = 0;
        for (int w2 = 0; w2 < warp; ++w2) a += s_warp[w2 * C + c];
        toff[tid * C + c] = a + (inc[c] - carry[c]);         // exclusive offset of this run inside the chunk
    }
    if constexpr (!EARLY) {
        if (tid < C) {
            TLoc t = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) t += s_warp[w2 * C + tid];
            const TAcc tot = (TAcc)t;
            s_tot[tid] = tot;
            st_desc(my_desc + tid, tile == 0 ? kScanPrefix : kScanAggregate, (unsigned long long)acc_bits<TAcc>(tot));
        }
        __syncwarp();
    }

    // ---- 3. look back for the exclusive prefix (warp 0), publish the inclusive prefix
    if (warp == 0) scan_look_back<TAcc, C>(desc, tile, lane, s_excl, s_tot);
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
        TAcc b0 = 0, b1 = 0;                         // base[] with a run-time channel: selects, no local memory
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c == c0) b0 = base[c];
            if (c == c1) b1 = base[c];
        }
        const TAcc a = b0 + (TAcc)(toff[run * C + c0] + loc[pidx(e)]);
        const TAcc b = b1 + (TAcc)(toff[run * C + c1] + loc[pidx(e + 1)]);
        if (out_vec && g + 1 < n) {
            *reinterpret_cast<double2*>(out + g) = make_double2(*reinterpret_cast<const double*>(&a),
                                                                *reinterpret_cast<const double*>(&b));
        } else {
            out[g] = a;
            if (g + 1 < n) out[g + 1] = b;
        }
    }
}

// ----------------------------------------------------------------------------------
// int16 -> int64 fast kernel (C = 1, 2, 4, 8; whole chunks only).  Same chain, same descriptors, same chunk size as
// scan_lookback_kernel<short, int, long long, C, 32768> (8192 elements, runs of 32), so the ragged last chunk of a
// signal is handled by that kernel with tile_base = number of whole chunks.  The ncu source view of the general kernel
// showed 45 executed instructions per element, most of them address arithmetic around 4-byte shared-memory accesses
// (11 in the load phase, 10.5 in the run scan, 20 in the output phase).  Here every shared access is a vector:
//   * chunk-local prefixes live at  pad(e) = e + 4 * (e / 32)  ints -- 16 bytes of padding per 128, which keeps every
//     8-element load vector, every 32-element run and every output pair contiguous and 16-byte aligned, and makes the
//     three patterns conflict-free per quarter warp (load phase: lanes 32 B apart, run scan: lanes 144 B apart);
//   * load phase: one 16-byte global load -> two STS.128; the aggregate is a dp2a on the packed words;
//   * run scan: 8 LDS.128 + 32 adds + 8 STS.128 per thread;
//   * output: one LDS.64 (two prefixes), one LDS (run offset), one 16-byte global store per pair.
// All loops are fully unrolled with compile-time strides, so addresses are one register plus an immediate.
// ----------------------------------------------------------------------------------
__device__ __forceinline__ int scan_dp2a(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Traits of the fast kernel per input type: chunk-local / accumulator types, padding (TLoc elements per 32), the
// vector that holds two chunk-local prefixes.
template <typename TIn> struct ScanFast;
template <> struct ScanFast<int16_t> {
    typedef int TLoc;
    typedef long long TAcc;
    typedef int2 Pair;                 // two prefixes of the output pair (LDS.64)
    static constexpr int kPad = 4;     // 16 bytes of padding per 128
};
template <> struct ScanFast<float> {
    typedef double TLoc;
    typedef double TAcc;
    typedef double2 Pair;              // LDS.128
    static constexpr int kPad = 2;     // 16 bytes of padding per 256
};

// NCH chunks per CTA (1 or 2): the global loads of ALL of them are issued first, so the second chunk arrives while the
// first goes through its phases, and only the first chunk of a CTA looks back -- the next one knows its exclusive
// prefix (the predecessor is local) and publishes a full PREFIX descriptor right behind its load barrier.
template <typename TIn, int C, int R, int NCH = 1>
__global__ void __launch_bounds__(kScanThreads, (sizeof(TIn) == 2 && NCH == 2 && C == 1 ? 5 : 0))   // mono int16: 48 registers, five resident CTAs (0.463 -> 0.445 ms); stereo loses with the cap (0.451 -> 0.497)
    scan_lookback_fast_kernel(const TIn* __restrict__ in, typename ScanFast<TIn>::TAcc* __restrict__ out,
                              ulonglong2* __restrict__ desc)
{
    typedef typename ScanFast<TIn>::TLoc TLoc;
    typedef typename ScanFast<TIn>::TAcc TAcc;
    typedef typename ScanFast<TIn>::Pair Pair;
    constexpr bool I16 = sizeof(TIn) == 2;
    constexpr int NT = kScanThreads, NW = NT / 32, E = NT * R;   // elements per chunk
    constexpr int VE = 16 / (int)sizeof(TIn);                     // elements per 16-byte global load: 8 or 4
    constexpr int LV = R / VE;                                    // load vectors per thread
    constexpr int EV = 16 / (int)sizeof(TLoc);                    // chunk-local prefixes per 16-byte shared access: 4 or 2
    constexpr int PAD = ScanFast<TIn>::kPad;
    static_assert(C == 1 || C == 2 || C == 4 || (C == 8 && I16), "channel of a loaded element known at compile time");
    static_assert(R == 32 || R == 16, "run length");
    constexpr int LOC = E + E / 32 * PAD;              // padded TLoc elements
    extern __shared__ __align__(16) uint8_t scan_smem[];
    TLoc* loc = reinterpret_cast<TLoc*>(scan_smem);    // [LOC]
    TLoc* toff = loc + LOC;                            // [NT][C] exclusive offset of each run inside the chunk
    TLoc* s_warp = toff + NT * C;                      // [NW][C]
    TLoc* s_part = s_warp + NW * C;                    // [NW][C]
    __shared__ __align__(16) TAcc s_excl[C < 2 ? 2 : C];
    __shared__ TAcc s_tot[C];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- 1a. loads of every chunk of this CTA: thread handles vectors q = tid + 256 * it (VE elements each), it < LV
    uint4 raw_all[NCH][LV];
#pragma unroll
    for (int s = 0; s < NCH; ++s) {
        const TIn* cin = in + ((uint64_t)blockIdx.x * NCH + s) * E;
#pragma unroll
        for (int it = 0; it < LV; ++it) raw_all[s][it] = __ldg(reinterpret_cast<const uint4*>(cin) + tid + NT * it);
    }
    // The aggregates of the LATER chunks of this CTA are published before the first chunk is processed (reduced straight
    // from the registers, same order as in phase 1b, so the bits agree): chunk 2b + 1 would otherwise stay unpublished
    // until CTA b is through its first chunk -- look-back included -- and the chain of CTAs would run one after the other
    // (measured: 33 ms instead of 0.5 ms on 2^28 samples).
#pragma unroll
    for (int s = 1; s < NCH; ++s) {
        TLoc cs1[C];
#pragma unroll
        for (int c = 0; c < C; ++c) cs1[c] = 0;
#pragma unroll
        for (int it = 0; it < LV; ++it) {
            const uint32_t w[4] = {raw_all[s][it].x, raw_all[s][it].y, raw_all[s][it].z, raw_all[s][it].w};
            if constexpr (I16) {
                int cs[C];
#pragma unroll
                for (int c = 0; c < C; ++c) cs[c] = (int)cs1[c];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (C == 1) {
                        cs[0] = scan_dp2a(w[j], 0x0101u, cs[0]);
                    } else {
                        cs[(2 * j) % C] = scan_dp2a(w[j], 0x0001u, cs[(2 * j) % C]);
                        cs[(2 * j + 1) % C] = scan_dp2a(w[j], 0x0100u, cs[(2 * j + 1) % C]);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) cs1[c] = (TLoc)cs[c];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) cs1[j % C] += (TLoc)__uint_as_float(w[j]);
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) cs1[c] += shfl_xor_t<TLoc>(cs1[c], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < C; ++c) s_part[warp * C + c] = cs1[c];
        }
        __syncthreads();
        if (tid < C) {
            TLoc t = 0;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) t += s_part[w2 * C + tid];
            st_desc(desc + ((uint64_t)blockIdx.x * NCH + s) * C + tid, kScanAggregate,
                    (unsigned long long)acc_bits<TAcc>((TAcc)t));
        }
        __syncthreads();   // s_part is reused below
    }
#pragma unroll
    for (int s = 0; s < NCH; ++s) {
    const uint32_t tile = blockIdx.x * NCH + s;
    TAcc* cout = out + (uint64_t)tile * E;
    const uint4* raw = raw_all[s];

    // ---- 1b. unpack into shared memory, reduce the aggregate on the way
    TLoc csum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) csum[c] = 0;
    {
        // pad(VE q) for q = tid, then + it * pad(VE * 256)
        TLoc* dst = loc + VE * tid + PAD * ((VE * tid) >> 5);
#pragma unroll
        for (int it = 0; it < LV; ++it) {
            const uint32_t w[4] = {raw[it].x, raw[it].y, raw[it].z, raw[it].w};
            TLoc v[VE];
            if constexpr (I16) {
                int cs[C];
#pragma unroll
                for (int c = 0; c < C; ++c) cs[c] = (int)csum[c];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[2 * j] = (TLoc)(int)(short)(w[j] & 0xffffu);
                    v[2 * j + 1] = (TLoc)((int)w[j] >> 16);
                    if constexpr (C == 1) {
                        cs[0] = scan_dp2a(w[j], 0x0101u, cs[0]);
                    } else {
                        cs[(2 * j) % C] = scan_dp2a(w[j], 0x0001u, cs[(2 * j) % C]);
                        cs[(2 * j + 1) % C] = scan_dp2a(w[j], 0x0100u, cs[(2 * j + 1) % C]);
                    }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) csum[c] = (TLoc)cs[c];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = (TLoc)__uint_as_float(w[j]);
                    csum[j % C] += v[j];
                }
            }
            TLoc* d = dst + it * (VE * NT + PAD * (VE * NT / 32));
            if constexpr (I16) {
                int4* d4 = reinterpret_cast<int4*>(d);
                d4[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
                d4[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
            } else {
                double2* d2 = reinterpret_cast<double2*>(d);
                d2[0] = make_double2((double)v[0], (double)v[1]);
                d2[1] = make_double2((double)v[2], (double)v[3]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) csum[c] += shfl_xor_t<TLoc>(csum[c], d);
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) s_part[warp * C + c] = csum[c];
    }
    __syncthreads();
    ulonglong2* my_desc = desc + (uint64_t)tile * C;
    if (tid < C) {   // fixed order over the warps
        TLoc t = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2) t += s_part[w2 * C + tid];
        if (s == 0) {
            s_tot[tid] = (TAcc)t;
            st_desc(my_desc + tid, tile == 0 ? kScanPrefix : kScanAggregate, (unsigned long long)acc_bits<TAcc>((TAcc)t));
        } else {
            // the predecessor is this CTA's previous chunk: exclusive prefix = its exclusive prefix + its aggregate
            const TAcc ex = s_excl[tid] + s_tot[tid];
            s_excl[tid] = ex;
            s_tot[tid] = (TAcc)t;
            st_desc(my_desc + tid, kScanPrefix, (unsigned long long)acc_bits<TAcc>(ex + (TAcc)t));
        }
    }

    // ---- 2. in-place scan of the own run: R contiguous prefixes at pad(R tid)
    TLoc carry[C];
#pragma unroll
    for (int c = 0; c < C; ++c) carry[c] = 0;
    {
        TLoc* run = loc + R * tid + PAD * ((R * tid) >> 5);
#pragma unroll
        for (int b = 0; b < R / EV; ++b) {
            TLoc v[EV];
            if constexpr (I16) {
                const int4 t = reinterpret_cast<const int4*>(run)[b];
                v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
            } else {
                const double2 t = reinterpret_cast<const double2*>(run)[b];
                v[0] = t.x, v[1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < EV; ++i) {
                carry[(EV * b + i) % C] += v[i];
                v[i] = carry[(EV * b + i) % C];
            }
            if constexpr (I16) reinterpret_cast<int4*>(run)[b] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
            else reinterpret_cast<double2*>(run)[b] = make_double2((double)v[0], (double)v[1]);
        }
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
#pragma unroll
    for (int c = 0; c < C; ++c) {
        TLoc a = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW; ++w2)
            if (w2 < warp) a += s_warp[w2 * C + c];
        toff[tid * C + c] = a + (inc[c] - carry[c]);
    }

    // ---- 3. look-back (warp 0)
    if (s == 0 && warp == 0) scan_look_back<TAcc, C>(desc, tile, lane, s_excl, s_tot);
    __syncthreads();

    // ---- 4. output: pair q = tid + 256 * it, it < R / 2: elements e = 2 q, e + 1 of run e / R
    {
        const Pair* lp = reinterpret_cast<const Pair*>(loc + 2 * tid + PAD * ((2 * tid) >> 5));   // pad(2 q); + it * pad(512)
        constexpr int kPairStep = (512 + PAD * 16) / 2;                                         // in Pair units
        longlong2* op = reinterpret_cast<longlong2*>(cout) + tid;                               // + it * 256 (16-byte units)
        const int c0 = (2 * tid) % C;                                                           // 2 q mod C does not depend on it
        const TLoc* tp = toff + ((2 * tid) / R) * C + c0;                                       // + it * (512 / R) * C
        TAcc b0, b1;
        if constexpr (C == 1) {
            b0 = b1 = s_excl[0];
        } else {
            b0 = s_excl[c0], b1 = s_excl[c0 + 1];
        }
#pragma unroll
        for (int it = 0; it < R / 2; ++it) {
            const Pair l = lp[it * kPairStep];
            TLoc t0, t1;
            if constexpr (C == 1) {
                t0 = t1 = tp[it * (512 / R) * C];
            } else {
                t0 = tp[it * (512 / R) * C], t1 = tp[it * (512 / R) * C + 1];
            }
            const TAcc r0 = b0 + (TAcc)(t0 + (TLoc)l.x);
            const TAcc r1 = b1 + (TAcc)(t1 + (TLoc)l.y);
            longlong2 r;
            r.x = acc_bits<TAcc>(r0);
            r.y = acc_bits<TAcc>(r1);
            op[it * NT] = r;
        }
    }
    if (s + 1 < NCH) __syncthreads();   // the next chunk reuses the shared arrays
    }   // chunks of this CTA
}

template <typename TIn, int C, int R>
constexpr uint32_t scan_fast_smem_bytes()
{
    typedef typename ScanFast<TIn>::TLoc TLoc;
    return (uint32_t)((kScanThreads * R + kScanThreads * R / 32 * ScanFast<TIn>::kPad) * sizeof(TLoc) +
                      kScanThreads * C * sizeof(TLoc) + 2 * (kScanThreads / 32) * C * sizeof(TLoc) + 64);
}

// bytes of dynamic shared memory scan_lookback_kernel needs
template <typename TLoc, int C, int CB>
constexpr uint32_t scan_smem_bytes()
{
    constexpr uint32_t E = (uint32_t)scan_chunk_elems<TLoc, C, CB>();
    return (E + E / 32) * sizeof(TLoc) + kScanThreads * C * sizeof(TLoc) + 2 * (kScanThreads / 32) * C * sizeof(TLoc) + 64;
}

// ----------------------------------------------------------------------------------
// Windows beyond every shared-memory ring, up to 8 interleaved channels: what the reference's scan binaries do
// (averager_kernel over the prefix, basics/hillis_steele_averager.cu:87-100) on top of the single-pass prefix above,
//     y[f, c] = (P[f, c] - P[f - k, c]) / k,
// in exact int64 (int16 input, C truncating division) or fp64 (float32 input; a prefix over 2^28 samples keeps 25
// bits of headroom over the window sum).  The cost no longer depends on k.  Frames in front of the shard come from
// the prefix HP of its left context (hf frames), or are zeros.  One thread per frame, C <= 8 values each.
// ----------------------------------------------------------------------------------
template <typename T, typename TAcc>
__global__ void __launch_bounds__(256)
    prefix_diff_kernel(const TAcc* __restrict__ P, const TAcc* __restrict__ HP, T* __restrict__ y, uint64_t frames,
                       uint32_t C, uint64_t k, uint64_t hf)
{
    const double inv = 1.0 / (double)k;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride) {
        for (uint32_t c = 0; c < C; ++c) {
            const uint64_t i = f * C + c;
            TAcc w = P[i];
            if (f >= k) {
                w -= P[i - k * C];
            } else if (HP != nullptr && hf > 0) {
                const uint64_t want = k - 1 - f;                 // frames of the window that lie in front of the shard
                TAcc h = HP[(hf - 1) * C + c];                   // the whole context ...
                if (want < hf) h -= HP[(hf - want - 1) * C + c]; // ... or its last `want` frames
                if (want > 0) w += h;
            }
            y[i] = GenericAcc<T>::finish(w, (uint32_t)k, inv);
        }
    }
}

}  // namespace mavg
