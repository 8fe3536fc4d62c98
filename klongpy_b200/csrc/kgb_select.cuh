// Stream compaction for find / where / group boundaries / unique, in three kernels with a bitmask in between:
//
//   select_mask_kernel   : pure streaming pass over the input (no inter-CTA dependency at all).  A warp takes 32
//                          rows of 32 elements, every lane issues its 32 loads before the first test, each row is
//                          balloted, and lane r keeps row r's mask — so a warp leaves ONE coalesced 128-byte store
//                          per 1024 elements, plus (only when it found something) one RED of its hit count into
//                          the counter of its 65536-element tile.  Traffic: the input once + n/8 bytes.
//   select_offsets_kernel: one CTA turns the tile counters into exclusive offsets in place and writes the total.
//   select_emit_kernel   : every tile with hits expands its 2048 mask words into ascending positions (per-thread
//                          popcounts, block scan, tile offset from the table); empty tiles return without reading.
//
// History: a single-pass kernel with the look-back in the streaming loop measured 0.16 -> 0.33 of
// the HBM roofline and no better when made persistent with prefetch (0.27): with ~150 tiles in flight in lock-step
// every look-back walks 2-5 dependent L2 round trips while the CTA's other warps idle (60 % of the samples at that
// barrier, ncu r01h).  Moving the dependency chain onto the 64x smaller mask stream (mask kernel + a compaction
// kernel with decoupled look-back over the mask words) gave 0.59 at 1e8 / 0.79 at 1e9 — the chain was still there,
// ~20 ns per 8 KB tile, 0.3 ms at 1e9.  Counting hits per tile in the mask pass removes the chain altogether.
#pragma once
#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

constexpr int SEL_MASK_THREADS = 256;             // 8 warps; a warp step covers 32 x 32 = 1024 elements
constexpr int SEL_WARP_ELEMS = 1024;
constexpr int SEL_THREADS = 512;                  // compaction kernel
constexpr int SEL_WPT = 4;                        // mask words per thread
constexpr int SEL_TILE_WORDS = SEL_THREADS * SEL_WPT;
constexpr i64 SEL_TILE = (i64)SEL_TILE_WORDS * 32;  // elements per compaction tile

inline size_t select_mask_words(i64 n) { return (size_t)((n + SEL_WARP_ELEMS - 1) / SEL_WARP_ELEMS) * 32; }
inline size_t select_tiles(i64 n) { return (select_mask_words(n) + SEL_TILE_WORDS - 1) / SEL_TILE_WORDS; }
constexpr int SEL_TILE_STEPS = SEL_TILE_WORDS / 32;  // mask-kernel warp steps per tile
// ws: [0,64) unused | [64, ...) u64 hit counter per tile (+1 slot for the total; zero on entry) | mask words
inline size_t select_state_bytes(i64 n) { return 64 + align_up((select_tiles(n) + 1) * 8, 64); }
inline size_t select_ws_bytes(i64 n) { return select_state_bytes(n) + align_up(select_mask_words(n) * 4, 64); }

template <typename Pred>
__global__ void __launch_bounds__(SEL_MASK_THREADS) select_mask_kernel(Pred pred, i64 n, u32* __restrict__ masks,
                                                                        unsigned long long* __restrict__ tilecnt) {
    const int lane = threadIdx.x & 31;
    const i64 nsteps = (n + SEL_WARP_ELEMS - 1) / SEL_WARP_ELEMS;
    const i64 wstride = (i64)gridDim.x * (SEL_MASK_THREADS / 32);
    for (i64 step = (i64)blockIdx.x * (SEL_MASK_THREADS / 32) + (threadIdx.x >> 5); step < nsteps; step += wstride) {
        const i64 wbase = step * SEL_WARP_ELEMS;
        // all of a thread's loads are issued before the first test: a fused load+compare per row lets the compiler
        // recycle one destination register, which serialises the loads (measured: 16 % of HBM roofline)
        typename Pred::V val[32];
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const i64 i = wbase + r * 32 + lane;
            val[r] = pred.load(i < n ? i : (n - 1));
        }
        u32 mine = 0;
        if (Pred::NEEDS_PREV) {
            // element i-1 comes from the neighbouring lane (or the previous row's lane 31); only lane 0 of row 0
            // loads it — one extra load per 1024 elements instead of a second value per element
            typename Pred::V prev0 = val[0];
            if (lane == 0 && wbase > 0) prev0 = pred.load(wbase - 1);
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const i64 i = wbase + r * 32 + lane;
                typename Pred::V up = __shfl_up_sync(0xffffffffu, val[r], 1);
                typename Pred::V last = __shfl_sync(0xffffffffu, val[r > 0 ? r - 1 : 0], 31);
                const typename Pred::V prev = lane > 0 ? up : (r > 0 ? last : prev0);
                const u32 m = __ballot_sync(0xffffffffu, (i < n) && pred.test2(i, val[r], prev));
                if (lane == r) mine = m;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const i64 i = wbase + r * 32 + lane;
                const u32 m = __ballot_sync(0xffffffffu, (i < n) && pred.test(i, val[r]));
                if (lane == r) mine = m;
            }
        }
        masks[step * 32 + lane] = mine;
        const u32 hits = __reduce_add_sync(0xffffffffu, (u32)__popc(mine));
        if (lane == 0 && hits) atomicAdd(&tilecnt[step / SEL_TILE_STEPS], (unsigned long long)hits);
    }
}

// counters -> exclusive offsets, in place; cnt[ntiles] and *count_out receive the total
static __global__ void __launch_bounds__(1024) select_offsets_kernel(unsigned long long* __restrict__ cnt, i64 ntiles,
                                                              i64* __restrict__ count_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (i64 base = 0; base < ntiles; base += 1024) {
        const i64 i = base + tid;
        const unsigned long long c = i < ntiles ? cnt[i] : 0ull;
        unsigned long long inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_w[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += v;
            }
            s_w[lane] = wi - w;  // exclusive over warps
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (i < ntiles) cnt[i] = carry + s_w[warp] + (inc - c);
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[warp] + inc;
        __syncthreads();
    }
    if (tid == 0) {
        cnt[ntiles] = s_carry;
        *count_out = (i64)s_carry;
        __threadfence_system();  // count_out may be mapped host memory that the caller polls while the emit kernel runs
    }
}

static __global__ void __launch_bounds__(SEL_THREADS) select_emit_kernel(const u32* __restrict__ masks, i64 nwords,
                                                                   i64* __restrict__ pos_out,
                                                                   const unsigned long long* __restrict__ offs) {
    __shared__ u32 s_wcnt[SEL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 tile = blockIdx.x;
    const unsigned long long base = offs[tile];
    if (offs[tile + 1] == base) return;  // no hit in this tile (uniform across the CTA): nothing to read or write
    const i64 w0 = tile * SEL_TILE_WORDS + (i64)tid * SEL_WPT;
    u32 m[SEL_WPT];
    if (w0 + SEL_WPT <= nwords) {
        const uint4 q = *reinterpret_cast<const uint4*>(masks + w0);  // nwords is a multiple of 32, w0 of 4
        m[0] = q.x, m[1] = q.y, m[2] = q.z, m[3] = q.w;
    } else {
#pragma unroll
        for (int j = 0; j < SEL_WPT; j++) m[j] = (w0 + j < nwords) ? masks[w0 + j] : 0u;
    }
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < SEL_WPT; j++) cnt += __popc(m[j]);
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wcnt[warp] = inc;
    __syncthreads();
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; w++) {
        const u32 c = s_wcnt[w];
        if (w < warp) woff += c;
    }
    u64 o = base + woff + (inc - cnt);
#pragma unroll
    for (int j = 0; j < SEL_WPT; j++) {
        u32 mm = m[j];
        const i64 ebase = (w0 + j) * 32;
        while (mm) {
            const int b = __ffs(mm) - 1;
            pos_out[o++] = ebase + b;
            mm &= mm - 1;
        }
    }
}

template <typename T> struct NonzeroPred {
    typedef T V;
    static constexpr bool NEEDS_PREV = false;
    const T* a;
    __device__ __forceinline__ V load(i64 i) const { return __ldcs(&a[i]); }
    __device__ __forceinline__ bool test(i64, V v) const { return v != T(0); }
    __device__ __forceinline__ bool test2(i64, V, V) const { return false; }
};
template <typename T> struct EqualPred {
    typedef T V;
    static constexpr bool NEEDS_PREV = false;
    const T* a;
    T needle;
    __device__ __forceinline__ V load(i64 i) const { return __ldcs(&a[i]); }
    __device__ __forceinline__ bool test(i64, V v) const { return v == needle; }
    __device__ __forceinline__ bool test2(i64, V, V) const { return false; }
};
// run heads of a sorted, order-encoded key array: k[i] != k[i-1]
struct BoundaryPred {
    typedef u64 V;
    static constexpr bool NEEDS_PREV = true;
    const u64* k;
    __device__ __forceinline__ V load(i64 i) const { return k[i]; }
    __device__ __forceinline__ bool test(i64, V) const { return false; }
    __device__ __forceinline__ bool test2(i64 i, V v, V prev) const { return i == 0 || v != prev; }
};

template <typename Pred>
inline int launch_select(Pred pred, i64 n, i64* pos_out, i64* count_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    KGB_REQUIRE(ws && ws_bytes >= select_ws_bytes(n), "select: workspace too small");
    if (n == 0) {
        KGB_CUDA_CHECK(cudaMemsetAsync(count_out, 0, 8, st));
        return KGB_OK;
    }
    KGB_CUDA_CHECK(cudaMemsetAsync(ws, 0, select_state_bytes(n), st));
    u32* masks = (u32*)((char*)ws + select_state_bytes(n));
    const size_t nwords = select_mask_words(n);
    const long long steps = (long long)(nwords / 32);
    long long mb = (steps + (SEL_MASK_THREADS / 32) - 1) / (SEL_MASK_THREADS / 32);
    const long long cap = (long long)sm_count() * 8;
    if (mb > cap) mb = cap;
    unsigned long long* tilecnt = (unsigned long long*)((char*)ws + 64);
    const i64 ntiles = (i64)select_tiles(n);
    select_mask_kernel<Pred><<<(unsigned)mb, SEL_MASK_THREADS, 0, st>>>(pred, n, masks, tilecnt);
    select_offsets_kernel<<<1, 1024, 0, st>>>(tilecnt, ntiles, count_out);
    select_emit_kernel<<<(unsigned)ntiles, SEL_THREADS, 0, st>>>(masks, (i64)nwords, pos_out, tilecnt);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

}  // namespace kgb
