// Single-pass stream compaction (decoupled look-back on the running count), shared by find / where /
// group-boundaries / unique.  Tile = 16 warps x 16 rows x 32 lanes = 8192 elements (64 KB of 8-byte input):
// the chained look-back advances about (32-tile window / state round trip) tiles per unit time, so a tile must
// carry enough bytes for that rate to exceed HBM bandwidth (4096-element tiles measured 16 % of roofline).
// Each warp ballots its rows so every lane knows the whole warp's masks without shuffles.
#pragma once
#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

constexpr int SEL_THREADS = 512;
constexpr int SEL_ROWS = 16;
constexpr int SEL_TILE = SEL_THREADS * SEL_ROWS;

inline size_t select_ws_bytes(i64 n) {
    const size_t tiles = (size_t)((n + SEL_TILE - 1) / SEL_TILE);
    return 64 + align_up(tiles * 8, 64);
}

// ws: [0,4) ticket | [64, ...) u64 state per tile — all zero on entry.
template <typename Pred>
__global__ void __launch_bounds__(SEL_THREADS) select_kernel(Pred pred, i64 n, i64* __restrict__ pos_out,
                                                              i64* __restrict__ count_out, void* ws) {
    __shared__ u32 s_tile;
    __shared__ u32 s_wcnt[SEL_THREADS / 32];
    __shared__ u64 s_base;
    u32* ticket = (u32*)ws;
    u64* states = (u64*)((char*)ws + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const i64 tile = s_tile;
    const i64 ntiles = (n + SEL_TILE - 1) / SEL_TILE;
    const i64 wbase = tile * SEL_TILE + (i64)warp * (SEL_ROWS * 32);

    // all of a thread's loads are issued before the first test: a fused load+compare per row lets the compiler
    // recycle one destination register, which serialises the loads (measured: 16 % of HBM roofline)
    typename Pred::V val[SEL_ROWS];
#pragma unroll
    for (int r = 0; r < SEL_ROWS; r++) {
        const i64 i = wbase + r * 32 + lane;
        val[r] = pred.load(i < n ? i : (n - 1));
    }
    bool hit[SEL_ROWS];
#pragma unroll
    for (int r = 0; r < SEL_ROWS; r++) {
        const i64 i = wbase + r * 32 + lane;
        hit[r] = (i < n) && pred.test(i, val[r]);
    }
    u32 masks[SEL_ROWS];
    u32 wtotal = 0;
#pragma unroll
    for (int r = 0; r < SEL_ROWS; r++) {
        masks[r] = __ballot_sync(0xffffffffu, hit[r]);
        wtotal += __popc(masks[r]);
    }
    if (lane == 0) s_wcnt[warp] = wtotal;
    __syncthreads();
    u32 woff = 0, ttotal = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; w++) {
        const u32 c = s_wcnt[w];
        if (w < warp) woff += c;
        ttotal += c;
    }
    if (warp == 0) {
        if (lane == 0)
            st_relaxed_u64(&states[tile], lb_pack(tile == 0 ? KGB_LB_PREFIX : KGB_LB_AGG, 1, ttotal));
        u64 excl = 0;
        if (tile > 0) {
            excl = lb_warp_lookback(states, tile, 1);
            if (lane == 0) st_relaxed_u64(&states[tile], lb_pack(KGB_LB_PREFIX, 1, excl + ttotal));
        }
        if (lane == 0) {
            s_base = excl;
            if (tile == ntiles - 1) *count_out = (i64)(excl + ttotal);
        }
    }
    __syncthreads();
    u64 o = s_base + woff;
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SEL_ROWS; r++) {
        if (hit[r]) pos_out[o + __popc(masks[r] & lt)] = wbase + r * 32 + lane;
        o += __popc(masks[r]);
    }
}

template <typename T> struct NonzeroPred {
    typedef T V;
    const T* a;
    __device__ __forceinline__ V load(i64 i) const { return __ldcs(&a[i]); }
    __device__ __forceinline__ bool test(i64, V v) const { return v != T(0); }
};
template <typename T> struct EqualPred {
    typedef T V;
    const T* a;
    T needle;
    __device__ __forceinline__ V load(i64 i) const { return __ldcs(&a[i]); }
    __device__ __forceinline__ bool test(i64, V v) const { return v == needle; }
};
// run heads of a sorted, order-encoded key array
struct BoundaryPred {
    typedef ulonglong2 V;
    const u64* k;
    __device__ __forceinline__ V load(i64 i) const { return make_ulonglong2(k[i], k[i > 0 ? i - 1 : 0]); }
    __device__ __forceinline__ bool test(i64 i, V v) const { return i == 0 || v.x != v.y; }
};

template <typename Pred>
inline int launch_select(Pred pred, i64 n, i64* pos_out, i64* count_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    KGB_REQUIRE(ws && ws_bytes >= select_ws_bytes(n), "select: workspace too small");
    const size_t tiles = (size_t)((n + SEL_TILE - 1) / SEL_TILE);
    KGB_CUDA_CHECK(cudaMemsetAsync(ws, 0, select_ws_bytes(n), st));
    if (tiles == 0) {
        KGB_CUDA_CHECK(cudaMemsetAsync(count_out, 0, 8, st));
        return KGB_OK;
    }
    select_kernel<Pred><<<(unsigned)tiles, SEL_THREADS, 0, st>>>(pred, n, pos_out, count_out, ws);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

}  // namespace kgb
