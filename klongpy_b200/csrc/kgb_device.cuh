// Device-side helpers shared by the AOT kernels (sort / select / group / groupagg).
#pragma once
#include <cstdint>

#include "../../include/kgb200.h"

namespace kgb {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef long long i64;

__device__ __forceinline__ u64 ld_acquire_u64(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(u64* p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back tile state packed in one 64-bit word (single-copy atomic):
//   [63:62] flag (0 empty, 1 aggregate, 2 inclusive prefix)  [61:56] epoch  [55:0] count
// The epoch lets several passes reuse one zero-initialised state array without re-zeroing: a word whose
// epoch differs from the reader's is treated as empty.
#define KGB_LB_AGG 1ull
#define KGB_LB_PREFIX 2ull
__device__ __forceinline__ u64 lb_pack(u64 flag, u32 epoch, u64 count) {
    return (flag << 62) | ((u64)(epoch & 63u) << 56) | (count & 0x00ffffffffffffffull);
}
__device__ __forceinline__ u64 lb_flag(u64 w, u32 epoch) {
    return (((w >> 56) & 63u) == (epoch & 63u)) ? (w >> 62) : 0ull;
}
__device__ __forceinline__ u64 lb_count(u64 w) { return w & 0x00ffffffffffffffull; }

// Exclusive prefix of `tile` over all previous tiles, one state word per tile.  Called by ALL 32 lanes of
// one warp; returns the prefix in every lane.  The caller must already have published its own aggregate.
__device__ __forceinline__ u64 lb_warp_lookback(const u64* states, i64 tile, u32 epoch) {
    const int lane = threadIdx.x & 31;
    u64 excl = 0;
    i64 look = tile - 1;
    while (look >= 0) {
        const i64 idx = look - lane;
        u64 w = lb_pack(KGB_LB_PREFIX, epoch, 0);  // virtual tiles before 0
        if (idx >= 0) w = ld_relaxed_u64(&states[idx]);  // the count travels inside the polled word: no acquire
        // all 32 lanes vote (lanes past tile 0 hold the virtual prefix and never spin)
        u32 spins = 0;
        while (__any_sync(0xffffffffu, lb_flag(w, epoch) == 0)) {
            if (lb_flag(w, epoch) == 0) w = ld_relaxed_u64(&states[idx]);
            if (++spins > (1u << 26)) __trap();  // a predecessor never published: fail, do not hang
        }
        const u32 pmask = __ballot_sync(0xffffffffu, lb_flag(w, epoch) == KGB_LB_PREFIX);
        const int first = __ffs(pmask) - 1;
        u64 c = (pmask == 0 || lane <= first) ? lb_count(w) : 0ull;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        excl += c;
        if (pmask != 0) break;
        look -= 32;
    }
    return excl;
}

// order-preserving key transforms (ascending unsigned order == numpy sort order, NaN last, -0.0 == +0.0)
__device__ __forceinline__ u64 key_from_i64(i64 x) { return (u64)x ^ 0x8000000000000000ull; }
__device__ __forceinline__ i64 key_to_i64(u64 k) { return (i64)(k ^ 0x8000000000000000ull); }
__device__ __forceinline__ u64 key_from_f64(double x) {
    if (x != x) return 0xffffffffffffffffull;
    if (x == 0.0) x = 0.0;  // -0.0 -> +0.0
    u64 b = (u64)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_f64(u64 k) {
    if (k == 0xffffffffffffffffull) return __longlong_as_double(0x7ff8000000000000ll);
    u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((i64)b);
}
__device__ __forceinline__ u64 key_from_i32(int x) { return (u64)((u32)x ^ 0x80000000u); }
__device__ __forceinline__ int key_to_i32(u64 k) { return (int)((u32)k ^ 0x80000000u); }
__device__ __forceinline__ u64 key_from_f32(float x) {
    if (x != x) return 0xffffffffull;
    if (x == 0.0f) x = 0.0f;
    u32 b = (u32)__float_as_int(x);
    return (u64)((b >> 31) ? ~b : (b | 0x80000000u));
}
__device__ __forceinline__ float key_to_f32(u64 k) {
    if (k == 0xffffffffull) return __int_as_float(0x7fc00000);
    u32 b = (u32)k;
    b = (b >> 31) ? (b & 0x7fffffffu) : ~b;
    return __int_as_float((int)b);
}

// internal dtype codes: float keys ordered by their raw bits (-0.0 < +0.0 stay distinct; NaNs collapse).
// `?a` needs this: the reference keys its set on str(x) (monads.py:284-294), so -0.0 and 0.0 differ.
#define KGB_F64_RAW 16
#define KGB_F32_RAW 18

template <int DT> struct KeyCodec;
template <> struct KeyCodec<KGB_F64_RAW> {
    typedef double T;
    static __device__ __forceinline__ u64 enc(T x) {
        if (x != x) return 0xffffffffffffffffull;
        u64 b = (u64)__double_as_longlong(x);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_f64(k); }
};
template <> struct KeyCodec<KGB_F32_RAW> {
    typedef float T;
    static __device__ __forceinline__ u64 enc(T x) {
        if (x != x) return 0xffffffffull;
        u32 b = (u32)__float_as_int(x);
        return (u64)((b >> 31) ? ~b : (b | 0x80000000u));
    }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_f32(k); }
};
template <> struct KeyCodec<KGB_F64> {
    typedef double T;
    static __device__ __forceinline__ u64 enc(T x) { return key_from_f64(x); }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_f64(k); }
};
template <> struct KeyCodec<KGB_I64> {
    typedef i64 T;
    static __device__ __forceinline__ u64 enc(T x) { return key_from_i64(x); }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_i64(k); }
};
template <> struct KeyCodec<KGB_F32> {
    typedef float T;
    static __device__ __forceinline__ u64 enc(T x) { return key_from_f32(x); }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_f32(k); }
};
template <> struct KeyCodec<KGB_I32> {
    typedef int T;
    static __device__ __forceinline__ u64 enc(T x) { return key_from_i32(x); }
    static __device__ __forceinline__ T dec(u64 k) { return key_to_i32(k); }
};

}  // namespace kgb
