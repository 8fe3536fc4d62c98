// kgb200 JIT skeleton — the hand-written part of every fused expression kernel (sm_100a).
//
// expr_codegen.cpp prepends a block of macros describing ONE expression tree:
//   KGB_MODE            0 map / 1 reduce / 2 scan / 3 scan+epilogue+reduce
//   KGB_VEC             elements per 128-bit (or narrower) vector group
//   KGB_OUT_T, KGB_ACC_T  C types of the stored result / the fold accumulator
//   KGB_SCALARS         loads of by-value and 0-d device scalars into locals s<i>
//   KGB_DECL_REGS(N)    register arrays a<i>[N] for each ARRAY argument
//   KGB_LOAD_VEC(V,off,idx)   a<i>[off..off+V) <- V elements at element index idx (vector load)
//   KGB_EVAL(j)         the expression value for register slot j
//   KGB_RED(a,b), KGB_RED_IDENT           fold operator + identity (modes 1,2,3)
//   KGB_EVAL2(j,sv), KGB_RED2(a,b), KGB_RED2_IDENT, KGB_ACC2_T   (mode 3 epilogue)
//   KGB_HAS_CARRY, KGB_CARRY   (mode 2, optional) scalar seed folded in front of element 0
// and this file supplies the memory movement: 128-bit streaming loads/stores, grid-stride unrolling
// for bytes-in-flight, warp-shuffle + block folds with a fixed-order last-block combine
// (deterministic), and the single-pass decoupled-lookback scan.
// Compiled by NVRTC with --fmad=false so a*b+c rounds twice like NumPy.

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

#define KGB_MAX_ARGS 12
#define KGB_BLOCK 256
#ifndef KGB_UNROLL
#define KGB_UNROLL 4
#endif

struct KgbParams {
    const void* ptr[KGB_MAX_ARGS];
    double sf[KGB_MAX_ARGS];
    i64 si[KGB_MAX_ARGS];
    void* out;
    void* ws;
    i64 n;
};

// ------------------------------------------------------------------ scalar helpers
__device__ __forceinline__ double kgb_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ float kgb_inff() { return __int_as_float(0x7f800000); }

template <typename T> struct kgb_lim;
template <> struct kgb_lim<double> {
    static __device__ __forceinline__ double lo() { return -kgb_inf(); }
    static __device__ __forceinline__ double hi() { return kgb_inf(); }
};
template <> struct kgb_lim<float> {
    static __device__ __forceinline__ float lo() { return -kgb_inff(); }
    static __device__ __forceinline__ float hi() { return kgb_inff(); }
};
template <> struct kgb_lim<i64> {
    static __device__ __forceinline__ i64 lo() { return (i64)0x8000000000000000LL; }
    static __device__ __forceinline__ i64 hi() { return 0x7fffffffffffffffLL; }
};
template <> struct kgb_lim<int> {
    static __device__ __forceinline__ int lo() { return (int)0x80000000; }
    static __device__ __forceinline__ int hi() { return 0x7fffffff; }
};
template <> struct kgb_lim<u8> {
    static __device__ __forceinline__ u8 lo() { return 0; }
    static __device__ __forceinline__ u8 hi() { return 1; }
};

// np.maximum / np.minimum: NaN-propagating for floats
// (a NaN -> a; b NaN -> b because a > NaN is false: two compares instead of three)
// NaN-propagating float64 max / min (np.maximum / np.minimum): `(a != a || a > b) ? a : b`, spelled in PTX so that it
// is two DSETP (the second OR-ed with the first) and one 64-bit select.  The C++ form compiles to two DSETP + four
// FSEL and, inside the unrolled scan loops, to divergent branches (BSSY/BRA/BSYNC): 56 instructions per element in
// `|/(|\x)-x` (ncu r01v), 26 of them selects and compares, 6.5 branch bookkeeping.
__device__ __forceinline__ double kgb_max(double a, double b) {
    double r;
    asm("{ .reg .pred p, q; setp.nan.f64 p, %1, %1; setp.gt.or.f64 q, %1, %2, p; selp.f64 %0, %1, %2, q; }"
        : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ double kgb_min(double a, double b) {
    double r;
    asm("{ .reg .pred p, q; setp.nan.f64 p, %1, %1; setp.lt.or.f64 q, %1, %2, p; selp.f64 %0, %1, %2, q; }"
        : "=d"(r) : "d"(a), "d"(b));
    return r;
}
__device__ __forceinline__ float kgb_max(float a, float b) { return (a != a || a > b) ? a : b; }
__device__ __forceinline__ float kgb_min(float a, float b) { return (a != a || a < b) ? a : b; }
__device__ __forceinline__ i64 kgb_max(i64 a, i64 b) { return a > b ? a : b; }
__device__ __forceinline__ i64 kgb_min(i64 a, i64 b) { return a < b ? a : b; }
__device__ __forceinline__ int kgb_max(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int kgb_min(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ u8 kgb_max(u8 a, u8 b) { return a > b ? a : b; }
__device__ __forceinline__ u8 kgb_min(u8 a, u8 b) { return a < b ? a : b; }

// numpy ** : exponent 2 is an exact square (NumPy's scalar fast path; also what a correctly rounded
// pow returns), 0.5 is sqrt; everything else goes to pow().
__device__ __forceinline__ double kgb_pow(double a, double b) {
    if (b == 2.0) return a * a;
    if (b == 0.5) return sqrt(a);
    if (b == 1.0) return a;
    if (b == -1.0) return 1.0 / a;
    return pow(a, b);
}
__device__ __forceinline__ float kgb_pow(float a, float b) {
    if (b == 2.0f) return a * a;
    if (b == 0.5f) return sqrtf(a);
    if (b == 1.0f) return a;
    return powf(a, b);
}
// integer power by squaring, wrapping like numpy int64.  Negative exponents (a ValueError in numpy)
// follow the mathematical truncation: 1 -> 1, -1 -> +-1, otherwise 0.
__device__ __forceinline__ i64 kgb_pow(i64 a, i64 b) {
    if (b < 0) {
        if (a == 1) return 1;
        if (a == -1) return (b & 1) ? -1 : 1;
        return 0;
    }
    u64 r = 1, x = (u64)a;
    while (b) {
        if (b & 1) r *= x;
        x *= x;
        b >>= 1;
    }
    return (i64)r;
}
__device__ __forceinline__ int kgb_pow(int a, int b) { return (int)kgb_pow((i64)a, (i64)b); }

// np.fmod: C truncated remainder; integer x % 0 -> 0 (numpy warns and returns 0)
__device__ __forceinline__ double kgb_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ float kgb_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ i64 kgb_fmod(i64 a, i64 b) { return b == 0 ? 0 : (b == -1 ? 0 : a % b); }
__device__ __forceinline__ int kgb_fmod(int a, int b) { return b == 0 ? 0 : (b == -1 ? 0 : a % b); }
// np.floor_divide / np.remainder.  Integers: exact, x // 0 == x % 0 == 0, INT_MIN // -1 wraps (numpy's loops).
// Floats: numpy's npy_divmod (fmod, then the Python sign convention, then the quotient snapped to an integer).
template <typename I> __device__ __forceinline__ I kgb_ifloordiv(I a, I b) {
    if (b == 0) return 0;
    if (b == -1) return (I)(0 - (unsigned long long)a);
    I q = a / b;
    if (((a > 0) != (b > 0)) && q * b != a) q--;
    return q;
}
template <typename I> __device__ __forceinline__ I kgb_imod(I a, I b) {
    if (b == 0 || b == -1) return 0;
    I r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) r += b;
    return r;
}
template <typename F> __device__ __forceinline__ F kgb_fdivmod(F a, F b, F* modulus) {
    F mod = (sizeof(F) == 8) ? (F)fmod((double)a, (double)b) : (F)fmodf((float)a, (float)b);
    if (b == (F)0) {
        *modulus = mod;
        return a / b;
    }
    F div = (a - mod) / b;
    if (mod != (F)0) {
        if ((b < (F)0) != (mod < (F)0)) {
            mod += b;
            div -= (F)1;
        }
    } else {
        mod = (sizeof(F) == 8) ? (F)copysign(0.0, (double)b) : (F)copysignf(0.0f, (float)b);
    }
    F fd;
    if (div != (F)0) {
        fd = (sizeof(F) == 8) ? (F)floor((double)div) : (F)floorf((float)div);
        if (div - fd > (F)0.5) fd += (F)1;
    } else {
        fd = (sizeof(F) == 8) ? (F)copysign(0.0, (double)(a / b)) : (F)copysignf(0.0f, (float)(a / b));
    }
    *modulus = mod;
    return fd;
}
__device__ __forceinline__ i64 kgb_floordiv(i64 a, i64 b) { return kgb_ifloordiv<i64>(a, b); }
__device__ __forceinline__ int kgb_floordiv(int a, int b) { return kgb_ifloordiv<int>(a, b); }
__device__ __forceinline__ double kgb_floordiv(double a, double b) { double m; return kgb_fdivmod<double>(a, b, &m); }
__device__ __forceinline__ float kgb_floordiv(float a, float b) { float m; return kgb_fdivmod<float>(a, b, &m); }
__device__ __forceinline__ i64 kgb_mod(i64 a, i64 b) { return kgb_imod<i64>(a, b); }
__device__ __forceinline__ int kgb_mod(int a, int b) { return kgb_imod<int>(a, b); }
__device__ __forceinline__ double kgb_mod(double a, double b) { double m; kgb_fdivmod<double>(a, b, &m); return m; }
__device__ __forceinline__ float kgb_mod(float a, float b) { float m; kgb_fdivmod<float>(a, b, &m); return m; }

__device__ __forceinline__ double kgb_abs(double a) { return fabs(a); }
__device__ __forceinline__ float kgb_abs(float a) { return fabsf(a); }
__device__ __forceinline__ i64 kgb_abs(i64 a) { return a < 0 ? -a : a; }
__device__ __forceinline__ int kgb_abs(int a) { return a < 0 ? -a : a; }
__device__ __forceinline__ u8 kgb_abs(u8 a) { return a; }

__device__ __forceinline__ double kgb_sign(double a) { return (a != a) ? a : (double)((a > 0.0) - (a < 0.0)); }
__device__ __forceinline__ float kgb_sign(float a) { return (a != a) ? a : (float)((a > 0.0f) - (a < 0.0f)); }
__device__ __forceinline__ i64 kgb_sign(i64 a) { return (i64)((a > 0) - (a < 0)); }
__device__ __forceinline__ int kgb_sign(int a) { return (a > 0) - (a < 0); }

__device__ __forceinline__ bool kgb_isclose(double a, double b) {
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    return fabs(a - b) <= 1e-08 + 1e-05 * fabs(b);
}

// ------------------------------------------------------------------ vector load / store
template <int BYTES> struct kgb_raw;
template <> struct kgb_raw<1> { typedef unsigned char type; };
template <> struct kgb_raw<2> { typedef unsigned short type; };
template <> struct kgb_raw<4> { typedef unsigned int type; };
template <> struct kgb_raw<8> { typedef uint2 type; };
template <> struct kgb_raw<16> { typedef uint4 type; };

// streaming (evict-first) load of V contiguous elements into registers dst[0..V).  32-byte groups use the 256-bit
// global accesses of sm_100 (LDG.E.256 / STG.E.256): half the memory instructions per byte moved.
#ifndef KGB_LDHINT
#define KGB_LDHINT ".cs"   // streaming (evict-first) by default; tuning: "" = default policy, ".cg", ".lu"
#define KGB_STHINT ".cs"
#endif
template <typename T, int V> __device__ __forceinline__ void kgb_ldv(T* dst, const T* src) {
    if constexpr (sizeof(T) * V == 32) {
        union {
            u64 r[4];
            T t[V];
        } u;
        asm volatile("ld.global" KGB_LDHINT ".v4.u64 {%0, %1, %2, %3}, [%4];"
                     : "=l"(u.r[0]), "=l"(u.r[1]), "=l"(u.r[2]), "=l"(u.r[3]) : "l"(src));
#pragma unroll
        for (int j = 0; j < V; j++) dst[j] = u.t[j];
    } else {
        typedef typename kgb_raw<sizeof(T) * V>::type R;
        union {
            R r;
            T t[V];
        } u;
        u.r = __ldcs(reinterpret_cast<const R*>(src));
#pragma unroll
        for (int j = 0; j < V; j++) dst[j] = u.t[j];
    }
}
template <typename T, int V> __device__ __forceinline__ void kgb_stv(T* dst, const T* src) {
    if constexpr (sizeof(T) * V == 32) {
        union {
            u64 r[4];
            T t[V];
        } u;
#pragma unroll
        for (int j = 0; j < V; j++) u.t[j] = src[j];
        asm volatile("st.global" KGB_STHINT ".v4.u64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(u.r[0]), "l"(u.r[1]), "l"(u.r[2]), "l"(u.r[3])
                     : "memory");
    } else {
        typedef typename kgb_raw<sizeof(T) * V>::type R;
        union {
            R r;
            T t[V];
        } u;
#pragma unroll
        for (int j = 0; j < V; j++) u.t[j] = src[j];
        __stcs(reinterpret_cast<R*>(dst), u.r);
    }
}

// ------------------------------------------------------------------ warp / block folds
template <typename T> __device__ __forceinline__ T kgb_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ u8 kgb_shfl_xor<u8>(u8 v, int m) { return (u8)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <typename T> __device__ __forceinline__ T kgb_shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <> __device__ __forceinline__ u8 kgb_shfl_up<u8>(u8 v, int d) { return (u8)__shfl_up_sync(0xffffffffu, (int)v, d); }
template <typename T> __device__ __forceinline__ T kgb_shfl(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }
template <> __device__ __forceinline__ u8 kgb_shfl<u8>(u8 v, int l) { return (u8)__shfl_sync(0xffffffffu, (int)v, l); }

// =====================================================================================================
#if KGB_MODE == 0
// ------------------------------------------------------------------ MAP: out[i] = f(args[i])
// Grid-stride over vector groups, KGB_UNROLL independent 128-bit loads per array in flight per thread.
template <int V> __device__ __forceinline__ void kgb_map_body(const KgbParams& p) {
    KGB_SCALARS
    KGB_OUT_T* __restrict__ out = (KGB_OUT_T*)p.out;
    const i64 n = p.n;
    const i64 ngroups = n / V;
    const i64 stride = (i64)gridDim.x * KGB_BLOCK;
    const i64 tid = (i64)blockIdx.x * KGB_BLOCK + threadIdx.x;
    for (i64 g0 = tid; g0 < ngroups; g0 += stride * KGB_UNROLL) {
        KGB_DECL_REGS(V * KGB_UNROLL)
#pragma unroll
        for (int u = 0; u < KGB_UNROLL; u++) {
            const i64 g = g0 + u * stride;
            if (g < ngroups) { KGB_LOAD_VEC(V, u * V, g * V) }
        }
#pragma unroll
        for (int u = 0; u < KGB_UNROLL; u++) {
            const i64 g = g0 + u * stride;
            if (g < ngroups) {
                KGB_OUT_T o[V];
#pragma unroll
                for (int j = 0; j < V; j++) o[j] = (KGB_OUT_T)(KGB_EVAL(u * V + j));
                kgb_stv<KGB_OUT_T, V>(out + g * V, o);
            }
        }
    }
    // scalar tail (n % V elements)
    const i64 t = ngroups * V + tid;
    if (t < n) {
        KGB_DECL_REGS(1)
        KGB_LOAD_VEC(1, 0, t)
        out[t] = (KGB_OUT_T)(KGB_EVAL(0));
    }
}
extern "C" __global__ void __launch_bounds__(KGB_BLOCK) KGB_KERNEL_NAME_V(const __grid_constant__ KgbParams p) { kgb_map_body<KGB_VEC>(p); }
extern "C" __global__ void __launch_bounds__(KGB_BLOCK) KGB_KERNEL_NAME_S(const __grid_constant__ KgbParams p) { kgb_map_body<1>(p); }
#endif

// =====================================================================================================
#if KGB_MODE == 1
// ------------------------------------------------------------------ REDUCE: out[0] = fold f(args[i])
// Per-thread V*UNROLL independent accumulators -> fixed tree -> warp shuffle -> block smem -> one
// partial per block in ws -> the last block to finish folds the partials in index order.  The shape of
// the tree depends only on (n, grid), so results are bit-reproducible run to run.
// ws layout: [32,36) ticket (zero on entry, zero again on exit), [64, 64+grid*sizeof(ACC)) partials.
__device__ __forceinline__ KGB_ACC_T kgb_block_fold(KGB_ACC_T v, KGB_ACC_T* sh) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        KGB_ACC_T o = kgb_shfl_xor<KGB_ACC_T>(v, m);
        v = KGB_RED(v, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < KGB_BLOCK / 32) ? sh[lane] : (KGB_ACC_T)(KGB_RED_IDENT);
#pragma unroll
        for (int m = (KGB_BLOCK / 64); m >= 1; m >>= 1) {
            KGB_ACC_T o = kgb_shfl_xor<KGB_ACC_T>(v, m);
            v = KGB_RED(v, o);
        }
    }
    return v;  // valid in thread 0
}

#ifdef KGB_HAS_RED2
// second fold of a two-output reduce (KGB_EVALB / KGB_RED2): `(+/p*s)%+/s` reads p and s ONCE for both sums, and a
// sharded scan gets its shard total from the same pass as another reduction over the same vector
__device__ __forceinline__ KGB_ACC2_T kgb_block_fold2(KGB_ACC2_T v, KGB_ACC2_T* sh) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(v, m);
        v = KGB_RED2(v, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = (lane < KGB_BLOCK / 32) ? sh[lane] : (KGB_ACC2_T)(KGB_RED2_IDENT);
#pragma unroll
        for (int m = (KGB_BLOCK / 64); m >= 1; m >>= 1) {
            KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(v, m);
            v = KGB_RED2(v, o);
        }
    }
    return v;  // valid in thread 0
}
#endif

template <int V> __device__ __forceinline__ void kgb_reduce_body(const KgbParams& p) {
    KGB_SCALARS
    __shared__ KGB_ACC_T sh[KGB_BLOCK / 32];
    __shared__ bool is_last;
    const i64 n = p.n;
    const i64 ngroups = n / V;
    const i64 stride = (i64)gridDim.x * KGB_BLOCK;
    const i64 tid = (i64)blockIdx.x * KGB_BLOCK + threadIdx.x;
    KGB_ACC_T acc[V * KGB_UNROLL];
#pragma unroll
    for (int k = 0; k < V * KGB_UNROLL; k++) acc[k] = (KGB_ACC_T)(KGB_RED_IDENT);
#ifdef KGB_HAS_RED2
    __shared__ KGB_ACC2_T sh2[KGB_BLOCK / 32];
    KGB_ACC2_T acc2[V * KGB_UNROLL];
#pragma unroll
    for (int k = 0; k < V * KGB_UNROLL; k++) acc2[k] = (KGB_ACC2_T)(KGB_RED2_IDENT);
#endif
    for (i64 g0 = tid; g0 < ngroups; g0 += stride * KGB_UNROLL) {
        KGB_DECL_REGS(V * KGB_UNROLL)
#pragma unroll
        for (int u = 0; u < KGB_UNROLL; u++) {
            const i64 g = g0 + u * stride;
            if (g < ngroups) { KGB_LOAD_VEC(V, u * V, g * V) }
        }
#pragma unroll
        for (int u = 0; u < KGB_UNROLL; u++) {
            const i64 g = g0 + u * stride;
            if (g < ngroups) {
#pragma unroll
                for (int j = 0; j < V; j++) {
                    KGB_ACC_T e = (KGB_ACC_T)(KGB_EVAL(u * V + j));
                    acc[u * V + j] = KGB_RED(acc[u * V + j], e);
#ifdef KGB_HAS_RED2
                    KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVALB(u * V + j));
                    acc2[u * V + j] = KGB_RED2(acc2[u * V + j], e2);
#endif
                }
            }
        }
    }
    const i64 t = ngroups * V + tid;
    if (t < n) {
        KGB_DECL_REGS(1)
        KGB_LOAD_VEC(1, 0, t)
        KGB_ACC_T e = (KGB_ACC_T)(KGB_EVAL(0));
        acc[0] = KGB_RED(acc[0], e);
#ifdef KGB_HAS_RED2
        KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVALB(0));
        acc2[0] = KGB_RED2(acc2[0], e2);
#endif
    }
#pragma unroll
    for (int w = (V * KGB_UNROLL) / 2; w >= 1; w >>= 1) {
#pragma unroll
        for (int k = 0; k < w; k++) {
            acc[k] = KGB_RED(acc[k], acc[k + w]);
#ifdef KGB_HAS_RED2
            acc2[k] = KGB_RED2(acc2[k], acc2[k + w]);
#endif
        }
    }
    KGB_ACC_T v = kgb_block_fold(acc[0], sh);

    // the fold's ticket has its own word (ws + 32): the scan kernels keep their ticket at ws + 0 and leave it dirty, this
    // one is returned to zero by the last CTA — a workspace that was zero-filled once stays valid for every later fold
    // without a memset in front of the launch (kgb_expr_launch_ex, KGB_LAUNCH_WS_CLEAN)
    u32* ticket = (u32*)((char*)p.ws + 32);
    volatile KGB_ACC_T* partial = (volatile KGB_ACC_T*)((char*)p.ws + 64);
#ifdef KGB_HAS_RED2
    KGB_ACC2_T v2 = kgb_block_fold2(acc2[0], sh2);
    // second partial array sits behind the first one (the host sizes the workspace for both)
    volatile KGB_ACC2_T* partial2 = (volatile KGB_ACC2_T*)((char*)p.ws + 64 + (((size_t)gridDim.x * sizeof(KGB_ACC_T) + 63) / 64) * 64);
#endif
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = v;
#ifdef KGB_HAS_RED2
        partial2[blockIdx.x] = v2;
#endif
        __threadfence();
        const u32 tk = atomicAdd(ticket, 1u);
        is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        KGB_ACC_T a = (KGB_ACC_T)(KGB_RED_IDENT);
        for (u32 b = threadIdx.x; b < gridDim.x; b += KGB_BLOCK) {
            KGB_ACC_T o = partial[b];
            a = KGB_RED(a, o);
        }
        __syncthreads();
        a = kgb_block_fold(a, sh);
#ifdef KGB_HAS_RED2
        KGB_ACC2_T a2 = (KGB_ACC2_T)(KGB_RED2_IDENT);
        for (u32 b = threadIdx.x; b < gridDim.x; b += KGB_BLOCK) {
            KGB_ACC2_T o = partial2[b];
            a2 = KGB_RED2(a2, o);
        }
        __syncthreads();
        a2 = kgb_block_fold2(a2, sh2);
#endif
        if (threadIdx.x == 0) {
#ifdef KGB_HAS_EPI
            *(KGB_EPI_T*)p.out = (KGB_EPI_T)(KGB_EPI(a, a2));  // ONE result: the epilogue of the two finished folds
#else
            *(KGB_OUT_T*)p.out = (KGB_OUT_T)a;
#ifdef KGB_HAS_RED2
            *(KGB_ACC2_T*)((char*)p.out + 8) = a2;  // second result in the next 8-byte slot
#endif
#endif
            *ticket = 0u;
        }
    }
}
extern "C" __global__ void __launch_bounds__(KGB_BLOCK) KGB_KERNEL_NAME_V(const __grid_constant__ KgbParams p) { kgb_reduce_body<KGB_VEC>(p); }
extern "C" __global__ void __launch_bounds__(KGB_BLOCK) KGB_KERNEL_NAME_S(const __grid_constant__ KgbParams p) { kgb_reduce_body<1>(p); }
#endif

// =====================================================================================================
#if KGB_MODE == 2 || KGB_MODE == 3
// ------------------------------------------------------------------ SCAN: out[i] = fold_{j<=i} f(args[j])
// Single-pass chained scan with decoupled look-back.  One tile = KGB_SBLOCK threads x KGB_ITEMS elements
// (KGB_SBLOCK / KGB_ITEMS are chosen by the host per accumulator width; DESIGN.md "scan").
//   1. coalesced 128-bit striped loads, expression evaluated, values parked in padded smem
//   2. each thread re-reads its contiguous chunk (bank-conflict-free thanks to the 16 B pad per chunk),
//      scans it in registers; warp-shuffle + cross-warp scan of thread totals
//   3. tile aggregate published; warp 0 looks back over predecessor tiles (32 per round)
//   4. results go back through smem and leave as coalesced 128-bit streaming stores
// Tiles take their index from an atomic ticket so a tile's predecessors are always already running.
//
// Tile state = ONE 16-byte record {w0, w1} read and written with a single 128-bit access and no fence: each
// 64-bit half carries the status in its upper 32 bits and half of the value's bit pattern in its lower 32, so a
// reader accepts a record only when both halves show the same status — correct with nothing more than
// 64-bit single-copy atomicity.  The look-back throughput of the whole scan is (window / round-trip) tiles per
// unit time, which is why the value travels inside the polled word instead of behind a second
// dependent load + __threadfence.
// ws layout: [0,4) ticket, [4,8) done-count (mode 3), [64,...) state uint4[T] (mode 3: then partial ACC2[T]).
// The host wrapper zeroes ticket + states before each launch.
#ifndef KGB_SBLOCK
#define KGB_SBLOCK 256
#endif
#ifndef KGB_ITEMS
#define KGB_ITEMS (128 / (int)sizeof(KGB_ACC_T))
#endif
#define KGB_PADE (16 / (int)sizeof(KGB_ACC_T))
#define KGB_TILE (KGB_SBLOCK * KGB_ITEMS)
#define KGB_SMEM_ELEMS (KGB_SBLOCK * (KGB_ITEMS + KGB_PADE))
#define KGB_ST_EMPTY 0u
#define KGB_ST_AGG 1u
#define KGB_ST_PREFIX 2u

__device__ __forceinline__ int kgb_pad(int e) { return e + (e / KGB_ITEMS) * KGB_PADE; }

template <typename T, int BYTES = sizeof(T)> struct kgb_bits;
template <typename T> struct kgb_bits<T, 8> {
    static __device__ __forceinline__ u64 to(T v) { return *reinterpret_cast<u64*>(&v); }
    static __device__ __forceinline__ T from(u64 b) { return *reinterpret_cast<T*>(&b); }
};
template <typename T> struct kgb_bits<T, 4> {
    static __device__ __forceinline__ u64 to(T v) { return (u64)(*reinterpret_cast<u32*>(&v)); }
    static __device__ __forceinline__ T from(u64 b) { u32 x = (u32)b; return *reinterpret_cast<T*>(&x); }
};
template <typename T> struct kgb_bits<T, 1> {
    static __device__ __forceinline__ u64 to(T v) { return (u64)v; }
    static __device__ __forceinline__ T from(u64 b) { return (T)b; }
};

__device__ __forceinline__ void kgb_state_store(uint4* p, u32 status, KGB_ACC_T v) {
    const u64 b = kgb_bits<KGB_ACC_T>::to(v);
    const u64 w0 = ((u64)status << 32) | (b >> 32), w1 = ((u64)status << 32) | (b & 0xffffffffull);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// returns the status (EMPTY when the two halves disagree, i.e. a write is still landing)
__device__ __forceinline__ u32 kgb_state_load(const uint4* p, KGB_ACC_T& v) {
    u64 w0, w1;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    const u32 s0 = (u32)(w0 >> 32), s1 = (u32)(w1 >> 32);
    v = kgb_bits<KGB_ACC_T>::from((w0 << 32) | (w1 & 0xffffffffull));
    return (s0 == s1) ? s0 : KGB_ST_EMPTY;
}

template <int V> __device__ __forceinline__ void kgb_scan_body(const KgbParams& p) {
    KGB_SCALARS
    extern __shared__ __align__(16) unsigned char kgb_dyn_smem[];
    KGB_ACC_T* sbuf = reinterpret_cast<KGB_ACC_T*>(kgb_dyn_smem);  // KGB_SMEM_ELEMS accumulators
    __shared__ KGB_ACC_T swarp[KGB_SBLOCK / 32];
    __shared__ KGB_ACC_T s_excl;
    __shared__ u32 s_tile;
    const i64 n = p.n;
    const u32 ntiles = (u32)((n + KGB_TILE - 1) / KGB_TILE);
    u32* ticket = (u32*)p.ws;
    uint4* states = (uint4*)((char*)p.ws + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const i64 base = (i64)tile * KGB_TILE;
    const bool full = (base + KGB_TILE <= n);

    // ---- 1. striped load + evaluate -> padded smem
    if (full) {
#pragma unroll
        for (int h = 0; h < KGB_ITEMS / V; h += KGB_UNROLL) {
            KGB_DECL_REGS(V * KGB_UNROLL)
#pragma unroll
            for (int u = 0; u < KGB_UNROLL; u++) {
                const int e = ((h + u) * KGB_SBLOCK + tid) * V;
                KGB_LOAD_VEC(V, u * V, base + e)
            }
#pragma unroll
            for (int u = 0; u < KGB_UNROLL; u++) {
                const int e = ((h + u) * KGB_SBLOCK + tid) * V;
                KGB_ACC_T o[V];
#pragma unroll
                for (int j = 0; j < V; j++) o[j] = (KGB_ACC_T)(KGB_EVAL(u * V + j));
                if (V * sizeof(KGB_ACC_T) == 16) {
                    *reinterpret_cast<uint4*>(&sbuf[kgb_pad(e)]) = *reinterpret_cast<uint4*>(o);
                } else {
#pragma unroll
                    for (int j = 0; j < V; j++) sbuf[kgb_pad(e + j)] = o[j];
                }
            }
        }
    } else {
        for (int k = 0; k < KGB_ITEMS; k++) {
            const int e = k * KGB_SBLOCK + tid;
            KGB_ACC_T o = (KGB_ACC_T)(KGB_RED_IDENT);
            if (base + e < n) {
                KGB_DECL_REGS(1)
                KGB_LOAD_VEC(1, 0, base + e)
                o = (KGB_ACC_T)(KGB_EVAL(0));
            }
            sbuf[kgb_pad(e)] = o;
        }
    }
    __syncthreads();

    // ---- 2. blocked read + in-register scan
    KGB_ACC_T x[KGB_ITEMS];
    {
        const KGB_ACC_T* src = &sbuf[tid * (KGB_ITEMS + KGB_PADE)];
#pragma unroll
        for (int q = 0; q < KGB_ITEMS * (int)sizeof(KGB_ACC_T) / 16; q++) {
            uint4 r = reinterpret_cast<const uint4*>(src)[q];
            *reinterpret_cast<uint4*>(&x[q * (16 / (int)sizeof(KGB_ACC_T))]) = r;
        }
    }
#pragma unroll
    for (int k = 1; k < KGB_ITEMS; k++) x[k] = KGB_RED(x[k - 1], x[k]);
    // inclusive warp scan of the thread totals
    KGB_ACC_T tot = x[KGB_ITEMS - 1];
    KGB_ACC_T inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        KGB_ACC_T o = kgb_shfl_up<KGB_ACC_T>(inc, d);
        if (lane >= d) inc = KGB_RED(o, inc);
    }
    if (lane == 31) swarp[warp] = inc;
    KGB_ACC_T lane_excl = kgb_shfl_up<KGB_ACC_T>(inc, 1);
    __syncthreads();
    // fixed-order prefix over the warp totals
    KGB_ACC_T wpre = (KGB_ACC_T)(KGB_RED_IDENT);
    KGB_ACC_T tile_agg = (KGB_ACC_T)(KGB_RED_IDENT);
#pragma unroll
    for (int w = 0; w < KGB_SBLOCK / 32; w++) {
        const KGB_ACC_T sw = swarp[w];
        if (w == warp) wpre = tile_agg;
        tile_agg = (w == 0) ? sw : KGB_RED(tile_agg, sw);
    }
    // identities are exact for every fold operator (0, 1, -+inf / INT_MIN / INT_MAX), so no validity
    // tracking is needed when a prefix is still empty
    KGB_ACC_T thread_excl = (lane == 0) ? wpre : KGB_RED(wpre, lane_excl);

    // ---- 3. publish aggregate, look back
#ifdef KGB_HAS_CARRY
    // seed of the whole scan (the exclusive offset of this shard in a multi-GPU scan)
    const KGB_ACC_T carry = (KGB_ACC_T)(KGB_CARRY);
    if (tile == 0) {
        thread_excl = KGB_RED(carry, thread_excl);
        tile_agg = KGB_RED(carry, tile_agg);
    }
#endif
    if (tid == 0) kgb_state_store(&states[tile], tile == 0 ? KGB_ST_PREFIX : KGB_ST_AGG, tile_agg);
    if (warp == 0 && tile > 0) {
        KGB_ACC_T excl = (KGB_ACC_T)(KGB_RED_IDENT);
        i64 look = (i64)tile - 1;
        while (true) {
            const i64 idx = look - lane;
            u32 f = KGB_ST_PREFIX;  // virtual tiles before 0: an identity prefix
            KGB_ACC_T c = (KGB_ACC_T)(KGB_RED_IDENT);
            if (idx >= 0) f = kgb_state_load(&states[idx], c);
            // every lane of the warp takes part in the vote (lanes past tile 0 hold the virtual prefix)
            u32 spins = 0;
            while (__any_sync(0xffffffffu, f == KGB_ST_EMPTY)) {
                if (f == KGB_ST_EMPTY) f = kgb_state_load(&states[idx], c);
                if (++spins > (1u << 26)) __trap();  // a predecessor never published: fail, do not hang
            }
            const u32 pmask = __ballot_sync(0xffffffffu, f == KGB_ST_PREFIX);
            const int first = __ffs(pmask) - 1;  // nearest tile holding an inclusive prefix
            if (!((pmask == 0 || lane <= first) && idx >= 0)) c = (KGB_ACC_T)(KGB_RED_IDENT);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                KGB_ACC_T o = kgb_shfl_xor<KGB_ACC_T>(c, d);
                c = KGB_RED(c, o);
            }
            excl = KGB_RED(c, excl);
            if (pmask != 0) break;
            look -= 32;
        }
        if (lane == 0) {
            s_excl = excl;
            kgb_state_store(&states[tile], KGB_ST_PREFIX, KGB_RED(excl, tile_agg));
        }
    }
    __syncthreads();
    if (tile > 0) thread_excl = KGB_RED(s_excl, thread_excl);
#pragma unroll
    for (int k = 0; k < KGB_ITEMS; k++) x[k] = KGB_RED(thread_excl, x[k]);

    // ---- 4. blocked -> smem -> striped (coalesced) epilogue
    {
        KGB_ACC_T* dst = &sbuf[tid * (KGB_ITEMS + KGB_PADE)];
#pragma unroll
        for (int q = 0; q < KGB_ITEMS * (int)sizeof(KGB_ACC_T) / 16; q++)
            reinterpret_cast<uint4*>(dst)[q] = *reinterpret_cast<uint4*>(&x[q * (16 / (int)sizeof(KGB_ACC_T))]);
    }
    __syncthreads();
#if KGB_MODE == 2
    KGB_OUT_T* __restrict__ out = (KGB_OUT_T*)p.out;
    if (full) {
#pragma unroll
        for (int h = 0; h < KGB_ITEMS / V; h++) {
            const int e = (h * KGB_SBLOCK + tid) * V;
            KGB_OUT_T o[V];
#pragma unroll
            for (int j = 0; j < V; j++) o[j] = (KGB_OUT_T)sbuf[kgb_pad(e + j)];
            kgb_stv<KGB_OUT_T, V>(out + base + e, o);
        }
    } else {
        for (int k = 0; k < KGB_ITEMS; k++) {
            const int e = k * KGB_SBLOCK + tid;
            if (base + e < n) out[base + e] = (KGB_OUT_T)sbuf[kgb_pad(e)];
        }
    }
#else
    // ---- 4'. epilogue g(scan value, args[i]) folded with RED2.  The arguments are re-read stripe-wise with the
    // same 128-bit coalesced loads as step 1 (the tile was streamed moments ago, so these hit L2); a row-wise
    // re-read (one element per lane at a 128-byte stride) measured 19 % of roofline.
    KGB_ACC2_T r2 = (KGB_ACC2_T)(KGB_RED2_IDENT);
    if (full) {
#pragma unroll
        for (int h = 0; h < KGB_ITEMS / V; h += KGB_UNROLL) {
            KGB_DECL_REGS(V * KGB_UNROLL)
#pragma unroll
            for (int u = 0; u < KGB_UNROLL; u++) {
                const int e = ((h + u) * KGB_SBLOCK + tid) * V;
                KGB_LOAD_VEC(V, u * V, base + e)
            }
#pragma unroll
            for (int u = 0; u < KGB_UNROLL; u++) {
                const int e = ((h + u) * KGB_SBLOCK + tid) * V;
#pragma unroll
                for (int j = 0; j < V; j++) {
                    KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVAL2(u * V + j, sbuf[kgb_pad(e + j)]));
                    r2 = KGB_RED2(r2, e2);
                }
            }
        }
    } else {
        for (int k = 0; k < KGB_ITEMS; k++) {
            const int e = k * KGB_SBLOCK + tid;
            if (base + e < n) {
                KGB_DECL_REGS(1)
                KGB_LOAD_VEC(1, 0, base + e)
                KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVAL2(0, sbuf[kgb_pad(e)]));
                r2 = KGB_RED2(r2, e2);
            }
        }
    }
    __shared__ KGB_ACC2_T sh2[KGB_SBLOCK / 32];
    __shared__ bool is_last;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(r2, m);
        r2 = KGB_RED2(r2, o);
    }
    if (lane == 0) sh2[warp] = r2;
    __syncthreads();
    KGB_ACC2_T* part2 = (KGB_ACC2_T*)(states + ntiles);
    u32* done = ticket + 1;
    if (tid == 0) {
        KGB_ACC2_T a = sh2[0];
#pragma unroll
        for (int w = 1; w < KGB_SBLOCK / 32; w++) a = KGB_RED2(a, sh2[w]);
        ((volatile KGB_ACC2_T*)part2)[tile] = a;
        __threadfence();
        is_last = (atomicAdd(done, 1u) == ntiles - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        KGB_ACC2_T a = (KGB_ACC2_T)(KGB_RED2_IDENT);
        for (u32 b = tid; b < ntiles; b += KGB_SBLOCK) {
            KGB_ACC2_T o = ((volatile KGB_ACC2_T*)part2)[b];
            a = KGB_RED2(a, o);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(a, m);
            a = KGB_RED2(a, o);
        }
        __syncthreads();
        if (lane == 0) sh2[warp] = a;
        __syncthreads();
        if (tid == 0) {
            a = sh2[0];
#pragma unroll
            for (int w = 1; w < KGB_SBLOCK / 32; w++) a = KGB_RED2(a, sh2[w]);
            *(KGB_OUT_T*)p.out = (KGB_OUT_T)a;
        }
    }
#endif
}
extern "C" __global__ void __launch_bounds__(KGB_SBLOCK) KGB_KERNEL_NAME_V(const __grid_constant__ KgbParams p) { kgb_scan_body<KGB_VEC>(p); }
extern "C" __global__ void __launch_bounds__(KGB_SBLOCK) KGB_KERNEL_NAME_S(const __grid_constant__ KgbParams p) { kgb_scan_body<1>(p); }
#endif

// =====================================================================================================
#if (KGB_MODE == 2 || KGB_MODE == 3) && defined(KGB_PIPE)
// ------------------------------------------------------------------ SCAN, persistent pipelined variant
// Used when the expression reads exactly ONE 8-byte array (running sum / max / min / product of x, of x*2, ...).
// The generic kernel above spends ~45 % of its warp time at the barrier behind the look-back (ncu, r1c): a
// look-back is 2-4 dependent L2 round trips at ~1 us each under a saturated memory system, and nothing else
// in the CTA can run meanwhile.  Here ONE persistent CTA per SM takes the look-back off the critical path:
//
//   cp.async ring  : KGB_PS stages of 32 KB (4096 elements).  Raw input of tile i+P streams into its stage with
//                    16-byte cp.async (no registers held) while tiles i .. i-D are being worked on.
//   phase A (i)    : each compute thread reads its 128-byte row from the stage (XOR-swizzled 16-byte chunks:
//                    conflict-free both row-wise and stripe-wise), evaluates the expression, scans in registers,
//                    adds the CTA-local prefix and writes the row back IN PLACE; thread 0 publishes the tile
//                    aggregate and hands the tile to a look-back warp through an mbarrier.
//   look-back warps: KGB_PLW dedicated warps; warp (i mod KGB_PLW) resolves tile i's exclusive prefix with a
//                    128-tile window per round (4 independent state loads per lane) and signals an mbarrier.
//   phase B (i-D)  : D tiles later the prefix is (almost always) already there: stripe-wise read of the parked
//                    tile, add the prefix, coalesced 128-bit streaming stores.  No transposition on the way out —
//                    adding a scalar does not care about layout.
// Mode 3 (scan + reduce epilogue) stores nothing: phase A only reduces the row (no write-back), phase B re-reads the
// raw row, re-scans it from (tile prefix (+) the thread's exclusive prefix, kept in a register) and folds the epilogue;
// because that read is row-wise, the load that recycles a stage is issued after the barrier (mode 2 issues it before).
// Tiles come from the same atomic ticket as the generic kernel, so no co-residency assumption is made.
#ifndef KGB_PT
#define KGB_PT 256   // compute threads (256 or 512); a tile is always 4096 elements
#endif
#ifndef KGB_PS
#define KGB_PS 7     // stages
#define KGB_PP 2     // tiles in flight from global memory
#define KGB_PD 4     // tiles parked while their look-back runs
#define KGB_PLW 2    // look-back warps (2 measured best: 0.87 of roofline vs 0.82 with 4, 0.74 with 1)
#endif
#ifndef KGB_PLBW
#define KGB_PLBW 1   // predecessor states fetched per lane per look-back round (window = 32 * KGB_PLBW tiles)
#endif
#ifndef KGB_PMINB
#define KGB_PMINB 1  // CTAs per SM
#endif
#define KGB_PTILE 4096
#define KGB_PITEMS (KGB_PTILE / KGB_PT)   // elements per thread row: 16 or 8
#define KGB_PCR (KGB_PITEMS / 2)          // 16-byte chunks per row: 8 or 4
#define KGB_PCT (KGB_PTILE / 2 / KGB_PT)  // chunks per thread: 8 or 4
#define KGB_PSTAGE_BYTES (KGB_PTILE * 8)
#define KGB_PRING 16

struct KgbPipeCtl {
    u64 agg_bar[KGB_PS];
    u64 pre_bar[KGB_PS];
    KGB_ACC_T agg[KGB_PS];
    KGB_ACC_T excl[KGB_PS];
    u32 aggtile[KGB_PS + 1];
    u32 tile_ring[KGB_PRING];
    KGB_ACC_T swarp[KGB_PT / 32];
#if KGB_MODE == 3
    KGB_ACC2_T sh2[KGB_PT / 32];
#endif
};

__device__ __forceinline__ u32 kgb_smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kgb_mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kgb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kgb_mbar_arrive(u64* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(kgb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void kgb_mbar_wait(u64* bar, u32 parity) {
    const u32 a = kgb_smem_u32(bar);
    u32 done = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done) __nanosleep(64);  // idle partners should not eat the issue slots of the working warps
        if (!done && (++spins & 1023u) == 0) {  // a partner that never arrives: fail after ~4 s, do not hang
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void kgb_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kgb_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void kgb_bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(KGB_PT) : "memory"); }

// physical byte offset of 16-byte chunk `c` (0..7) of row `row` inside a stage
// (rows of 128 B: chunk ^ (row & 7); rows of 64 B: chunk ^ ((row >> 1) & 3) — either way the 8 threads of a
// quarter-warp hit 8 different 16-byte bank groups whether they walk down rows or along them)
__device__ __forceinline__ int kgb_swz(int row, int c) {
    return KGB_PCR == 8 ? row * 128 + ((c ^ (row & 7)) << 4) : row * 64 + ((c ^ ((row >> 1) & 3)) << 4);
}

extern "C" __global__ void __launch_bounds__(KGB_PT + 32 * KGB_PLW, KGB_PMINB) KGB_KERNEL_NAME_P(const __grid_constant__ KgbParams p) {
    KGB_SCALARS
    extern __shared__ __align__(128) unsigned char kgb_dyn_smem[];
    unsigned char* stages = kgb_dyn_smem;
    KgbPipeCtl& ctl = *reinterpret_cast<KgbPipeCtl*>(kgb_dyn_smem + KGB_PS * KGB_PSTAGE_BYTES);
    const i64 n = p.n;
    const u32 ntiles = (u32)((n + KGB_PTILE - 1) / KGB_PTILE);
    u32* ticket = (u32*)p.ws;
    uint4* states = (uint4*)((char*)p.ws + 64);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KGB_PIPE_T* __restrict__ src = (const KGB_PIPE_T*)KGB_PIPE_PTR;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < KGB_PS; s++) {
            kgb_mbar_init(&ctl.agg_bar[s], 1);
            kgb_mbar_init(&ctl.pre_bar[s], 1);
        }
        for (int j = 0; j < KGB_PP + 2; j++) ctl.tile_ring[j] = atomicAdd(ticket, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= KGB_PT / 32) {
        // ================================================================ look-back warps
        const u32 lw = (u32)warp - KGB_PT / 32;
        for (u32 i = lw;; i += KGB_PLW) {
            const u32 s = i % KGB_PS;
            kgb_mbar_wait(&ctl.agg_bar[s], (i / KGB_PS) & 1u);
            const u32 tile = ctl.aggtile[s];
            if (tile == 0xffffffffu) break;
            KGB_ACC_T excl = (KGB_ACC_T)(KGB_RED_IDENT);
            if (tile == 0) {
#ifdef KGB_HAS_CARRY
                excl = (KGB_ACC_T)(KGB_CARRY);
#endif
            } else {
                const KGB_ACC_T agg = ctl.agg[s];
                i64 look = (i64)tile - 1;
                bool done = false;
                while (!done) {
                    KGB_ACC_T c[KGB_PLBW];
                    u32 f[KGB_PLBW];
#pragma unroll
                    for (int j = 0; j < KGB_PLBW; j++) {
                        const i64 idx = look - (j * 32 + lane);
                        f[j] = KGB_ST_PREFIX;  // virtual tiles before 0: an identity prefix
                        c[j] = (KGB_ACC_T)(KGB_RED_IDENT);
                        if (idx >= 0) f[j] = kgb_state_load(&states[idx], c[j]);
                    }
#pragma unroll
                    for (int j = 0; j < KGB_PLBW; j++) {
                        if (!done) {
                            const i64 idx = look - (j * 32 + lane);
                            u32 spins = 0;
                            while (__any_sync(0xffffffffu, f[j] == KGB_ST_EMPTY)) {
                                if (f[j] == KGB_ST_EMPTY) f[j] = kgb_state_load(&states[idx], c[j]);
                                if (++spins > (1u << 24)) __trap();  // a predecessor never published: fail, do not hang
                            }
                            const u32 pmask = __ballot_sync(0xffffffffu, f[j] == KGB_ST_PREFIX);
                            const int first = __ffs(pmask) - 1;  // nearest tile holding an inclusive prefix
                            KGB_ACC_T v = ((pmask == 0 || lane <= first) && idx >= 0) ? c[j] : (KGB_ACC_T)(KGB_RED_IDENT);
#pragma unroll
                            for (int d = 16; d >= 1; d >>= 1) {
                                KGB_ACC_T o = kgb_shfl_xor<KGB_ACC_T>(v, d);
                                v = KGB_RED(v, o);
                            }
                            excl = KGB_RED(v, excl);
                            done = (pmask != 0);
                        }
                    }
                    look -= 32 * KGB_PLBW;
                }
                if (lane == 0) kgb_state_store(&states[tile], KGB_ST_PREFIX, KGB_RED(excl, agg));
            }
            if (lane == 0) {
                ctl.excl[s] = excl;
                kgb_mbar_arrive(&ctl.pre_bar[s]);
            }
            __syncwarp();
        }
        return;
    }

    // ==================================================================== compute warps (tid < KGB_PT)
#if KGB_MODE == 2
    KGB_OUT_T* __restrict__ out = (KGB_OUT_T*)p.out;
#else
    // mode 3: the epilogue g(scan value, args[i]) is folded into one accumulator per thread across all of the
    // CTA's tiles; one partial per CTA at the end (max/min/int folds are order-independent; a float sum here is
    // already association-dependent through the look-back window, like the scan itself)
    KGB_ACC2_T r2 = (KGB_ACC2_T)(KGB_RED2_IDENT);
#endif

    // stream one tile into its stage: 8 x 16-byte cp.async per thread, stripe-wise (coalesced)
    auto issue_load = [&](u32 li) {
        const u32 t = ctl.tile_ring[li % KGB_PRING];
        if (t < ntiles) {
            unsigned char* sb = stages + (size_t)(li % KGB_PS) * KGB_PSTAGE_BYTES;
            const i64 base = (i64)t * KGB_PTILE;
            if (base + KGB_PTILE <= n) {
#pragma unroll
                for (int h = 0; h < KGB_PCT; h++) {
                    const int cf = h * KGB_PT + tid;
                    kgb_cp_async16(sb + kgb_swz(cf / KGB_PCR, cf % KGB_PCR), src + base + cf * 2);
                }
            } else {  // the last, partial tile: guarded element loads (zero-filled beyond n, masked in phase A)
#pragma unroll
                for (int h = 0; h < KGB_PCT; h++) {
                    const int cf = h * KGB_PT + tid;
                    KGB_PIPE_T v0 = 0, v1 = 0;
                    if (base + cf * 2 < n) v0 = src[base + cf * 2];
                    if (base + cf * 2 + 1 < n) v1 = src[base + cf * 2 + 1];
                    KGB_PIPE_T* d = reinterpret_cast<KGB_PIPE_T*>(sb + kgb_swz(cf / KGB_PCR, cf % KGB_PCR));
                    d[0] = v0;
                    d[1] = v1;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // phase B: parked tile + exclusive prefix -> global, stripe-wise
#if KGB_MODE == 2
    auto phase_b = [&](u32 li, u32 t) {
#else
    auto phase_b = [&](u32 li, u32 t, KGB_ACC_T te) {
#endif
        const u32 s = li % KGB_PS;
        kgb_mbar_wait(&ctl.pre_bar[s], (li / KGB_PS) & 1u);
        const KGB_ACC_T excl = ctl.excl[s];
        const unsigned char* sb = stages + (size_t)s * KGB_PSTAGE_BYTES;
        const i64 base = (i64)t * KGB_PTILE;
        const bool full = (base + KGB_PTILE <= n);
#if KGB_MODE == 2
        // NOTE the chunk a thread reads here is exactly the chunk the same thread's cp.async overwrites when the stage
        // is reused (issue_load uses the same h * KGB_PT + tid mapping): that is what makes the stage hand-over safe
        // without another barrier.  A 256-bit store variant (two adjacent chunks per thread) broke this invariant
        // and raced with the prefetch; it bought no bandwidth either, so stores stay 128-bit.
#pragma unroll
        for (int h = 0; h < KGB_PCT; h++) {
            const int cf = h * KGB_PT + tid;
            KGB_ACC_T v[2];
            *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(sb + kgb_swz(cf / KGB_PCR, cf % KGB_PCR));
            KGB_OUT_T o[2];
            o[0] = (KGB_OUT_T)(KGB_RED(excl, v[0]));
            o[1] = (KGB_OUT_T)(KGB_RED(excl, v[1]));
            if (full) {
                kgb_stv<KGB_OUT_T, 2>(out + base + cf * 2, o);
            } else {
                if (base + cf * 2 < n) out[base + cf * 2] = o[0];
                if (base + cf * 2 + 1 < n) out[base + cf * 2 + 1] = o[1];
            }
        }
#else
        // Mode 3 never stores the scan: phase A only REDUCES the thread's row (nothing is written back to the stage), and
        // here the same thread re-reads its raw row, re-runs the serial scan from (tile prefix (+) its exclusive prefix
        // inside the tile, kept in a register since phase A) and folds the epilogue g(scan value, x) on the fly.  The
        // pipelined kernel exists only for programs with ONE array argument, so g's operand is the staged row itself:
        // no global re-read (the stripe-wise L2 re-read stalled the 8 compute warps: ncu r01v, DADD behind LDG).
        {
            KGB_ACC_T run = KGB_RED(excl, te);
            KGB_DECL_REGS(KGB_PITEMS)
#pragma unroll
            for (int q = 0; q < KGB_PCR; q++)
                *reinterpret_cast<uint4*>(&KGB_PIPE_ARG[q * 2]) = *reinterpret_cast<const uint4*>(sb + kgb_swz(tid, q));
            if (full) {
                // two half-rows scanned independently (the second from its own first element), then the first half's
                // total is folded into the second: 7 more folds, but dependent chains of 8 + 1 instead of 16; the
                // epilogue folds into four accumulators for the same reason
                constexpr int H = KGB_PITEMS / 2;
                KGB_ACC_T sv[KGB_PITEMS];
                sv[0] = KGB_RED(run, (KGB_ACC_T)(KGB_EVAL(0)));
                sv[H] = (KGB_ACC_T)(KGB_EVAL(H));
#pragma unroll
                for (int k = 1; k < H; k++) {
                    sv[k] = KGB_RED(sv[k - 1], (KGB_ACC_T)(KGB_EVAL(k)));
                    sv[H + k] = KGB_RED(sv[H + k - 1], (KGB_ACC_T)(KGB_EVAL(H + k)));
                }
#pragma unroll
                for (int k = H; k < KGB_PITEMS; k++) sv[k] = KGB_RED(sv[H - 1], sv[k]);
                KGB_ACC2_T q2[4];
#pragma unroll
                for (int k = 0; k < 4; k++) q2[k] = (KGB_ACC2_T)(KGB_EVAL2(k, sv[k]));
#pragma unroll
                for (int k = 4; k < KGB_PITEMS; k++) {
                    KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVAL2(k, sv[k]));
                    q2[k & 3] = KGB_RED2(q2[k & 3], e2);
                }
                r2 = KGB_RED2(r2, KGB_RED2(KGB_RED2(q2[0], q2[1]), KGB_RED2(q2[2], q2[3])));
            } else {
                for (int k = 0; k < KGB_PITEMS; k++) {
                    if (base + tid * KGB_PITEMS + k < n) {
                        run = KGB_RED(run, (KGB_ACC_T)(KGB_EVAL(k)));
                        KGB_ACC2_T e2 = (KGB_ACC2_T)(KGB_EVAL2(k, run));
                        r2 = KGB_RED2(r2, e2);
                    }
                }
            }
        }
#endif
    };

#pragma unroll
    for (int j = 0; j < KGB_PP; j++) issue_load((u32)j);

    u32 tq[KGB_PD];  // tiles parked in shared memory, oldest first
#if KGB_MODE == 3
    KGB_ACC_T teq[KGB_PD];  // this thread's exclusive prefix inside each parked tile
#pragma unroll
    for (int j = 0; j < KGB_PD; j++) teq[j] = (KGB_ACC_T)(KGB_RED_IDENT);
#endif
#pragma unroll
    for (int j = 0; j < KGB_PD; j++) tq[j] = 0xffffffffu;
    u32 tk = 0;
    u32 i = 0;
    for (;; i++) {
        const u32 s = i % KGB_PS;
        if (tid == 0) {
            if (i > 0) ctl.tile_ring[(i + KGB_PP + 1) % KGB_PRING] = tk;  // ticket taken one iteration ago
            tk = atomicAdd(ticket, 1u);                                    // for local tile i + P + 2
        }
        const u32 tile = ctl.tile_ring[i % KGB_PRING];
        if (tile >= ntiles) break;
#if KGB_MODE == 2
        issue_load(i + KGB_PP);
        asm volatile("cp.async.wait_group %0;" ::"n"(KGB_PP) : "memory");
        kgb_bar_compute();
#else
        // mode 3 reads the parked stage row-wise in phase B (not with the cp.async mapping), so the load that recycles
        // a stage is issued only after the barrier that puts every thread past that stage's phase B
        asm volatile("cp.async.wait_group %0;" ::"n"(KGB_PP - 1) : "memory");
        kgb_bar_compute();
        issue_load(i + KGB_PP);
#endif

        // ---- phase A
        unsigned char* sb = stages + (size_t)s * KGB_PSTAGE_BYTES;
        const i64 base = (i64)tile * KGB_PTILE;
        const bool full = (base + KGB_PTILE <= n);
        KGB_ACC_T x[KGB_PITEMS];
        {
            KGB_DECL_REGS(KGB_PITEMS)
#pragma unroll
            for (int q = 0; q < KGB_PCR; q++)
                *reinterpret_cast<uint4*>(&KGB_PIPE_ARG[q * 2]) = *reinterpret_cast<const uint4*>(sb + kgb_swz(tid, q));
#pragma unroll
            for (int k = 0; k < KGB_PITEMS; k++) {
                x[k] = (KGB_ACC_T)(KGB_EVAL(k));
                if (!full && base + tid * KGB_PITEMS + k >= n) x[k] = (KGB_ACC_T)(KGB_RED_IDENT);
            }
        }
#if KGB_MODE == 2
#pragma unroll
        for (int k = 1; k < KGB_PITEMS; k++) x[k] = KGB_RED(x[k - 1], x[k]);
        KGB_ACC_T inc = x[KGB_PITEMS - 1];
#else
        // only the row total is needed here: a pairwise tree (depth log2 instead of a 15-long dependent chain; with 8
        // compute warps per SM the float64 compare/select latency is what the kernel waits for)
#pragma unroll
        for (int w = 1; w < KGB_PITEMS; w <<= 1) {
#pragma unroll
            for (int k = 0; k + w < KGB_PITEMS; k += 2 * w) x[k] = KGB_RED(x[k], x[k + w]);
        }
        KGB_ACC_T inc = x[0];
#endif
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            KGB_ACC_T o = kgb_shfl_up<KGB_ACC_T>(inc, d);
            if (lane >= d) inc = KGB_RED(o, inc);
        }
        if (lane == 31) ctl.swarp[warp] = inc;
        KGB_ACC_T lane_excl = kgb_shfl_up<KGB_ACC_T>(inc, 1);
        kgb_bar_compute();
        KGB_ACC_T wpre = (KGB_ACC_T)(KGB_RED_IDENT);
        KGB_ACC_T tile_agg = (KGB_ACC_T)(KGB_RED_IDENT);
#pragma unroll
        for (int w = 0; w < KGB_PT / 32; w++) {
            const KGB_ACC_T sw = ctl.swarp[w];
            if (w == warp) wpre = tile_agg;
            tile_agg = (w == 0) ? sw : KGB_RED(tile_agg, sw);
        }
        const KGB_ACC_T thread_excl = (lane == 0) ? wpre : KGB_RED(wpre, lane_excl);
        if (tid == 0) {
            KGB_ACC_T pub = tile_agg;
#ifdef KGB_HAS_CARRY
            if (tile == 0) pub = KGB_RED((KGB_ACC_T)(KGB_CARRY), tile_agg);
#endif
            kgb_state_store(&states[tile], tile == 0 ? KGB_ST_PREFIX : KGB_ST_AGG, pub);
            ctl.agg[s] = tile_agg;
            ctl.aggtile[s] = tile;
            kgb_mbar_arrive(&ctl.agg_bar[s]);
        }
#if KGB_MODE == 2
#pragma unroll
        for (int k = 0; k < KGB_PITEMS; k++) x[k] = KGB_RED(thread_excl, x[k]);
#pragma unroll
        for (int q = 0; q < KGB_PCR; q++)
            *reinterpret_cast<uint4*>(sb + kgb_swz(tid, q)) = *reinterpret_cast<uint4*>(&x[q * 2]);
        // the row-wise write-back must be visible to the stripe-wise readers of phase B: the barrier at the top of
        // a later iteration orders it (phase B of this tile runs D >= 1 iterations from now)

        // ---- phase B of the tile parked D iterations ago
        if (tq[0] != 0xffffffffu) phase_b(i - KGB_PD, tq[0]);
#else
        if (tq[0] != 0xffffffffu) phase_b(i - KGB_PD, tq[0], teq[0]);
#pragma unroll
        for (int j = 0; j < KGB_PD - 1; j++) teq[j] = teq[j + 1];
        teq[KGB_PD - 1] = thread_excl;
#endif
#pragma unroll
        for (int j = 0; j < KGB_PD - 1; j++) tq[j] = tq[j + 1];
        tq[KGB_PD - 1] = tile;
    }
    // ---- drain: remaining parked tiles (i is the first local index without a tile)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    kgb_bar_compute();
#pragma unroll
    for (int j = 0; j < KGB_PD; j++) {
#if KGB_MODE == 2
        if (tq[j] != 0xffffffffu) phase_b(i - KGB_PD + j, tq[j]);
#else
        if (tq[j] != 0xffffffffu) phase_b(i - KGB_PD + j, tq[j], teq[j]);
#endif
    }
    kgb_bar_compute();
    if (tid == 0) {
        // release the look-back warps: one sentinel per warp, on the stages of the next KGB_PLW local indices
        for (u32 k = 0; k < KGB_PLW; k++) {
            const u32 s = (i + k) % KGB_PS;
            ctl.aggtile[s] = 0xffffffffu;
            kgb_mbar_arrive(&ctl.agg_bar[s]);
        }
    }
#if KGB_MODE == 3
    // one partial per CTA -> the last CTA to finish folds them in CTA order
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(r2, m);
        r2 = KGB_RED2(r2, o);
    }
    if (lane == 0) ctl.sh2[warp] = r2;
    kgb_bar_compute();
    if (warp == 0) {
        KGB_ACC2_T* part2 = (KGB_ACC2_T*)(states + ntiles);
        u32* done = ticket + 1;
        u32 last = 0;
        if (lane == 0) {
            KGB_ACC2_T a = ctl.sh2[0];
#pragma unroll
            for (int w = 1; w < KGB_PT / 32; w++) a = KGB_RED2(a, ctl.sh2[w]);
            ((volatile KGB_ACC2_T*)part2)[blockIdx.x] = a;
            __threadfence();
            last = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            KGB_ACC2_T a = (KGB_ACC2_T)(KGB_RED2_IDENT);
            for (u32 b = lane; b < gridDim.x; b += 32) {
                KGB_ACC2_T o = ((volatile KGB_ACC2_T*)part2)[b];
                a = KGB_RED2(a, o);
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                KGB_ACC2_T o = kgb_shfl_xor<KGB_ACC2_T>(a, m);
                a = KGB_RED2(a, o);
            }
            if (lane == 0) *(KGB_OUT_T*)p.out = (KGB_OUT_T)a;
        }
    }
#endif
}
#endif
