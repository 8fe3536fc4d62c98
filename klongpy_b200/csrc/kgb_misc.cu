// Support verbs: iota (!n), gather (a@b), expand (&a), reverse (|a), array_equal.  All HBM-bound streaming
// kernels: coalesced accesses, grid sized to the data with a cap of a few waves over the 148 SMs.
#include <cstdlib>

#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

static inline unsigned grid_for(long long work_items, int block, int per_thread = 1) {
    long long need = (work_items + (long long)block * per_thread - 1) / ((long long)block * per_thread);
    long long cap = (long long)sm_count() * 16;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

__global__ void __launch_bounds__(256) iota_kernel(i64* __restrict__ out, i64 n, i64 start, i64 step) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 n2 = n / 2;
    // 128-bit stores when the base is 16-byte aligned
    if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n2; g += stride) {
            longlong2 v;
            v.x = start + (2 * g) * step;
            v.y = start + (2 * g + 1) * step;
            __stcs(reinterpret_cast<longlong2*>(out) + g, v);
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = start + (n - 1) * step;
    } else {
        for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = start + i * step;
    }
}

// gather of rows: out[i, :] = a[idx[i], :], rows of `units` words of type W
template <typename W>
__global__ void __launch_bounds__(256) gather_kernel(const W* __restrict__ a, i64 n_a, const i64* __restrict__ idx,
                                                      i64 n_idx, i64 units, W* __restrict__ out, int* err) {
    const i64 total = n_idx * units;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        i64 i, u;
        if (units == 1) {
            i = g;
            u = 0;
        } else {
            i = g / units;
            u = g - i * units;
        }
        i64 j = __ldg(&idx[i]);
        if (j < 0) j += n_a;
        W v = W(0);
        if (j < 0 || j >= n_a) {
            *err = 1;
        } else {
            v = __ldg(&a[j * units + u]);
        }
        out[g] = v;
    }
}

// &a : out[k] = i for incl[i-1] <= k < incl[i]   (binary search; incl is the inclusive scan of counts)
__global__ void __launch_bounds__(256) expand_kernel(const i64* __restrict__ incl, i64 n, i64 total,
                                                      i64* __restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        i64 lo = 0, hi = n - 1;  // first i with incl[i] > k
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            if (__ldg(&incl[mid]) > k)
                hi = mid;
            else
                lo = mid + 1;
        }
        out[k] = lo;
    }
}

template <typename W>
__global__ void __launch_bounds__(256) reverse_kernel(const W* __restrict__ a, i64 n, i64 units, W* __restrict__ out) {
    const i64 total = n * units;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        i64 i, u;
        if (units == 1) {
            i = g;
            u = 0;
        } else {
            i = g / units;
            u = g - i * units;
        }
        out[g] = __ldcs(&a[(n - 1 - i) * units + u]);
    }
}

__global__ void set_flag_kernel(int* flag, int v) { *flag = v; }

template <typename T>
__global__ void __launch_bounds__(256) array_equal_kernel(const T* __restrict__ a, const T* __restrict__ b, i64 n,
                                                           int* flag) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    bool ok = true;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T x = __ldcs(&a[i]), y = __ldcs(&b[i]);
        ok &= (x == y);  // value equality: NaN != NaN, -0.0 == 0.0 (np.array_equal)
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if ((threadIdx.x & 31) == 0) *flag = 0;
    }
}

// gather of single elements (`x@<x`, the common case): four independent index loads, then four independent
// random loads in flight per thread — a random 8-byte read costs a whole 32-byte sector, so the only lever is
// memory-level parallelism
template <typename W, int U, int BLOCK>
__global__ void __launch_bounds__(BLOCK) gather1_kernel(const W* __restrict__ a, i64 n_a, const i64* __restrict__ idx,
                                                       i64 n_idx, W* __restrict__ out, int* err) {
    // 32-bit arithmetic inside a CTA's contiguous chunk of U*256 indices: thread t owns chunk slots t, t+BLOCK, ... so the
    // index loads and the result stores of a warp are coalesced and U independent random loads are in flight per thread
    const u64 n_chunks = ((u64)n_idx + (U * BLOCK) - 1) / (U * BLOCK);
    bool bad = false;
    for (u64 c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const i64 base = (i64)(c * (U * BLOCK)) + threadIdx.x;
        i64 j[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 g = base + u * BLOCK;
            j[u] = g < n_idx ? __ldcs(&idx[g]) : 0;
            if (j[u] < 0) j[u] += n_a;
        }
        W v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = (u64)j[u] < (u64)n_a;
            v[u] = ok ? __ldg(&a[j[u]]) : W(0);
            bad = bad || (!ok && base + u * BLOCK < n_idx);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 g = base + u * BLOCK;
            if (g < n_idx) __stcs(&out[g], v[u]);
        }
    }
    if (bad) *err = 1;
}

template <typename W>
static int launch_gather(const void* a, i64 n_a, const i64* idx, i64 n_idx, i64 units, void* out, int* err,
                         cudaStream_t st) {
    if (units == 1) {
        // One small CTA per chunk of 4 x 256 indices, all chunks launched (no persistent loop): measured at 1e8 random
        // 8-byte reads (profiles/r02e_gather_probe.log) the many-small-CTA shape beats every persistent geometry — the block
        // scheduler keeps the SMs' load queues evenly full (2.16 ms = torch's index kernel = the DRAM random-sector rate).  KGB200_GATHER_U / _BLOCK / _CTAS are tuning overrides.
        int U = 4, block = 256, per_sm = 0;
        if (const char* e = getenv("KGB200_GATHER_U")) U = atoi(e);
        if (const char* e = getenv("KGB200_GATHER_BLOCK")) block = atoi(e);
        if (const char* e = getenv("KGB200_GATHER_CTAS")) per_sm = atoi(e);
        const long long chunks = (n_idx + (long long)U * block - 1) / ((long long)U * block);
        long long grid = per_sm > 0 ? (long long)sm_count() * per_sm : chunks;
        if (grid > chunks) grid = chunks;
        if (grid < 1) grid = 1;
#define KGB_GATHER1(UU, BB) gather1_kernel<W, UU, BB><<<(unsigned)grid, BB, 0, st>>>((const W*)a, n_a, idx, n_idx, (W*)out, err)
        if (block == 256) {
            if (U == 8) KGB_GATHER1(8, 256);
            else if (U == 16) KGB_GATHER1(16, 256);
            else KGB_GATHER1(4, 256);
        } else {
            if (U == 8) KGB_GATHER1(8, 128);
            else if (U == 2) KGB_GATHER1(2, 128);
            else KGB_GATHER1(4, 128);
        }
#undef KGB_GATHER1
    } else
        gather_kernel<W><<<grid_for(n_idx * units, 256, 2), 256, 0, st>>>((const W*)a, n_a, idx, n_idx, units, (W*)out, err);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}
template <typename W>
static int launch_reverse(const void* a, i64 n, i64 units, void* out, cudaStream_t st) {
    reverse_kernel<W><<<grid_for(n * units, 256, 2), 256, 0, st>>>((const W*)a, n, units, (W*)out);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_iota_i64(int64_t* out, int64_t n, int64_t start, int64_t step, void* stream) {
    KGB_REQUIRE(n >= 0 && (out || n == 0), "kgb_iota_i64: bad arguments");
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    if (n == 0) return KGB_OK;
    iota_kernel<<<grid_for(n / 2 + 1, 256, 4), 256, 0, (cudaStream_t)stream>>>((i64*)out, n, start, step);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

extern "C" int kgb_gather(const void* a, int32_t elem_bytes, int64_t n_a, const int64_t* idx, int64_t n_idx,
                          void* out, int32_t* err_flag, void* stream) {
    KGB_REQUIRE(a && idx && out && err_flag && elem_bytes > 0 && n_a >= 0 && n_idx >= 0, "kgb_gather: bad arguments");
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    if (n_idx == 0) return KGB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t al = (uintptr_t)a | (uintptr_t)out | (uintptr_t)elem_bytes;
    if ((al & 7) == 0) return launch_gather<u64>(a, n_a, (const i64*)idx, n_idx, elem_bytes / 8, out, err_flag, st);
    if ((al & 3) == 0) return launch_gather<u32>(a, n_a, (const i64*)idx, n_idx, elem_bytes / 4, out, err_flag, st);
    return launch_gather<unsigned char>(a, n_a, (const i64*)idx, n_idx, elem_bytes, out, err_flag, st);
}

extern "C" int kgb_expand(const int64_t* counts, const int64_t* incl_offsets, int64_t n, int64_t total, int64_t* out,
                          void* stream) {
    (void)counts;
    KGB_REQUIRE(incl_offsets && n > 0 && total >= 0 && (out || total == 0), "kgb_expand: bad arguments");
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    if (total == 0) return KGB_OK;
    expand_kernel<<<grid_for(total, 256, 2), 256, 0, (cudaStream_t)stream>>>((const i64*)incl_offsets, n, total, (i64*)out);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

extern "C" int kgb_reverse(const void* a, int32_t elem_bytes, int64_t n, void* out, void* stream) {
    KGB_REQUIRE(a && out && elem_bytes > 0 && n >= 0, "kgb_reverse: bad arguments");
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    if (n == 0) return KGB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t al = (uintptr_t)a | (uintptr_t)out | (uintptr_t)elem_bytes;
    if ((al & 7) == 0) return launch_reverse<u64>(a, n, elem_bytes / 8, out, st);
    if ((al & 3) == 0) return launch_reverse<u32>(a, n, elem_bytes / 4, out, st);
    return launch_reverse<unsigned char>(a, n, elem_bytes, out, st);
}

extern "C" int kgb_array_equal(const void* a, const void* b, int32_t dtype, int64_t n, int32_t* out_flag, void* stream) {
    KGB_REQUIRE(out_flag && n >= 0 && dtype_size(dtype) > 0, "kgb_array_equal: bad arguments");
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    cudaStream_t st = (cudaStream_t)stream;
    set_flag_kernel<<<1, 1, 0, st>>>(out_flag, 1);
    if (n > 0) {
        KGB_REQUIRE(a && b, "kgb_array_equal: null input");
        const unsigned g = grid_for(n, 256, 4);
        switch (dtype) {
            case KGB_F64: array_equal_kernel<double><<<g, 256, 0, st>>>((const double*)a, (const double*)b, n, out_flag); break;
            case KGB_I64: array_equal_kernel<i64><<<g, 256, 0, st>>>((const i64*)a, (const i64*)b, n, out_flag); break;
            case KGB_F32: array_equal_kernel<float><<<g, 256, 0, st>>>((const float*)a, (const float*)b, n, out_flag); break;
            case KGB_I32: array_equal_kernel<int><<<g, 256, 0, st>>>((const int*)a, (const int*)b, n, out_flag); break;
            default: array_equal_kernel<unsigned char><<<g, 256, 0, st>>>((const unsigned char*)a, (const unsigned char*)b, n, out_flag); break;
        }
    }
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}
