// find (a?b), where / nonzero, group (=a) and range (?a): compaction- and sort-based kernels.
#include "kgb_select.cuh"
#include "kgb_sort.cuh"

namespace kgb {

static int dispatch_nonzero(const void* a, int dtype, i64 n, i64* pos, i64* cnt, void* ws, size_t wsb, cudaStream_t st) {
    switch (dtype) {
        case KGB_F64: return launch_select(NonzeroPred<double>{(const double*)a}, n, pos, cnt, ws, wsb, st);
        case KGB_I64: return launch_select(NonzeroPred<i64>{(const i64*)a}, n, pos, cnt, ws, wsb, st);
        case KGB_F32: return launch_select(NonzeroPred<float>{(const float*)a}, n, pos, cnt, ws, wsb, st);
        case KGB_I32: return launch_select(NonzeroPred<int>{(const int*)a}, n, pos, cnt, ws, wsb, st);
        case KGB_B8: return launch_select(NonzeroPred<unsigned char>{(const unsigned char*)a}, n, pos, cnt, ws, wsb, st);
    }
    set_error("nonzero: unsupported dtype %d", dtype);
    return KGB_EINVAL;
}

// first[j] = idx[heads[j]] for j < *count
__global__ void __launch_bounds__(256) gather_counted_i64(const i64* __restrict__ src, const i64* __restrict__ pos,
                                                           const i64* __restrict__ count, i64* __restrict__ out) {
    const i64 c = *count;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < c; j += stride) out[j] = src[pos[j]];
}
// starts_out[G] = n  (sentinel closing the last group)
__global__ void close_starts(i64* starts, const i64* count, i64 n) { starts[*count] = n; }

template <typename T>
__global__ void __launch_bounds__(256) gather_typed(const T* __restrict__ a, const i64* __restrict__ pos, i64 c,
                                                     T* __restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < c; j += stride) out[j] = a[pos[j]];
}

static unsigned cap_grid(long long items) {
    long long need = (items + 255) / 256;
    long long cap = (long long)sm_count() * 16;
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

// group workspace: [enc sorted keys 8n][select ws][sort ws]
struct GroupWs {
    u64* enc;
    void* sel;
    size_t sel_bytes;
    void* sort;
    size_t sort_bytes;
    size_t total;
};
static GroupWs carve_group(void* ws, i64 n) {
    GroupWs g;
    char* p = (char*)ws;
    size_t o = 0;
    g.enc = (u64*)(p + o);
    o += align_up((size_t)n * 8, 256);
    g.sel = p + o;
    g.sel_bytes = select_ws_bytes(n);
    o += align_up(g.sel_bytes, 256);
    g.sort = p + o;
    g.sort_bytes = sort_ws_bytes(n);
    o += g.sort_bytes;
    g.total = o;
    return g;
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_select_workspace_bytes(int64_t n, size_t* out_bytes) {
    KGB_REQUIRE(out_bytes && n >= 0, "kgb_select_workspace_bytes: bad arguments");
    *out_bytes = select_ws_bytes(n);
    return KGB_OK;
}

extern "C" int kgb_nonzero(const void* flags, int32_t dtype, int64_t n, int64_t* pos_out, int64_t* count_out,
                           void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(count_out && n >= 0 && (n == 0 || (flags && pos_out)), "kgb_nonzero: bad arguments");
    return dispatch_nonzero(flags, dtype, n, (i64*)pos_out, (i64*)count_out, workspace, workspace_bytes,
                                         (cudaStream_t)stream);
}

extern "C" int kgb_find_equal(const void* a, int32_t dtype, int64_t n, const void* needle_host, int64_t* pos_out,
                              int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(count_out && needle_host && n >= 0 && (n == 0 || (a && pos_out)), "kgb_find_equal: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    i64* pos = (i64*)pos_out;
    i64* cnt = (i64*)count_out;
    switch (dtype) {
        case KGB_F64: return launch_select(EqualPred<double>{(const double*)a, *(const double*)needle_host}, n, pos, cnt, workspace, workspace_bytes, st);
        case KGB_I64: return launch_select(EqualPred<i64>{(const i64*)a, *(const i64*)needle_host}, n, pos, cnt, workspace, workspace_bytes, st);
        case KGB_F32: return launch_select(EqualPred<float>{(const float*)a, *(const float*)needle_host}, n, pos, cnt, workspace, workspace_bytes, st);
        case KGB_I32: return launch_select(EqualPred<int>{(const int*)a, *(const int*)needle_host}, n, pos, cnt, workspace, workspace_bytes, st);
    }
    set_error("kgb_find_equal: unsupported dtype %d", dtype);
    return KGB_EINVAL;
}

extern "C" int kgb_group_workspace_bytes(int32_t dtype, int64_t n, size_t* out_bytes) {
    KGB_REQUIRE(out_bytes && n >= 0 && dtype_size(dtype) >= 4, "kgb_group_workspace_bytes: bad arguments");
    *out_bytes = carve_group(nullptr, n).total;
    return KGB_OK;
}

extern "C" int kgb_group(const void* keys, int32_t dtype, int64_t n, int64_t* order_out, int64_t* starts_out,
                         int64_t* ngroups_out, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(ngroups_out && starts_out && n >= 0, "kgb_group: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        KGB_CUDA_CHECK(cudaMemsetAsync(ngroups_out, 0, 8, st));
        KGB_CUDA_CHECK(cudaMemsetAsync(starts_out, 0, 8, st));
        return KGB_OK;
    }
    KGB_REQUIRE(keys && order_out && workspace && workspace_bytes >= carve_group(nullptr, n).total,
                "kgb_group: null pointer or workspace too small");
    GroupWs g = carve_group(workspace, n);
    int rc = sort_pairs(keys, dtype, n, 0, (i64*)order_out, nullptr, g.enc, g.sort, g.sort_bytes, st);
    if (rc != KGB_OK) return rc;
    rc = launch_select(BoundaryPred{g.enc}, n, (i64*)starts_out, (i64*)ngroups_out, g.sel, g.sel_bytes, st);
    if (rc != KGB_OK) return rc;
    close_starts<<<1, 1, 0, st>>>((i64*)starts_out, (const i64*)ngroups_out, n);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

// unique workspace: [idx1 8n][heads 8n][first 8n][first_sorted_idx 8n] + group-style buffers
extern "C" int kgb_unique_first_workspace_bytes(int32_t dtype, int64_t n, size_t* out_bytes) {
    KGB_REQUIRE(out_bytes && n >= 0 && dtype_size(dtype) >= 4, "kgb_unique_first_workspace_bytes: bad arguments");
    *out_bytes = 4 * align_up((size_t)n * 8, 256) + carve_group(nullptr, n).total;
    return KGB_OK;
}

extern "C" int kgb_unique_first(const void* a, int32_t dtype, int64_t n, void* out, int64_t* first_idx_out,
                                int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(count_out && n >= 0, "kgb_unique_first: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        KGB_CUDA_CHECK(cudaMemsetAsync(count_out, 0, 8, st));
        return KGB_OK;
    }
    size_t need = 0;
    kgb_unique_first_workspace_bytes(dtype, n, &need);
    KGB_REQUIRE(a && out && first_idx_out && workspace && workspace_bytes >= need,
                "kgb_unique_first: null pointer or workspace too small");
    char* p = (char*)workspace;
    const size_t nb = align_up((size_t)n * 8, 256);
    i64* idx1 = (i64*)p;
    i64* heads = (i64*)(p + nb);
    i64* first = (i64*)(p + 2 * nb);
    i64* tmpidx = (i64*)(p + 3 * nb);
    GroupWs g = carve_group(p + 4 * nb, n);
    // 1. stable sort -> runs of equal keys, smallest position first inside each run
    const int sort_dt = dtype == KGB_F64 ? KGB_F64_RAW : dtype == KGB_F32 ? KGB_F32_RAW : dtype;
    int rc = sort_pairs(a, sort_dt, n, 0, idx1, nullptr, g.enc, g.sort, g.sort_bytes, st);
    if (rc != KGB_OK) return rc;
    // 2. run heads
    rc = launch_select(BoundaryPred{g.enc}, n, heads, (i64*)count_out, g.sel, g.sel_bytes, st);
    if (rc != KGB_OK) return rc;
    // 3. first occurrence position of every distinct key
    gather_counted_i64<<<cap_grid(n), 256, 0, st>>>(idx1, heads, (const i64*)count_out, first);
    KGB_CUDA_CHECK(cudaGetLastError());
    // the number of distinct keys sizes the second sort: one 8-byte D2H read (the caller needs it anyway)
    i64 c = 0;
    KGB_CUDA_CHECK(cudaMemcpyAsync(&c, count_out, 8, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    // 4. order of first appearance = ascending first positions
    rc = sort_pairs(first, KGB_I64, c, 0, tmpidx, first_idx_out, nullptr, g.sort, g.sort_bytes, st);
    if (rc != KGB_OK) return rc;
    // 5. the values
    const unsigned gr = cap_grid(c);
    switch (dtype) {
        case KGB_F64:
        case KGB_I64: gather_typed<u64><<<gr, 256, 0, st>>>((const u64*)a, (const i64*)first_idx_out, c, (u64*)out); break;
        default: gather_typed<u32><<<gr, 256, 0, st>>>((const u32*)a, (const i64*)first_idx_out, c, (u32*)out); break;
    }
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}
