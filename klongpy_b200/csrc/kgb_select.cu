// find (a?b), where / nonzero, group (=a) and range (?a): compaction- and sort-based kernels.
#include <cstdlib>

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


// ------------------------------------------------------------------------------------------- range (?a) by hashing
// Low-cardinality inputs (10 k stations in 1e8 rows) do not need a sort: an L2-resident open-addressing table keyed by
// the raw bits keeps, per distinct key, the smallest position that holds it.  A row whose key is already present
// with a smaller position costs one 16-byte L2 read and no atomic — which is every row but the first few of each key —
// so the pass runs near the streaming rate.  The table has 2^20 slots (16 MB); when more than half fill up the kernel
// raises a flag, every CTA stops at its next iteration and the caller falls back to the sort-based path, so the
// attempt costs a high-cardinality input a few tens of microseconds.
struct __align__(16) UniqSlot {
    u64 key;    // raw bits; UNIQ_EMPTY while free
    u64 first;  // smallest position seen so far
};
constexpr u64 UNIQ_EMPTY = ~0ull;
constexpr int UNIQ_LOG2_SLOTS = 20;
constexpr u64 UNIQ_SLOTS = 1ull << UNIQ_LOG2_SLOTS;
constexpr u64 UNIQ_LIMIT = UNIQ_SLOTS / 2;
struct UniqCtl {
    unsigned long long distinct;       // keys inserted so far
    unsigned long long special_first;  // first position of the key whose bits equal UNIQ_EMPTY (all ones)
    unsigned long long out_count;      // collect kernel's cursor
    int overflow;
};

__device__ __forceinline__ u64 uniq_hash(u64 k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

template <typename T> __device__ __forceinline__ u64 uniq_bits(T v);
template <> __device__ __forceinline__ u64 uniq_bits<u64>(u64 v) { return v; }
template <> __device__ __forceinline__ u64 uniq_bits<u32>(u32 v) { return (u64)v; }

template <typename T>
__global__ void __launch_bounds__(256) uniq_insert_kernel(const T* __restrict__ a, i64 m, i64 step, UniqSlot* __restrict__ tab,
                                                           UniqCtl* __restrict__ ctl) {
    // visits rows j * step for j < m: step 1 is the full pass, a larger step the cardinality probe on a strided sample
    constexpr int U = 4;
    const i64 stride = (i64)gridDim.x * 256 * U;
    u32 fresh = 0;
    for (i64 base = (i64)blockIdx.x * 256 * U; base < m; base += stride) {
        if (__any_sync(0xffffffffu, *(volatile int*)&ctl->overflow != 0)) return;  // warp-uniform exit
        bool bail = false;
        T v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 j = base + u * 256 + threadIdx.x;
            v[u] = a[(j < m ? j : m - 1) * step];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 j = base + u * 256 + threadIdx.x;
            if (j >= m || bail) continue;
            const i64 i = j * step;
            const u64 k = uniq_bits<T>(v[u]);
            if (k == UNIQ_EMPTY) {
                if ((u64)i < *(volatile unsigned long long*)&ctl->special_first) atomicMin(&ctl->special_first, (unsigned long long)i);
                continue;
            }
            u64 slot = uniq_hash(k) & (UNIQ_SLOTS - 1);
            for (u64 probe = 0; probe < UNIQ_SLOTS; probe++) {
                ulonglong2 e;  // one 16-byte relaxed load of {key, first}; either half may be stale, both only move one way
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(e.x), "=l"(e.y) : "l"(&tab[slot]) : "memory");
                if (e.x == k) {
                    if ((u64)i < e.y) atomicMin((unsigned long long*)&tab[slot].first, (unsigned long long)i);
                    break;
                }
                if (e.x == UNIQ_EMPTY) {
                    // no new keys once the table is over its limit: the rows in flight when the flag goes up would
                    // otherwise fill it completely and turn every probe sequence into a scan of the whole table
                    if (*(volatile int*)&ctl->overflow) {
                        bail = true;
                        break;
                    }
                    const u64 prev = atomicCAS((unsigned long long*)&tab[slot].key, (unsigned long long)UNIQ_EMPTY, (unsigned long long)k);
                    if (prev == UNIQ_EMPTY) {
                        fresh++;
                        atomicMin((unsigned long long*)&tab[slot].first, (unsigned long long)i);
                        break;
                    }
                    if (prev == k) {
                        atomicMin((unsigned long long*)&tab[slot].first, (unsigned long long)i);
                        break;
                    }
                }
                slot = (slot + 1) & (UNIQ_SLOTS - 1);
                if ((probe & 15) == 15) {
                    // a probe sequence this long means the table is far past half full (the counter below lags by
                    // one iteration of every thread): give up, the caller sorts instead
                    if (probe >= 255) ctl->overflow = 1;
                    if (*(volatile int*)&ctl->overflow) {
                        bail = true;
                        break;
                    }
                }
            }
        }
        // new keys are counted once per warp and iteration, not once per insertion: a permutation would otherwise
        // serialise half a million atomics on this one word (measured: 1 ms)
        fresh = __reduce_add_sync(0xffffffffu, fresh);
        if ((threadIdx.x & 31) == 0 && fresh) {
            if (atomicAdd(&ctl->distinct, (unsigned long long)fresh) + fresh > UNIQ_LIMIT) ctl->overflow = 1;
        }
        fresh = 0;
        if (__any_sync(0xffffffffu, bail)) return;
    }
}

// The full pass with a per-CTA cache in front of the table.  One persistent 1024-thread CTA per SM walks its tiles in
// ascending order; a 16384-entry direct-mapped cache of keys in shared memory remembers which keys THIS CTA has already
// taken to the table.  Entries are inserted only at the end of an iteration (between two barriers), so a hit always
// refers to an earlier iteration of the same CTA, i.e. to a smaller position: the row cannot be a first occurrence and
// costs one shared-memory load instead of a 16-byte L2 read.  With 10 k keys three rows in four hit the cache (the
// keys that share a cache line evict each other); with more keys than entries the pass degrades to the plain one.
constexpr int UNIQ_CACHE_LOG2 = 14;
constexpr int UNIQ_CACHE = 1 << UNIQ_CACHE_LOG2;
constexpr int UNIQ_CT = 1024;
constexpr int UNIQ_CU = 4;

// the table protocol for one row (key k at position i), starting at `slot`: shared by both passes' slow paths
__device__ __forceinline__ void uniq_probe(UniqSlot* __restrict__ tab, UniqCtl* __restrict__ ctl, u64 k, i64 i, u64 slot, u32& fresh,
                                           bool& bail) {
    for (u64 probe = 0; probe < UNIQ_SLOTS; probe++) {
        ulonglong2 e;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(e.x), "=l"(e.y) : "l"(&tab[slot]) : "memory");
        if (e.x == k) {
            if ((u64)i < e.y) atomicMin((unsigned long long*)&tab[slot].first, (unsigned long long)i);
            return;
        }
        if (e.x == UNIQ_EMPTY) {
            if (*(volatile int*)&ctl->overflow) {
                bail = true;
                return;
            }
            const u64 prev = atomicCAS((unsigned long long*)&tab[slot].key, (unsigned long long)UNIQ_EMPTY, (unsigned long long)k);
            if (prev == UNIQ_EMPTY) fresh++;
            if (prev == UNIQ_EMPTY || prev == k) {
                atomicMin((unsigned long long*)&tab[slot].first, (unsigned long long)i);
                return;
            }
        }
        slot = (slot + 1) & (UNIQ_SLOTS - 1);
        if ((probe & 15) == 15) {
            if (probe >= 255) ctl->overflow = 1;
            if (*(volatile int*)&ctl->overflow) {
                bail = true;
                return;
            }
        }
    }
}

constexpr size_t uniq_cached_smem() { return (size_t)UNIQ_CACHE * 8; }

// Measured (1e8 int64 rows, 10 k keys / 53 keys incl. the all-ones key / 1002 float64 keys, whole primitive):
// plain pass 0.88 / 1.31 / 0.79 ms; this kernel 0.77 / 0.48 / 0.57 ms.  A variant that staged the input through a
// three-stage cp.async ring (64 KB in flight per SM) was slower (0.86 / 0.55 / 0.65 ms): ncu r01y_uniq_cached_raw.csv
// shows the pass issue-bound (126 warp-instructions per 32 rows, most of them the 64-bit hash and index arithmetic),
// not latency-bound, so the ring's extra shared-memory traffic only added work.
template <typename T>
__global__ void __launch_bounds__(UNIQ_CT, 1) uniq_insert_cached_kernel(const T* __restrict__ a, i64 n, UniqSlot* __restrict__ tab,
                                                                        UniqCtl* __restrict__ ctl) {
    extern __shared__ __align__(16) u64 s_cache[];
    __shared__ int s_special;  // this CTA has already reported the all-ones key at a smaller position
    constexpr int U = UNIQ_CU;
    const int tid = threadIdx.x;
    for (int c = tid; c < UNIQ_CACHE; c += UNIQ_CT) s_cache[c] = UNIQ_EMPTY;
    if (tid == 0) s_special = 0;
    __syncthreads();
    const i64 tile = (i64)UNIQ_CT * U;
    const i64 stride = (i64)gridDim.x * tile;
    i64 base = (i64)blockIdx.x * tile;
    T cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const i64 j = base + u * UNIQ_CT + tid;
        cur[u] = a[j < n ? j : n - 1];
    }
    u32 fresh = 0;
    for (; base < n; base += stride) {
        const i64 nb = base + stride;  // next tile's loads fly while this one is looked up
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 j = nb + u * UNIQ_CT + tid;
            nxt[u] = a[j < n ? j : n - 1];
        }
        // the abort flag is read by one thread and only consumed at the barrier below: a global load every thread
        // waits for would put an L2 round trip on the critical path of every iteration
        bool bail = tid == 0 ? *(volatile int*)&ctl->overflow != 0 : false;
        bool special = false;
        u32 miss = 0;
        u64 hs[U];
        // 1. the cache: shared memory only
#pragma unroll
        for (int u = 0; u < U; u++) {
            const i64 i = base + u * UNIQ_CT + tid;
            const u64 k = uniq_bits<T>(cur[u]);
            hs[u] = uniq_hash(k);
            if (i >= n) continue;
            if (k == UNIQ_EMPTY) {
                if (!s_special) {
                    atomicMin(&ctl->special_first, (unsigned long long)i);
                    special = true;
                }
                continue;
            }
            if (s_cache[(hs[u] >> 40) & (UNIQ_CACHE - 1)] != k) miss |= 1u << u;
        }
        // 2. the misses: all table reads are issued before the first one is looked at
        ulonglong2 e[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (miss & (1u << u))
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];"
                             : "=l"(e[u].x), "=l"(e[u].y)
                             : "l"(&tab[hs[u] & (UNIQ_SLOTS - 1)])
                             : "memory");
#pragma unroll
        for (int u = 0; u < U; u++)
            if (miss & (1u << u)) {
                const i64 i = base + u * UNIQ_CT + tid;
                const u64 k = uniq_bits<T>(cur[u]);
                if (e[u].x == k) {  // the common case: present; touch it only if this row comes earlier
                    if ((u64)i < e[u].y) atomicMin((unsigned long long*)&tab[hs[u] & (UNIQ_SLOTS - 1)].first, (unsigned long long)i);
                } else {
                    uniq_probe(tab, ctl, k, i, hs[u] & (UNIQ_SLOTS - 1), fresh, bail);
                }
            }
        fresh = __reduce_add_sync(0xffffffffu, fresh);
        if ((tid & 31) == 0 && fresh) {
            if (atomicAdd(&ctl->distinct, (unsigned long long)fresh) + fresh > UNIQ_LIMIT) ctl->overflow = 1;
        }
        fresh = 0;
        if (__syncthreads_or(bail)) return;  // every lookup of this iteration is done; leave together on overflow
        // 3. remember what went to the table (visible to the NEXT iteration's lookups only)
#pragma unroll
        for (int u = 0; u < U; u++)
            if (miss & (1u << u)) s_cache[(hs[u] >> 40) & (UNIQ_CACHE - 1)] = uniq_bits<T>(cur[u]);
        if (special) s_special = 1;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; u++) cur[u] = nxt[u];
    }
}

// first positions of all occupied slots, in no particular order (they are sorted afterwards)
__global__ void __launch_bounds__(256) uniq_collect_kernel(const UniqSlot* __restrict__ tab, UniqCtl* __restrict__ ctl,
                                                            i64* __restrict__ first, i64* __restrict__ count_out) {
    const u64 slot = (u64)blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool occ = false;
    u64 f = 0;
    if (slot < UNIQ_SLOTS) {
        const UniqSlot e = tab[slot];
        occ = e.key != UNIQ_EMPTY;
        f = e.first;
    } else if (slot == UNIQ_SLOTS) {
        f = ctl->special_first;
        occ = f != ~0ull;
    }
    const u32 m = __ballot_sync(0xffffffffu, occ);
    if (m) {
        unsigned long long b = 0;
        if (lane == __ffs(m) - 1) b = atomicAdd(&ctl->out_count, (unsigned long long)__popc(m));
        b = __shfl_sync(0xffffffffu, b, __ffs(m) - 1);
        if (occ) first[b + __popc(m & ((1u << lane) - 1u))] = (i64)f;
    }
    (void)count_out;
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
    int rc;
    i64 c = -1;
    // 0. low-cardinality fast path: first positions through an L2-resident hash table (no sort of the n rows)
    static const bool hash_on = !(getenv("KGB200_UNIQUE_HASH") && atoi(getenv("KGB200_UNIQUE_HASH")) == 0);
    const size_t tab_bytes = (size_t)UNIQ_SLOTS * sizeof(UniqSlot);
    if (hash_on && nb >= tab_bytes && nb >= 256 + sizeof(UniqCtl)) {
        UniqSlot* tab = (UniqSlot*)idx1;          // the table borrows the sort path's index buffer
        UniqCtl* ctl = (UniqCtl*)heads;
        KGB_CUDA_CHECK(cudaMemsetAsync(tab, 0xff, tab_bytes, st));
        KGB_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(UniqCtl), st));
        KGB_CUDA_CHECK(cudaMemsetAsync(&ctl->special_first, 0xff, 8, st));
        auto insert = [&](i64 m, i64 step) {
            long long ib = (m + 1023) / 1024;
            const long long icap = (long long)sm_count() * 8;
            if (ib > icap) ib = icap;
            if (dtype_size(dtype) == 8)
                uniq_insert_kernel<u64><<<(unsigned)ib, 256, 0, st>>>((const u64*)a, m, step, tab, ctl);
            else
                uniq_insert_kernel<u32><<<(unsigned)ib, 256, 0, st>>>((const u32*)a, m, step, tab, ctl);
        };
        UniqCtl h;
        // cardinality probe: 512k rows spread over the whole input go in first (they are real rows, nothing is wasted);
        // if they alone bring in more than 300k distinct keys the table would overflow — go straight to the sort
        const i64 sample = n < (1ll << 19) ? n : (1ll << 19);
        insert(sample, n / sample);
        KGB_CUDA_CHECK(cudaMemcpyAsync(&h, ctl, sizeof(UniqCtl), cudaMemcpyDeviceToHost, st));
        KGB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (!h.overflow && h.distinct <= 300000ull) {
            // the full pass: persistent CTAs with the shared-memory key cache in front of the table
            static const bool cache_on = !(getenv("KGB200_UNIQUE_CACHE") && atoi(getenv("KGB200_UNIQUE_CACHE")) == 0);
            if (cache_on) {
                static bool opted[64] = {};
                const int dv = current_device();
                const size_t smem8 = uniq_cached_smem(), smem4 = uniq_cached_smem();
                if (dv >= 0 && dv < 64 && !opted[dv]) {
                    KGB_CUDA_CHECK(cudaFuncSetAttribute(uniq_insert_cached_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
                    KGB_CUDA_CHECK(cudaFuncSetAttribute(uniq_insert_cached_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
                    opted[dv] = true;
                }
                long long cb = (n + UNIQ_CT * UNIQ_CU - 1) / (UNIQ_CT * UNIQ_CU);
                if (cb > sm_count()) cb = sm_count();
                if (dtype_size(dtype) == 8)
                    uniq_insert_cached_kernel<u64><<<(unsigned)cb, UNIQ_CT, smem8, st>>>((const u64*)a, n, tab, ctl);
                else
                    uniq_insert_cached_kernel<u32><<<(unsigned)cb, UNIQ_CT, smem4, st>>>((const u32*)a, n, tab, ctl);
            } else {
                insert(n, 1);
            }
            uniq_collect_kernel<<<(unsigned)(UNIQ_SLOTS / 256 + 1), 256, 0, st>>>(tab, ctl, first, (i64*)count_out);
            KGB_CUDA_CHECK(cudaGetLastError());
            KGB_CUDA_CHECK(cudaMemcpyAsync(&h, ctl, sizeof(UniqCtl), cudaMemcpyDeviceToHost, st));
            KGB_CUDA_CHECK(cudaStreamSynchronize(st));
            if (!h.overflow) {
                c = (i64)h.out_count;
                KGB_CUDA_CHECK(cudaMemcpyAsync(count_out, &ctl->out_count, 8, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    if (c < 0) {
        // 1. stable sort -> runs of equal keys, smallest position first inside each run
        const int sort_dt = dtype == KGB_F64 ? KGB_F64_RAW : dtype == KGB_F32 ? KGB_F32_RAW : dtype;
        rc = sort_pairs(a, sort_dt, n, 0, idx1, nullptr, g.enc, g.sort, g.sort_bytes, st);
        if (rc != KGB_OK) return rc;
        // 2. run heads
        rc = launch_select(BoundaryPred{g.enc}, n, heads, (i64*)count_out, g.sel, g.sel_bytes, st);
        if (rc != KGB_OK) return rc;
        // 3. first occurrence position of every distinct key
        gather_counted_i64<<<cap_grid(n), 256, 0, st>>>(idx1, heads, (const i64*)count_out, first);
        KGB_CUDA_CHECK(cudaGetLastError());
        // the number of distinct keys sizes the second sort: one 8-byte D2H read (the caller needs it anyway)
        KGB_CUDA_CHECK(cudaMemcpyAsync(&c, count_out, 8, cudaMemcpyDeviceToHost, st));
        KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    }
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
