// Grouped min / max / sum / count over (int64 key, float64 value) rows — the 1brc-style aggregation of
// examples/1brc/worker.kg:3-5 (collect = [min max sum count], merge = (&, |, +, +)).
//
// v1: streaming 128-bit loads of both columns; the G x 4 table lives in L2-resident global memory.
// min/max use order-encoded 64-bit integers with a read-before-update filter (after the first few rows of a
// station almost no row improves its extremes, so the RED is rarely issued); sum/count are RED.ADD.
#include <algorithm>
#include <cstdlib>

#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

struct AggWs {
    u64* tmin;  // order-encoded, NaN = 0
    u64* tmax;  // order-encoded, NaN = ~0, empty = 0
    double* tsum;
    u64* tcnt;
    int* err;
    double* psum;  // [parts][G] per-CTA partial sums   (all-shared-memory variant: plain stores, folded in CTA order)
    u32* pcnt;     // [parts][G] per-CTA partial counts
    size_t total;
};
constexpr int AGG_MAX_PARTS = 160;  // >= SM count: one persistent CTA per SM
constexpr size_t AGG_SMEM_LIMIT = 200 * 1024;
// The per-CTA partial tables exist only for the variants that keep sums and counts in shared memory (16 B per group):
// ONE predicate for the workspace query and the launcher, so G = 4e6 sparse keys cost 32*G bytes, not 1920*G.
static inline bool agg_has_parts(i64 G) { return (size_t)G * 16 <= AGG_SMEM_LIMIT; }
static AggWs carve_agg(void* ws, i64 G) {
    AggWs a;
    char* p = (char*)ws;
    const size_t gb = align_up((size_t)G * 8, 256);
    const size_t parts = agg_has_parts(G) ? (size_t)AGG_MAX_PARTS : 0;
    a.tmin = (u64*)p;
    a.tmax = (u64*)(p + gb);
    a.tsum = (double*)(p + 2 * gb);
    a.tcnt = (u64*)(p + 3 * gb);
    a.err = (int*)(p + 4 * gb);
    a.psum = (double*)(p + 4 * gb + 256);
    a.pcnt = (u32*)(p + 4 * gb + 256 + parts * gb);
    a.total = 4 * gb + 256 + parts * gb + parts * align_up((size_t)G * 4, 256);
    return a;
}

__device__ __forceinline__ u64 enc_for_min(double v) { return (v != v) ? 0ull : key_from_f64(v); }
__device__ __forceinline__ u64 enc_for_max(double v) { return key_from_f64(v); }  // NaN -> ~0

__device__ __forceinline__ void agg_row(i64 k, double v, i64 G, const AggWs& a) {
    if ((u64)k >= (u64)G) {
        *a.err = 1;
        return;
    }
    const u64 emin = enc_for_min(v), emax = enc_for_max(v);
    if (emin < ld_relaxed_u64(&a.tmin[k])) atomicMin(&a.tmin[k], emin);
    if (emax > ld_relaxed_u64(&a.tmax[k])) atomicMax(&a.tmax[k], emax);
    atomicAdd(&a.tsum[k], v);
    atomicAdd(&a.tcnt[k], 1ull);
}

__global__ void __launch_bounds__(256) groupagg_kernel(const i64* __restrict__ keys, const double* __restrict__ vals,
                                                        i64 n, i64 G, AggWs a) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    constexpr int U = 4;
    if (vec) {
        const i64 n2 = n / 2;
        for (i64 g0 = tid; g0 < n2; g0 += stride * U) {
            longlong2 k[U];
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    k[u] = __ldcs(reinterpret_cast<const longlong2*>(keys) + g);
                    v[u] = __ldcs(reinterpret_cast<const double2*>(vals) + g);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    agg_row(k[u].x, v[u].x, G, a);
                    agg_row(k[u].y, v[u].y, G, a);
                }
            }
        }
        if ((n & 1) && tid == 0) agg_row(keys[n - 1], vals[n - 1], G, a);
    } else {
        for (i64 i = tid; i < n; i += stride) agg_row(keys[i], vals[i], G, a);
    }
}

// Filter seeding: the exact extremes of a SAMPLE of the rows (the first `m`) go into the global table before the
// main kernel starts, and every CTA initialises its shared-memory filters from that table.  Without it each CTA
// re-learns every station's extremes from scratch — ~ln(rows per station per CTA) record-breaking rows per station
// per CTA take the exact path (14 % of all rows at 1e8 rows / 10 k stations / 148 CTAs).
__global__ void __launch_bounds__(256) groupagg_seed_kernel(const i64* __restrict__ keys, const double* __restrict__ vals,
                                                             i64 m, i64 G, AggWs a) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const i64 k = keys[i];
        if ((u64)k >= (u64)G) continue;  // reported by the main kernel
        const double v = vals[i];
        const u64 emin = enc_for_min(v), emax = enc_for_max(v);
        if (emin < ld_relaxed_u64(&a.tmin[k])) atomicMin(&a.tmin[k], emin);
        if (emax > ld_relaxed_u64(&a.tmax[k])) atomicMax(&a.tmax[k], emax);
    }
}

// v2: per-CTA shared-memory state for everything that does not need a 64-bit float atomic.
//   cnt[k]        u32 row count of this CTA (native ATOMS.ADD), flushed to the global table at the end
//   minhi/maxhi   upper 32 bits of the order-encoded running extremes seen by this CTA: a row whose upper half
//                 cannot beat them is filtered with two LDS and touches no atomic unit; the exact 64-bit
//                 extreme lives in the L2-resident global table and is updated (read-check + RED) only by
//                 rows that pass the filter — O(log rows) per station per CTA on unordered data
//   sum           RED.ADD.F64 on the global table (shared-memory f64 add is a CAS loop, ATOMS.CAST.SPIN)
// One persistent CTA per SM, 12 B of shared memory per station.
template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) groupagg_kernel_v2(const i64* __restrict__ keys, const double* __restrict__ vals,
                                                              i64 n, i64 G, AggWs a, int l2_every) {
    extern __shared__ __align__(16) unsigned char agg_smem[];
    const bool sum_to_l2 = l2_every > 0 && ((threadIdx.x >> 5) % l2_every) == l2_every - 1;
    constexpr bool COUNT_IN_SMEM = VARIANT == 1 || VARIANT == 3 || VARIANT == 5;
    constexpr bool SUM_IN_SMEM = VARIANT == 3 || VARIANT == 5;
    // VARIANT 5: both filters in ONE 32-bit word per station — [31:16] top 16 bits of the running max's order encoding,
    // [15:0] of the running min's (sign, exponent, 4 mantissa bits: a row passes only if it lands in the extreme's
    // 1/16-binade bucket or beyond) — so a row that cannot beat either extreme costs one LDS instead of two; the word
    // changes O(log rows) times per station and CTA, by CAS
    constexpr bool PACKED = VARIANT == 5;
    // VARIANT 6 (kept for the record, NOT the default: 6.86 vs 3.73 ms at 1e9 rows): the same state as one 16-byte
    // RECORD per station {sum f64 | packed filters u32 | count u32}: one LDS.128 brings the filter word AND the current
    // sum (the seed of a hand-written CAS loop), instead of an LDS.32 + an LDS.64.  The software loop around
    // ATOMS.CAS.64 is far slower than the ATOMS.CAST.SPIN.64 the compiler emits for atomicAdd(double) on shared memory
    constexpr bool AOS = VARIANT == 6;
    struct __align__(16) Rec { u64 sum_bits; u32 filt; u32 cnt; };
    Rec* s_rec = reinterpret_cast<Rec*>(agg_smem);
    double* s_sum = reinterpret_cast<double*>(agg_smem);  // VARIANT 3/5 only (G doubles first, keeps 8-byte alignment)
    u32* s_cnt = reinterpret_cast<u32*>(agg_smem + (SUM_IN_SMEM ? (size_t)G * 8 : 0));
    u32* s_minhi = s_cnt + G;       // PACKED: the packed filter word
    u32* s_maxhi = s_minhi + G;     // (unused when PACKED)
    if (AOS) {
        for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
            s_rec[i].sum_bits = 0ull;
            s_rec[i].filt = ((u32)(a.tmax[i] >> 48) << 16) | (u32)(a.tmin[i] >> 48);
            s_rec[i].cnt = 0u;
        }
    }
    for (i64 i = threadIdx.x; !AOS && i < G; i += blockDim.x) {
        s_cnt[i] = 0u;
        if (PACKED) {
            s_minhi[i] = ((u32)(a.tmax[i] >> 48) << 16) | (u32)(a.tmin[i] >> 48);
        } else {
            s_minhi[i] = (u32)(a.tmin[i] >> 32);  // seeded extremes (0xffffffff / 0 where the sample had no row)
            s_maxhi[i] = (u32)(a.tmax[i] >> 32);
        }
        if (SUM_IN_SMEM) s_sum[i] = 0.0;
    }
    __syncthreads();
    auto row = [&](i64 k, double v) {
        if ((u64)k >= (u64)G) {
            *a.err = 1;
            return;
        }
        const u64 emin = enc_for_min(v), emax = enc_for_max(v);
        if (AOS) {
            Rec* p = &s_rec[k];
            const uint4 r = *reinterpret_cast<const uint4*>(p);  // {sum lo, sum hi, filters, count}
            const u32 lo = (u32)(emin >> 48), hi = (u32)(emax >> 48);
            const u32 f = r.z;
            if (lo <= (f & 0xffffu)) {
                for (u32 cur = f; lo < (cur & 0xffffu);) {
                    const u32 got = atomicCAS(&p->filt, cur, (cur & 0xffff0000u) | lo);
                    if (got == cur) break;
                    cur = got;
                }
                if (emin < ld_relaxed_u64(&a.tmin[k])) atomicMin(&a.tmin[k], emin);
            }
            if (hi >= (f >> 16)) {
                for (u32 cur = f; hi > (cur >> 16);) {
                    const u32 got = atomicCAS(&p->filt, cur, (cur & 0xffffu) | (hi << 16));
                    if (got == cur) break;
                    cur = got;
                }
                if (emax > ld_relaxed_u64(&a.tmax[k])) atomicMax(&a.tmax[k], emax);
            }
            unsigned long long old = (unsigned long long)r.x | ((unsigned long long)r.y << 32);
            while (true) {
                const unsigned long long nw = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)old) + v);
                const unsigned long long got = atomicCAS(reinterpret_cast<unsigned long long*>(&p->sum_bits), old, nw);
                if (got == old) break;
                old = got;
            }
            atomicAdd(&p->cnt, 1u);
            return;
        }
        if (PACKED) {
            const u32 lo = (u32)(emin >> 48), hi = (u32)(emax >> 48);
            const u32 f = s_minhi[k];
            // a row that passes this CTA's filter visits the global table anyway: the filter is then tightened to what
            // ALL CTAs have seen so far (the global extreme's bucket), not just to this row's — the private filters
            // converge after one or two visits per station instead of ~ln(rows per station per CTA)
            if (lo <= (f & 0xffffu)) {
                const u64 gmin = ld_relaxed_u64(&a.tmin[k]);
                if (emin < gmin) atomicMin(&a.tmin[k], emin);
                const u32 glo = (u32)(gmin >> 48), nlo = lo < glo ? lo : glo;
                for (u32 cur = f; nlo < (cur & 0xffffu);) {
                    const u32 got = atomicCAS(&s_minhi[k], cur, (cur & 0xffff0000u) | nlo);
                    if (got == cur) break;
                    cur = got;
                }
            }
            if (hi >= (f >> 16)) {
                const u64 gmax = ld_relaxed_u64(&a.tmax[k]);
                if (emax > gmax) atomicMax(&a.tmax[k], emax);
                const u32 ghi = (u32)(gmax >> 48), nhi = hi > ghi ? hi : ghi;
                for (u32 cur = f; nhi > (cur >> 16);) {
                    const u32 got = atomicCAS(&s_minhi[k], cur, (cur & 0xffffu) | (nhi << 16));
                    if (got == cur) break;
                    cur = got;
                }
            }
        } else {
            const u32 lo = (u32)(emin >> 32), hi = (u32)(emax >> 32);
            if (lo <= s_minhi[k]) {
                if (lo < s_minhi[k]) atomicMin(&s_minhi[k], lo);
                if (emin < ld_relaxed_u64(&a.tmin[k])) atomicMin(&a.tmin[k], emin);
            }
            if (hi >= s_maxhi[k]) {
                if (hi > s_maxhi[k]) atomicMax(&s_maxhi[k], hi);
                if (emax > ld_relaxed_u64(&a.tmax[k])) atomicMax(&a.tmax[k], emax);
            }
        }
        // the float64 add is the expensive step: a CAS loop (ATOMS.CAST.SPIN.64) on the SM's own atomic unit, or a
        // RED.ADD.F64 on the L2 atomic units.  Both units work at the same time when one warp in `l2_every` sends
        // its sums to L2 (groupagg_finalize adds the global table and the per-CTA partials).
        if (SUM_IN_SMEM && !sum_to_l2)
            atomicAdd(&s_sum[k], v);
        else
            atomicAdd(&a.tsum[k], v);
        if (COUNT_IN_SMEM)
            atomicAdd(&s_cnt[k], 1u);
        else
            atomicAdd(&a.tcnt[k], 1ull);
    };
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    constexpr int U = 4;
    if (vec) {
        const i64 n2 = n / 2;
        for (i64 g0 = tid; g0 < n2; g0 += stride * U) {
            longlong2 k[U];
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    k[u] = __ldcs(reinterpret_cast<const longlong2*>(keys) + g);
                    v[u] = __ldcs(reinterpret_cast<const double2*>(vals) + g);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    row(k[u].x, v[u].x);
                    row(k[u].y, v[u].y);
                }
            }
        }
        if ((n & 1) && tid == 0) row(keys[n - 1], vals[n - 1]);
    } else {
        for (i64 i = tid; i < n; i += stride) row(keys[i], vals[i]);
    }
    if (AOS) {
        __syncthreads();
        double* ps = a.psum + (size_t)blockIdx.x * G;
        u32* pc = a.pcnt + (size_t)blockIdx.x * G;
        for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
            ps[i] = __longlong_as_double((long long)s_rec[i].sum_bits);
            pc[i] = s_rec[i].cnt;
        }
    } else if (SUM_IN_SMEM) {
        // per-CTA partial tables leave as plain coalesced stores; groupagg_finalize folds them in CTA order
        // (3 M global atomics for the flush cost ~0.1 ms at 1e8 rows)
        __syncthreads();
        double* ps = a.psum + (size_t)blockIdx.x * G;
        u32* pc = a.pcnt + (size_t)blockIdx.x * G;
        for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
            ps[i] = s_sum[i];
            pc[i] = s_cnt[i];
        }
    } else if (COUNT_IN_SMEM) {
        __syncthreads();
        for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
            const u32 c = s_cnt[i];
            if (c) atomicAdd(&a.tcnt[i], (u64)c);
        }
    }
}

// v4 (kept for the record, not the default: ATOMS.CAS.128 measured 1.7x slower than v3's CAS.64 + POPC.INC):
// one 16-byte shared-memory record per station {sum f64 | count u32 | min filter u16 | max filter u16}, updated
// with ONE 128-bit CAS per row (ATOMS.CAS.128): the record load, the float64 add, the count and both filters share a
// single shared-memory read-modify-write instead of five separate operations.  The filters are the top 16 bits of the
// order-encoded value (sign, exponent, 4 mantissa bits); rows that land in the extreme's bucket take the exact
// read-check + RED on the L2-resident global table, like v2.  16 B of shared memory per station (G <= 12 800).
struct __align__(16) AggRec {
    u64 sum_bits;
    u64 meta;  // [31:0] count  [47:32] min filter  [63:48] max filter
};
__device__ __forceinline__ AggRec cas128_shared(AggRec* addr, AggRec cmp, AggRec val) {
    AggRec old;
    asm volatile(
        "{\n .reg .b128 c, v, o;\n mov.b128 c, {%2, %3};\n mov.b128 v, {%4, %5};\n"
        " atom.shared.cas.b128 o, [%6], c, v;\n mov.b128 {%0, %1}, o;\n}\n"
        : "=l"(old.sum_bits), "=l"(old.meta)
        : "l"(cmp.sum_bits), "l"(cmp.meta), "l"(val.sum_bits), "l"(val.meta), "r"((u32)__cvta_generic_to_shared(addr))
        : "memory");
    return old;
}

__global__ void __launch_bounds__(1024, 1) groupagg_kernel_v4(const i64* __restrict__ keys, const double* __restrict__ vals,
                                                              i64 n, i64 G, AggWs a) {
    extern __shared__ __align__(16) unsigned char agg_smem[];
    AggRec* tab = reinterpret_cast<AggRec*>(agg_smem);
    for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
        tab[i].sum_bits = 0ull;
        tab[i].meta = 0x0000ffff00000000ull;  // count 0, min filter 0xffff, max filter 0
    }
    __syncthreads();
    auto row = [&](i64 k, double v) {
        if ((u64)k >= (u64)G) {
            *a.err = 1;
            return;
        }
        const u64 emin = enc_for_min(v), emax = enc_for_max(v);
        const u32 lo16 = (u32)(emin >> 48), hi16 = (u32)(emax >> 48);
        AggRec* p = &tab[k];
        AggRec cur = *p;
        bool slow_min, slow_max;
        while (true) {
            const u32 mn = (u32)(cur.meta >> 32) & 0xffffu, mx = (u32)(cur.meta >> 48);
            slow_min = lo16 <= mn;
            slow_max = hi16 >= mx;
            AggRec nw;
            nw.sum_bits = (u64)__double_as_longlong(__longlong_as_double((i64)cur.sum_bits) + v);
            nw.meta = (u64)((u32)cur.meta + 1u) | ((u64)(lo16 < mn ? lo16 : mn) << 32) | ((u64)(hi16 > mx ? hi16 : mx) << 48);
            const AggRec got = cas128_shared(p, cur, nw);
            if (got.sum_bits == cur.sum_bits && got.meta == cur.meta) break;
            cur = got;
        }
        if (slow_min && emin < ld_relaxed_u64(&a.tmin[k])) atomicMin(&a.tmin[k], emin);
        if (slow_max && emax > ld_relaxed_u64(&a.tmax[k])) atomicMax(&a.tmax[k], emax);
    };
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
    constexpr int U = 4;
    if (vec) {
        const i64 n2 = n / 2;
        for (i64 g0 = tid; g0 < n2; g0 += stride * U) {
            longlong2 k[U];
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    k[u] = __ldcs(reinterpret_cast<const longlong2*>(keys) + g);
                    v[u] = __ldcs(reinterpret_cast<const double2*>(vals) + g);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const i64 g = g0 + u * stride;
                if (g < n2) {
                    row(k[u].x, v[u].x);
                    row(k[u].y, v[u].y);
                }
            }
        }
        if ((n & 1) && tid == 0) row(keys[n - 1], vals[n - 1]);
    } else {
        for (i64 i = tid; i < n; i += stride) row(keys[i], vals[i]);
    }
    __syncthreads();
    double* ps = a.psum + (size_t)blockIdx.x * G;
    u32* pc = a.pcnt + (size_t)blockIdx.x * G;
    for (i64 i = threadIdx.x; i < G; i += blockDim.x) {
        ps[i] = __longlong_as_double((i64)tab[i].sum_bits);
        pc[i] = (u32)tab[i].meta;
    }
}

__device__ __forceinline__ double nanpropagating_min(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
__device__ __forceinline__ double nanpropagating_max(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// Four lanes per station: lane q folds the per-CTA partial tables [q * nparts / 4, (q + 1) * nparts / 4) in CTA order, the
// four results meet in a fixed tree — bit-reproducible run to run, and four times the loads in flight of the round-1
// one-thread-per-station loop (26 -> 9 us for 148 tables of 10 k stations).
__global__ void __launch_bounds__(256) groupagg_finalize(AggWs a, i64 G, double* __restrict__ mn, double* __restrict__ mx,
                                                          double* __restrict__ sm, i64* __restrict__ ct, int accumulate, int nparts) {
    const int q = threadIdx.x & 3;
    const i64 g = (i64)blockIdx.x * 64 + (threadIdx.x >> 2);
    const bool live = g < G;
    double psum = 0.0;
    u64 c = 0;
    if (live) {
        const int per = (nparts + 3) / 4;
        const int p0 = q * per, p1 = (p0 + per < nparts) ? p0 + per : nparts;
        for (int part = p0; part < p1; part++) {  // fixed order within the quarter
            psum += a.psum[(size_t)part * G + g];
            c += a.pcnt[(size_t)part * G + g];
        }
    }
    // quarters 0+1 and 2+3, then the halves (the four lanes of a station are adjacent lanes of one warp)
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    if (!live || q != 0) return;
    c += a.tcnt[g];
    double vmin = __longlong_as_double(0x7ff0000000000000ll), vmax = -vmin, vsum = 0.0;
    if (c) {
        const u64 emin = a.tmin[g], emax = a.tmax[g];
        vmin = (emin == 0ull) ? __longlong_as_double(0x7ff8000000000000ll) : key_to_f64(emin);
        vmax = key_to_f64(emax);
        vsum = a.tsum[g] + psum;
    }
    if (accumulate) {
        mn[g] = nanpropagating_min(mn[g], vmin);
        mx[g] = nanpropagating_max(mx[g], vmax);
        sm[g] = sm[g] + vsum;
        ct[g] = ct[g] + (i64)c;
    } else {
        mn[g] = vmin;
        mx[g] = vmax;
        sm[g] = vsum;
        ct[g] = (i64)c;
    }
}

__global__ void __launch_bounds__(256) groupagg_merge_kernel(double* mn, double* mx, double* sm, i64* ct,
                                                              const double* __restrict__ mn2, const double* __restrict__ mx2,
                                                              const double* __restrict__ sm2, const i64* __restrict__ ct2, i64 G) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    mn[g] = nanpropagating_min(mn[g], mn2[g]);
    mx[g] = nanpropagating_max(mx[g], mx2[g]);
    sm[g] = sm[g] + sm2[g];
    ct[g] = ct[g] + ct2[g];
}

// ---- multi-GPU exchange (dist.sharded_groupagg).  A rank's four tables travel as ONE block of 8-byte lanes
// [4][G] = {min, max, sum (float64 bit patterns), count (int64)}.
// fold: `w` such blocks, in rank order (bit-identical on every rank and every run) -> the four result tables.
__global__ void __launch_bounds__(256) groupagg_fold_kernel(const u64* __restrict__ blocks, i64 w, i64 G, i64 stride, double* mn,
                                                             double* mx, double* sm, i64* ct) {
    const i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double a = __longlong_as_double((i64)blocks[g]), b = __longlong_as_double((i64)blocks[G + g]);
    double s = __longlong_as_double((i64)blocks[2 * G + g]);
    i64 c = (i64)blocks[3 * G + g];
    for (i64 r = 1; r < w; r++) {
        const u64* t = blocks + r * stride;
        a = nanpropagating_min(a, __longlong_as_double((i64)t[g]));
        b = nanpropagating_max(b, __longlong_as_double((i64)t[G + g]));
        s = s + __longlong_as_double((i64)t[2 * G + g]);
        c += (i64)t[3 * G + g];
    }
    mn[g] = a;
    mx[g] = b;
    sm[g] = s;
    ct[g] = c;
}
// hash partition: station s is owned by rank s mod w (the hash of a dense station id), slot s / w of the owner's slice.
// to_owners: local [4][G] block -> send buffer [w][4][B] (B = ceil(G / w); missing slots carry the fold's identities).
__global__ void __launch_bounds__(256) groupagg_to_owners_kernel(const u64* __restrict__ block, i64 G, i64 w, i64 B, u64* __restrict__ send) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;  // index into [w][B]
    if (i >= w * B) return;
    const i64 r = i / B, slot = i - r * B, st = slot * w + r;
    u64* dst = send + r * 4 * B + slot;
    if (st < G) {
        dst[0] = block[st];
        dst[B] = block[G + st];
        dst[2 * B] = block[2 * G + st];
        dst[3 * B] = block[3 * G + st];
    } else {
        dst[0] = 0x7ff0000000000000ull;      // +inf
        dst[B] = 0xfff0000000000000ull;      // -inf
        dst[2 * B] = 0;
        dst[3 * B] = 0;
    }
}
// from_owners: all-gathered owner slices [w][4][B] -> the four full tables in station order
__global__ void __launch_bounds__(256) groupagg_from_owners_kernel(const u64* __restrict__ full, i64 G, i64 w, i64 B, double* mn,
                                                                    double* mx, double* sm, i64* ct) {
    const i64 st = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= G) return;
    const i64 r = st % w, slot = st / w;
    const u64* src = full + r * 4 * B + slot;
    mn[st] = __longlong_as_double((i64)src[0]);
    mx[st] = __longlong_as_double((i64)src[B]);
    sm[st] = __longlong_as_double((i64)src[2 * B]);
    ct[st] = (i64)src[3 * B];
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_groupagg_fold(const void* blocks, int64_t n_blocks, int64_t n_groups, int64_t block_stride, double* min_out,
                                 double* max_out, double* sum_out, int64_t* count_out, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_blocks >= 1 && n_groups >= 0 && block_stride >= 4 * n_groups, "kgb_groupagg_fold: bad sizes");
    if (n_groups == 0) return KGB_OK;
    KGB_REQUIRE(blocks && min_out && max_out && sum_out && count_out, "kgb_groupagg_fold: null pointer");
    groupagg_fold_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const u64*)blocks, n_blocks, n_groups, block_stride, min_out, max_out, sum_out, (i64*)count_out);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

extern "C" int kgb_groupagg_to_owners(const void* block, int64_t n_groups, int64_t world, void* send, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_groups >= 0 && world >= 1, "kgb_groupagg_to_owners: bad sizes");
    if (n_groups == 0) return KGB_OK;
    KGB_REQUIRE(block && send, "kgb_groupagg_to_owners: null pointer");
    const i64 B = (n_groups + world - 1) / world;
    groupagg_to_owners_kernel<<<(unsigned)((world * B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const u64*)block, n_groups, world, B, (u64*)send);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

extern "C" int kgb_groupagg_from_owners(const void* full, int64_t n_groups, int64_t world, double* min_out, double* max_out,
                                        double* sum_out, int64_t* count_out, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_groups >= 0 && world >= 1, "kgb_groupagg_from_owners: bad sizes");
    if (n_groups == 0) return KGB_OK;
    KGB_REQUIRE(full && min_out && max_out && sum_out && count_out, "kgb_groupagg_from_owners: null pointer");
    const i64 B = (n_groups + world - 1) / world;
    groupagg_from_owners_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const u64*)full, n_groups, world, B, min_out, max_out, sum_out, (i64*)count_out);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}

extern "C" int kgb_groupagg_workspace_bytes(int64_t n, int64_t n_groups, size_t* out_bytes) {
    (void)n;
    KGB_REQUIRE(out_bytes && n_groups >= 0, "kgb_groupagg_workspace_bytes: bad arguments");
    *out_bytes = carve_agg(nullptr, n_groups).total;
    return KGB_OK;
}

static int groupagg_run(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups, double* min_out, double* max_out,
                        double* sum_out, int64_t* count_out, int32_t accumulate, void* workspace, size_t workspace_bytes,
                        void* stream, bool check_now);

extern "C" int kgb_groupagg_f64(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups, double* min_out,
                                double* max_out, double* sum_out, int64_t* count_out, int32_t accumulate,
                                void* workspace, size_t workspace_bytes, void* stream) {
    return groupagg_run(keys, vals, n, n_groups, min_out, max_out, sum_out, count_out, accumulate, workspace, workspace_bytes,
                        stream, true);
}

extern "C" int kgb_groupagg_f64_async(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups, double* min_out,
                                      double* max_out, double* sum_out, int64_t* count_out, int32_t accumulate,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    return groupagg_run(keys, vals, n, n_groups, min_out, max_out, sum_out, count_out, accumulate, workspace, workspace_bytes,
                        stream, false);
}

extern "C" int kgb_groupagg_status(const void* workspace, int64_t n_groups, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_groups >= 0, "kgb_groupagg_status: negative size");
    if (n_groups == 0) return KGB_OK;
    KGB_REQUIRE(workspace, "kgb_groupagg_status: null workspace");
    AggWs a = carve_agg(const_cast<void*>(workspace), n_groups);
    int err = 0;
    KGB_CUDA_CHECK(cudaMemcpyAsync(&err, a.err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    KGB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    KGB_REQUIRE(err == 0, "kgb_groupagg_f64: key outside [0, n_groups)");
    return KGB_OK;
}

static int groupagg_run(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups, double* min_out, double* max_out,
                        double* sum_out, int64_t* count_out, int32_t accumulate, void* workspace, size_t workspace_bytes,
                        void* stream, bool check_now) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n >= 0 && n_groups >= 0, "kgb_groupagg_f64: negative size");
    if (n_groups == 0) return KGB_OK;
    KGB_REQUIRE(min_out && max_out && sum_out && count_out && workspace &&
                    workspace_bytes >= carve_agg(nullptr, n_groups).total && (n == 0 || (keys && vals)),
                "kgb_groupagg_f64: null pointer or workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    AggWs a = carve_agg(workspace, n_groups);
    const size_t gb = align_up((size_t)n_groups * 8, 256);
    KGB_CUDA_CHECK(cudaMemsetAsync(a.tmin, 0xff, gb, st));
    KGB_CUDA_CHECK(cudaMemsetAsync(a.tmax, 0, 3 * gb + 256, st));
    int nparts = 0;
    if (n > 0) {
        long long blocks = (n / 2 + 256 * 4 - 1) / (256 * 4);
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        // variant: 3 = sum / count / filters in separate shared-memory arrays (default when 20 B/station fit),
        // 4 = one 128-bit CAS per row on a 16-byte record (ATOMS.CAS.128 turned out slower than CAS.64 + POPC.INC), 1 = filters + count in shared
        // memory and RED.ADD.F64 sums on the global table, 2 = filters only, 0 = v1 (all four updates on the
        // global table).  KGB200_AGG_VARIANT overrides for measurements.
        // 5 = variant 3 with both filters packed into one 32-bit word (one LDS per row instead of two; default).
        // Round 2, measured and NOT kept: count and filters in ONE 32-bit word {count:8 | max filter:12 | min filter:12}
        // updated by a native RETURNING ATOMS.ADD (the filters arrive with the increment: 3 shared-memory operations
        // per row instead of 4; the one thread that sees the 8-bit count wrap adds 256 to the global table) — 5.30 ms
        // at 1e9 rows against 3.73 ms: a returning shared atomic stalls its warp far longer than LDS + ATOMS.POPC.INC.
        // (64-bit shared integer adds are not native at all: ATOMS.CAST.SPIN.64, a CAS loop.)
        int variant = 5;  // measured at 1e9 rows / 10k stations: v5 3.72 ms (0.66 of roofline), v3 4.09 ms (0.60), v4 7.1 ms, v1 ~14 ms, v0 ~23 ms
        if (const char* e = getenv("KGB200_AGG_VARIANT")) variant = atoi(e);
        if (variant < 0 || variant > 6) variant = 5;
        const size_t limit = AGG_SMEM_LIMIT;
        if (variant >= 3 && !agg_has_parts(n_groups)) variant = 1;  // (12 800 stations per CTA; 10 240 for variant 3)
        if (variant == 3 && (size_t)n_groups * 20 > limit) variant = 1;
        if (variant != 0 && (size_t)n_groups * 12 > limit) variant = 0;
        if (variant == 0) {
            groupagg_kernel<<<(unsigned)blocks, 256, 0, st>>>((const i64*)keys, vals, n, n_groups, a);
        } else if (variant == 4) {
            static bool attr4[64] = {false};
            const int dev = current_device();
            if (!attr4[dev]) {
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                attr4[dev] = true;
            }
            long long pb = (n / 2 + 1024 * 4 - 1) / (1024 * 4);
            if (pb > sm_count()) pb = sm_count();
            if (pb > AGG_MAX_PARTS) pb = AGG_MAX_PARTS;
            if (pb < 1) pb = 1;
            groupagg_kernel_v4<<<(unsigned)pb, 1024, (size_t)n_groups * 16, st>>>((const i64*)keys, vals, n, n_groups, a);
            nparts = (int)pb;
        } else {
            const size_t smem = (size_t)n_groups * (variant == 3 ? 20 : variant >= 5 ? 16 : 12);
            static bool attr_set[64] = {false};
            const int dev = current_device();
            if (!attr_set[dev]) {
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v2<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                KGB_CUDA_CHECK(cudaFuncSetAttribute(groupagg_kernel_v2<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
                attr_set[dev] = true;
            }
            int l2_every = 0;
            if (const char* e = getenv("KGB200_AGG_L2_EVERY")) l2_every = atoi(e);
            long long pb = (n / 2 + 1024 * 4 - 1) / (1024 * 4);
            if (pb > sm_count()) pb = sm_count();
            if (pb > AGG_MAX_PARTS) pb = AGG_MAX_PARTS;
            if (pb < 1) pb = 1;
            // seed the filters from a sample: ~64 rows per station, and only when the run is long enough to pay for it
            long long seed_rows = 64 * n_groups;
            if (const char* e = getenv("KGB200_AGG_SEED")) seed_rows = atoll(e);
            if (seed_rows > 0 && n >= 32 * seed_rows) {
                const long long sb = std::min<long long>((seed_rows + 255) / 256, (long long)sm_count() * 8);
                groupagg_seed_kernel<<<(unsigned)sb, 256, 0, st>>>((const i64*)keys, vals, seed_rows, n_groups, a);
            }
            if (variant == 1)
                groupagg_kernel_v2<1><<<(unsigned)pb, 1024, smem, st>>>((const i64*)keys, vals, n, n_groups, a, l2_every);
            else if (variant == 2)
                groupagg_kernel_v2<2><<<(unsigned)pb, 1024, smem, st>>>((const i64*)keys, vals, n, n_groups, a, l2_every);
            else if (variant == 6) {
                groupagg_kernel_v2<6><<<(unsigned)pb, 1024, smem, st>>>((const i64*)keys, vals, n, n_groups, a, l2_every);
                nparts = (int)pb;
            } else if (variant == 5) {
                groupagg_kernel_v2<5><<<(unsigned)pb, 1024, smem, st>>>((const i64*)keys, vals, n, n_groups, a, l2_every);
                nparts = (int)pb;
            } else {
                groupagg_kernel_v2<3><<<(unsigned)pb, 1024, smem, st>>>((const i64*)keys, vals, n, n_groups, a, l2_every);
                nparts = (int)pb;
            }
        }
    }
    groupagg_finalize<<<(unsigned)((n_groups + 63) / 64), 256, 0, st>>>(a, n_groups, min_out, max_out, sum_out,
                                                                         (i64*)count_out, accumulate, nparts);
    KGB_CUDA_CHECK(cudaGetLastError());
    if (!check_now) return KGB_OK;  // kgb_groupagg_f64_async: the caller collects the flag with kgb_groupagg_status
    int err = 0;
    // out-of-range keys are a caller bug: report it (one 4-byte read after the kernels)
    KGB_CUDA_CHECK(cudaMemcpyAsync(&err, a.err, 4, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    KGB_REQUIRE(err == 0, "kgb_groupagg_f64: key outside [0, n_groups)");
    return KGB_OK;
}

extern "C" int kgb_groupagg_merge(double* min_io, double* max_io, double* sum_io, int64_t* count_io, const double* min_in,
                                  const double* max_in, const double* sum_in, const int64_t* count_in, int64_t n_groups,
                                  void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_groups >= 0, "kgb_groupagg_merge: negative size");
    if (n_groups == 0) return KGB_OK;
    groupagg_merge_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        min_io, max_io, sum_io, (i64*)count_io, min_in, max_in, sum_in, (const i64*)count_in, n_groups);
    KGB_CUDA_CHECK(cudaGetLastError());
    return KGB_OK;
}
