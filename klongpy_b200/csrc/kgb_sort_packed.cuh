// Packed-payload LSD radix passes for grade / group (round 2).
//
// The general pass (kgb_sort.cu) carries 12 bytes per key — order-encoded u64 key + u32 position — through registers,
// shared memory and HBM in every pass, although the histogram pass already knows which key bits vary at all.  Here the
// LIVE key bits and the position travel in ONE 64-bit word
//
//        packed = ((enc(key) >> lo) & kmask) << pb  |  position            pb = ceil(log2 n)
//
// whenever live bits + pb <= 64 (a permutation of 1e8: 27 + 27; 10 k distinct keys in 1e9 rows: 14 + 30), so a pass
// moves 16 bytes per key instead of 24, a tile needs 8 instead of 12 bytes of shared memory per key and a thread a third
// fewer payload registers; the digit width is chosen per sort (7..10 bits) so that 27 live bits take three passes, not
// four.  Keys with more live bits than fit (full-range int64, float64) are sorted on their top 64-pb live bits and a
// fix-up kernel orders the (rare) runs that tie on that prefix by the remaining bits (kgb_sort.cu: prefix_fixup).
//
// One pass = one persistent CTA per SM:
//   * CT compute threads hold KPT packed keys each in registers (tile = CT*KPT keys).  Rank = warp multisplit (BITS
//     ballots folded with LOP3) + one leader update of the warp's digit counter per distinct digit and row; tile digit
//     totals and the exclusive digit offsets inside the tile by the first BINS/E threads; the tile is then scattered into
//     a shared-memory stage in digit order ("parked") and the NEXT tile's keys are requested from HBM into the freed
//     registers before the tile parked one iteration ago is written out as digit-contiguous runs.
//   * LBT look-back threads own the digits' chained decoupled look-back (LB predecessor states per round and chain) and
//     hand the global digit bases back through an mbarrier, one tile period off the compute threads' critical path.
//   * tiles come from an atomic ticket, so no co-residency is assumed; every wait traps instead of hanging.
//
// Replaces np.argsort behind kg_argsort (klongpy/writer.py:97-122, klongpy/backends/numpy_backend.py:75-80).
#pragma once
#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {
namespace pk {

constexpr int MAXP = 6;  // passes of one packed sort (<= 54 key bits at 9 bits per pass; the planner never needs more)
constexpr int MINBITS = 7, MAXBITS = 10;
constexpr int MAXBINS = 1 << MAXBITS;
constexpr int ROLE_FIRST = 1, ROLE_LAST = 2;
constexpr u32 NO_TILE = 0xffffffffu;

#ifndef KGB_PK_ATOM
#define KGB_PK_ATOM 0  // 1: warp digit counters are u32 updated with ATOMS.ADD; 0: u16 counters, leader LDS + STS
#endif
#ifndef KGB_PK_LB
#define KGB_PK_LB 4    // look-back states fetched per chain and round
#endif

struct PassArgs {
    const void* in_keys;    // FIRST: the caller's typed keys
    const u64* src;         // packed input of a later pass
    u64* dst;               // packed output (not LAST)
    i64* idx_out;           // LAST: the permutation
    void* keys_sorted_out;  // LAST, optional: typed sorted keys
    u64* enc_sorted_out;    // LAST, optional: order-encoded sorted keys
    const u32* digit_base;  // [1 << BITS] exclusive global digit offsets of this pass
    u64* states;            // [tiles][1 << BITS] look-back words
    u32* ticket;
    u32 n;
    int shift;              // bit position of this pass's digit inside the packed word
    int lo;                 // lowest live bit of the encoded key
    int pb;                 // position bits
    u64 kmask;              // mask of the live key field (after >> lo)
    u64 enc_const;          // the bits all keys agree on (reconstructs encoded keys in the LAST pass)
    u32 epoch;
    int dtype;              // KGB_* (+ _RAW) code of the caller's keys
    int descending;
};

// ------------------------------------------------------------------------------------------- small device helpers
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity, bool sleep) {
    const u32 a = smem_u32(bar);
    u32 done = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && sleep) __nanosleep(64);  // idle partners should not eat the issue slots of the working warps
        if (!done && (++spins & 1023u) == 0) {  // a partner that never arrives: fail after ~4 s, do not hang
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > 4000000000ull) __trap();
        }
    }
}
template <int N> __device__ __forceinline__ void bar_named(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}

// order-preserving encoding chosen at run time (the FIRST pass is templated on the key width only)
__device__ __forceinline__ u64 enc_rt(int dt, u64 raw) {
    switch (dt) {
        case KGB_I64: return raw ^ 0x8000000000000000ull;
        case KGB_F64: return KeyCodec<KGB_F64>::enc(__longlong_as_double((i64)raw));
        case KGB_F64_RAW: return KeyCodec<KGB_F64_RAW>::enc(__longlong_as_double((i64)raw));
        case KGB_I32: return (u64)((u32)raw ^ 0x80000000u);
        case KGB_F32: return KeyCodec<KGB_F32>::enc(__int_as_float((int)(u32)raw));
        default: return KeyCodec<KGB_F32_RAW>::enc(__int_as_float((int)(u32)raw));
    }
}
__device__ __forceinline__ void dec_store_rt(int dt, void* out, u64 at, u64 enc) {
    switch (dt) {
        case KGB_I64: ((i64*)out)[at] = key_to_i64(enc); break;
        case KGB_F64:
        case KGB_F64_RAW: ((double*)out)[at] = key_to_f64(enc); break;
        case KGB_I32: ((int*)out)[at] = key_to_i32(enc); break;
        default: ((float*)out)[at] = key_to_f32(enc); break;
    }
}

// lanes of the warp that hold the same BITS-bit digit
template <int BITS> __device__ __forceinline__ u32 digit_peers(u32 d) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const bool bit = (d >> b) & 1u;
        const u32 m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

// ------------------------------------------------------------------------------------------- the pass
template <int BITS, int CT, int KPT, int LBT> struct PassCfg {
    static constexpr int BINS = 1 << BITS;
    static constexpr int T = CT * KPT;                      // keys per tile
    static constexpr int W = CT / 32;                       // compute warps
    static constexpr int E = BINS > CT ? BINS / CT : 1;     // digits per totals thread
    static constexpr int NA = BINS / E;                     // threads that own digits in the totals step
    static constexpr int NAW = (NA + 31) / 32;
    static constexpr int ELB = BINS > LBT ? BINS / LBT : 1; // chains per look-back thread
    static constexpr int LBN = ELB >= 8 ? (KGB_PK_LB < 2 ? KGB_PK_LB : 2) : KGB_PK_LB;  // states per chain and round
    static_assert(T <= 32768, "tile offsets are 16-bit");
    static_assert(BINS % E == 0 && (BINS <= CT || BINS % CT == 0) && (BINS <= LBT || BINS % LBT == 0), "digit ownership");
};

#if KGB_PK_ATOM
typedef u32 wcount_t;
#else
typedef unsigned short wcount_t;
#endif

template <int BITS, int CT, int KPT, int LBT> struct PassSmem {
    typedef PassCfg<BITS, CT, KPT, LBT> C;
    u64 stage[2][C::T];
    wcount_t whist[C::W][C::BINS];
    u32 cntd[2][C::BINS];   // (tile count of the digit) << 16 | exclusive digit offset inside the tile
    u32 gbase[2][C::BINS];  // global position of tile-sorted slot p with digit d is gbase[d] + p   (mod 2^32)
    u32 wsum[32];
    u64 cnt_bar[2], gb_bar[2];
    u32 tile_of[2];
    u32 tile_slot[2];
};

template <int BITS, int CT, int KPT, int LBT, int ROLE, int KW, int OCC>
__global__ void __launch_bounds__(CT + LBT, OCC) pass_kernel(const PassArgs a) {
    typedef PassCfg<BITS, CT, KPT, LBT> C;
    typedef PassSmem<BITS, CT, KPT, LBT> S;
    constexpr int BINS = C::BINS, T = C::T, W = C::W, E = C::E, NA = C::NA, ELB = C::ELB, LBN = C::LBN;
    constexpr bool FIRST = (ROLE & ROLE_FIRST) != 0, LAST = (ROLE & ROLE_LAST) != 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S& sm = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 n = a.n;
    const u32 ntiles = (n + T - 1) / T;
    const int shift = a.shift;
    const u32 epoch = a.epoch;

    if (tid == 0) {
        mbar_init(&sm.cnt_bar[0], 1);
        mbar_init(&sm.cnt_bar[1], 1);
        mbar_init(&sm.gb_bar[0], LBT);
        mbar_init(&sm.gb_bar[1], LBT);
        sm.tile_slot[0] = atomicAdd(a.ticket, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int WROW = BINS * (int)sizeof(wcount_t) / 4;  // one warp's digit counters, in 32-bit words
    if (tid < CT) {
        for (int q = lane; q < WROW; q += 32) reinterpret_cast<u32*>(&sm.whist[warp][0])[q] = 0;
    }
    __syncthreads();

    if (tid >= CT) {
        // ============================================================ look-back group
        const int lt = tid - CT;
        u32 dbase[ELB];
#pragma unroll
        for (int e = 0; e < ELB; e++) {
            const int d = lt + e * LBT;
            dbase[e] = d < BINS ? a.digit_base[d] : 0u;
        }
        for (u32 i = 0;; i++) {
            const u32 s = i & 1u;
            mbar_wait(&sm.cnt_bar[s], (i >> 1) & 1u, true);
            const u32 tile = sm.tile_of[s];
            if (tile == NO_TILE) break;
            u32 excl[ELB], back[ELB];  // running exclusive prefix; number of predecessors already folded in
            bool done[ELB];
#pragma unroll
            for (int e = 0; e < ELB; e++) {
                excl[e] = 0;
                back[e] = 0;
                done[e] = (tile == 0) || (lt + e * LBT >= BINS);
            }
            u32 spins = 0;
            for (;;) {
                bool all = true;
#pragma unroll
                for (int e = 0; e < ELB; e++) all = all && done[e];
                if (all) break;
                u64 w[ELB][LBN];
#pragma unroll
                for (int e = 0; e < ELB; e++) {
                    const int d = lt + e * LBT;
#pragma unroll
                    for (int j = 0; j < LBN; j++) {
                        const u32 bk = back[e] + (u32)j;  // predecessor tile - 1 - bk
                        w[e][j] = (!done[e] && bk < tile) ? ld_relaxed_u64(&a.states[(size_t)(tile - 1u - bk) * BINS + d])
                                                          : lb_pack(KGB_LB_PREFIX, epoch, 0);
                    }
                }
                bool progress = false;
#pragma unroll
                for (int e = 0; e < ELB; e++) {
                    if (!done[e]) {
                        bool stop = false;
#pragma unroll
                        for (int j = 0; j < LBN; j++) {
                            if (!stop && !done[e]) {
                                const u64 f = lb_flag(w[e][j], epoch);
                                if (f == 0) {
                                    stop = true;  // not published yet: retry from here next round
                                } else {
                                    excl[e] += (u32)lb_count(w[e][j]);
                                    back[e] += 1u;
                                    progress = true;
                                    if (f == KGB_LB_PREFIX || back[e] >= tile) done[e] = true;
                                }
                            }
                        }
                    }
                }
                if (!progress && ++spins > (1u << 24)) __trap();  // a predecessor never published: fail, do not hang
            }
#pragma unroll
            for (int e = 0; e < ELB; e++) {
                const int d = lt + e * LBT;
                if (d < BINS) {
                    const u32 cd = sm.cntd[s][d];
                    if (tile > 0) st_relaxed_u64(&a.states[(size_t)tile * BINS + d], lb_pack(KGB_LB_PREFIX, epoch, (u64)excl[e] + (cd >> 16)));
                    sm.gbase[s][d] = dbase[e] + excl[e] - (cd & 0xffffu);
                }
            }
            mbar_arrive(&sm.gb_bar[s]);
        }
        return;
    }

    // ================================================================ compute threads
    const u64 posmask = (a.pb >= 64) ? ~0ull : ((1ull << a.pb) - 1ull);
    u64 key[KPT];
    const int wslot = warp * (KPT * 32);

    // request tile t's keys from HBM into `key` (raw for the FIRST pass; finished by `finish_load`)
    auto load_tile = [&](u32 t) {
        if (t >= ntiles) return;
        const u32 base = t * (u32)T;
        const bool full = (n - base) >= (u32)T;
        if (FIRST) {
            if (KW == 8) {
                const u64* p = reinterpret_cast<const u64*>(a.in_keys) + base + wslot + lane;
#pragma unroll
                for (int k = 0; k < KPT; k++) key[k] = (full || base + wslot + k * 32 + lane < n) ? __ldcs(p + k * 32) : 0ull;
            } else {
                const u32* p = reinterpret_cast<const u32*>(a.in_keys) + base + wslot + lane;
#pragma unroll
                for (int k = 0; k < KPT; k++) key[k] = (full || base + wslot + k * 32 + lane < n) ? (u64)__ldcs(p + k * 32) : 0ull;
            }
        } else {
            const u64* p = a.src + base + wslot + lane;
#pragma unroll
            for (int k = 0; k < KPT; k++) key[k] = (full || base + wslot + k * 32 + lane < n) ? __ldcs(p + k * 32) : ~0ull;
        }
    };
    auto finish_load = [&](u32 t) {
        if (FIRST) {
            const u32 base = t * (u32)T;
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const u32 g = base + wslot + k * 32 + lane;
                const u64 e = enc_rt(a.dtype, key[k]);
                key[k] = (g < n) ? ((((e >> a.lo) & a.kmask) << a.pb) | (u64)g) : ~0ull;
            }
        }
    };
    // tile parked in stage st -> global, as digit-contiguous runs
    auto write_out = [&](u32 li, u32 t) {
        const u32 st = li & 1u;
        mbar_wait(&sm.gb_bar[st], (li >> 1) & 1u, false);
        const u32 base = t * (u32)T;
        const u32 tile_n = (n - base) < (u32)T ? (n - base) : (u32)T;
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const u32 p = j * CT + tid;
            if (p < tile_n) {
                const u64 v = sm.stage[st][p];
                const u32 d = (u32)(v >> shift) & (BINS - 1);
                const u32 g = sm.gbase[st][d] + p;
                if (LAST) {
                    const u32 op = a.descending ? (n - 1u - g) : g;
                    a.idx_out[op] = (i64)(v & posmask);
                    if (a.keys_sorted_out || a.enc_sorted_out) {
                        const u64 enc = (((v >> a.pb) & a.kmask) << a.lo) | a.enc_const;
                        if (a.enc_sorted_out) a.enc_sorted_out[op] = enc;
                        if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, enc);
                    }
                } else {
                    a.dst[g] = v;
                }
            }
        }
    };

    u32 tile = sm.tile_slot[0];
    u32 prev = NO_TILE;
    load_tile(tile);
    u32 i = 0;
    for (;; i++) {
        if (tile >= ntiles) break;
        const u32 st = i & 1u;
        if (tid == 0) sm.tile_slot[(i + 1) & 1u] = atomicAdd(a.ticket, 1u);
        finish_load(tile);
        // ---- rank: warp w owns slots [w*32*KPT, (w+1)*32*KPT), key k of lane l is slot w*32*KPT + k*32 + l.
        // Three sweeps so that the ballots, the counter updates and the shuffles of a thread are each issued back to back.
        u32 rank[KPT];
        {
            const u32 lt = (1u << lane) - 1u;
            u32 peers[KPT];
#pragma unroll
            for (int k = 0; k < KPT; k++) peers[k] = digit_peers<BITS>((u32)(key[k] >> shift) & (BINS - 1));
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const u32 d = (u32)(key[k] >> shift) & (BINS - 1);
                const bool leader = (peers[k] & lt) == 0u;
                rank[k] = 0;
#if KGB_PK_ATOM
                if (leader) rank[k] = atomicAdd(&sm.whist[warp][d], (u32)__popc(peers[k]));
#else
                if (leader) {
                    const u32 c = sm.whist[warp][d];
                    rank[k] = c;
                    sm.whist[warp][d] = (wcount_t)(c + (u32)__popc(peers[k]));
                }
                __syncwarp();  // the next row's leader of the same digit may be another lane
#endif
            }
#pragma unroll
            for (int k = 0; k < KPT; k++)
                rank[k] = __shfl_sync(0xffffffffu, rank[k], __ffs(peers[k]) - 1) + __popc(peers[k] & lt);
        }
        bar_named<CT>(1);  // B: warp histograms complete
        // ---- digit totals: exclusive offsets across warps, tile count, publish, exclusive digit offsets in the tile
        u32 cnt[E], inc = 0, tsum = 0;
        if (tid < NA) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int d = tid * E + e;
                u32 run = 0;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const u32 c = sm.whist[w][d];
                    sm.whist[w][d] = (wcount_t)run;
                    run += c;
                }
                cnt[e] = run;
                tsum += run;
                st_relaxed_u64(&a.states[(size_t)tile * BINS + d], lb_pack(tile == 0 ? KGB_LB_PREFIX : KGB_LB_AGG, epoch, run));
            }
            inc = tsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) sm.wsum[warp] = inc;
        }
        bar_named<CT>(1);  // C'
        if (tid < NA) {
            u32 run = inc - tsum;
#pragma unroll
            for (int w = 0; w < C::NAW; w++)
                if (w < warp) run += sm.wsum[w];
#pragma unroll
            for (int e = 0; e < E; e++) {
                sm.cntd[st][tid * E + e] = (cnt[e] << 16) | run;
                run += cnt[e];
            }
        }
        bar_named<CT>(1);  // C: counts + digit offsets of the tile are in place
        if (tid == 0) {
            sm.tile_of[st] = tile;
            mbar_arrive(&sm.cnt_bar[st]);  // hand the tile to the look-back group
        }
        // ---- scatter into tile-sorted order
#pragma unroll
        for (int k = 0; k < KPT; k++) {
            const u32 d = (u32)(key[k] >> shift) & (BINS - 1);
            const u32 p = (sm.cntd[st][d] & 0xffffu) + sm.whist[warp][d] + rank[k];
            sm.stage[st][p] = key[k];
        }
        bar_named<CT>(1);  // D: the sorted tile is parked; whist is free again
#pragma unroll
        for (int q = lane; q < WROW; q += 32) reinterpret_cast<u32*>(&sm.whist[warp][0])[q] = 0;
        const u32 next = sm.tile_slot[(i + 1) & 1u];
        load_tile(next);  // lands while the parked tile leaves
        if (prev != NO_TILE) write_out(i - 1, prev);
        prev = tile;
        tile = next;
    }
    if (prev != NO_TILE) write_out(i - 1, prev);
    bar_named<CT>(1);
    if (tid == 0) {
        sm.tile_of[i & 1u] = NO_TILE;  // release the look-back group
        mbar_arrive(&sm.cnt_bar[i & 1u]);
    }
}

// ------------------------------------------------------------------------------------------- analysis + histograms
// OR of enc(key) ^ enc(key[0]) over all keys: the bits that vary at all.  out[0] = diff mask, out[1] = enc(key[0]).
template <int KW>
__global__ void __launch_bounds__(256) analyze_kernel(const void* __restrict__ keys, u32 n, int dtype, u64* __restrict__ out) {
    const u64 first = enc_rt(dtype, KW == 8 ? reinterpret_cast<const u64*>(keys)[0] : (u64)reinterpret_cast<const u32*>(keys)[0]);
    u64 diff = 0;
    constexpr int U = 8;
    for (u64 base = (u64)blockIdx.x * (256 * U); base < n; base += (u64)gridDim.x * (256 * U)) {
        u64 r[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const u64 i = base + u * 256 + threadIdx.x;
            const u64 ii = i < n ? i : 0;
            r[u] = KW == 8 ? __ldg(reinterpret_cast<const u64*>(keys) + ii) : (u64)__ldg(reinterpret_cast<const u32*>(keys) + ii);
        }
#pragma unroll
        for (int u = 0; u < U; u++) diff |= enc_rt(dtype, r[u]) ^ first;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, o);
    if ((threadIdx.x & 31) == 0 && diff) atomicOr((unsigned long long*)&out[0], diff);
    if (blockIdx.x == 0 && threadIdx.x == 0) out[1] = first;
}

struct HistArgs {
    int npass;
    int shift[MAXP];  // digit position inside the KEY FIELD (packed shift minus pb)
    int bits[MAXP];
    int lo;
    u64 kmask;
    int dtype;
};

// digit histograms of all passes in one read of the keys: hist[p][1 << bits[p]] at hist + p * MAXBINS
template <int KW>
__global__ void __launch_bounds__(256) hist_kernel(const void* __restrict__ keys, u32 n, const HistArgs h, u32* __restrict__ ghist) {
    __shared__ u32 sh[MAXP * MAXBINS];
    const int tot = h.npass * MAXBINS;
    for (int i = threadIdx.x; i < tot; i += 256) sh[i] = 0;
    __syncthreads();
    constexpr int U = 4;
    for (u64 base = (u64)blockIdx.x * (256 * U); base < n; base += (u64)gridDim.x * (256 * U)) {
        u64 r[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const u64 i = base + u * 256 + threadIdx.x;
            valid[u] = i < n;
            const u64 ii = valid[u] ? i : 0;
            r[u] = KW == 8 ? __ldcs(reinterpret_cast<const u64*>(keys) + ii) : (u64)__ldcs(reinterpret_cast<const u32*>(keys) + ii);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (valid[u]) {
                const u64 f = (enc_rt(h.dtype, r[u]) >> h.lo) & h.kmask;
                // atomicAdd(.., 1) compiles to ATOMS.POPC.INC, which folds lanes that hit the same bin in hardware
                for (int p = 0; p < h.npass; p++) atomicAdd(&sh[p * MAXBINS + ((u32)(f >> h.shift[p]) & ((1u << h.bits[p]) - 1u))], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tot; i += 256) {
        const u32 c = sh[i];
        if (c) atomicAdd(&ghist[i], c);
    }
}

// exclusive scan of every pass's histogram (in place): one CTA per pass, MAXBINS threads
__global__ void __launch_bounds__(MAXBINS) base_kernel(u32* __restrict__ ghist) {
    __shared__ u32 ws[32];
    u32* h = ghist + (size_t)blockIdx.x * MAXBINS;
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;
    const u32 c = h[d];
    u32 inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    u32 pre = 0;
    for (int w = 0; w < warp; w++) pre += ws[w];
    h[d] = pre + inc - c;
}

// no live bit at all (or n == 1): the stable permutation is the identity
__global__ void __launch_bounds__(256) identity_kernel(const PassArgs a) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const u64 op = a.descending ? (a.n - 1 - i) : i;
        a.idx_out[op] = (i64)i;
        if (a.enc_sorted_out) a.enc_sorted_out[op] = a.enc_const;
        if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, a.enc_const);
    }
}

// ------------------------------------------------------------------------------------------- prefix fix-up
// Keys with more live bits than the packed word can carry were sorted on their top bits only.  `sorted` holds the packed
// words in (prefix, position) order; elements whose prefix is unique are final.  The head of a run of equal prefixes
// fetches the run's full keys and orders them by (key, position) with a stable insertion sort; a run longer than
// FIX_LIMIT raises flags[0] and the caller re-sorts with the general (12 bytes per key) pass.
constexpr int FIX_LIMIT = 32;
struct FixArgs {
    const u64* sorted;
    const void* keys;
    i64* idx_out;
    void* keys_sorted_out;
    u64* enc_sorted_out;
    u64* flags;
    u32 n;
    int pb, dtype, descending;
};
template <int KW>
__global__ void __launch_bounds__(256) fixup_kernel(const FixArgs a) {
    const u64 posmask = (1ull << a.pb) - 1ull;
    const bool want_keys = a.keys_sorted_out || a.enc_sorted_out;
    auto fetch = [&](u64 pos) -> u64 {
        return enc_rt(a.dtype, KW == 8 ? __ldg(reinterpret_cast<const u64*>(a.keys) + pos)
                                       : (u64)__ldg(reinterpret_cast<const u32*>(a.keys) + pos));
    };
    auto emit = [&](u64 i, u64 pos, u64 enc) {
        const u64 op = a.descending ? ((u64)a.n - 1 - i) : i;
        a.idx_out[op] = (i64)pos;
        if (a.enc_sorted_out) a.enc_sorted_out[op] = enc;
        if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, enc);
    };
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const u64 v = a.sorted[i];
        const u64 pf = v >> a.pb;
        const bool same_l = i > 0 && (a.sorted[i - 1] >> a.pb) == pf;
        const bool same_r = i + 1 < a.n && (a.sorted[i + 1] >> a.pb) == pf;
        if (!same_l && !same_r) {
            const u64 pos = v & posmask;
            emit(i, pos, want_keys ? fetch(pos) : 0ull);
        } else if (!same_l) {
            u64 e[FIX_LIMIT];
            u32 ps[FIX_LIMIT];
            int m = 0;
            bool over = false;
            for (;;) {
                if (i + m >= a.n) break;
                const u64 vv = a.sorted[i + m];
                if ((vv >> a.pb) != pf) break;
                if (m == FIX_LIMIT) {
                    over = true;
                    break;
                }
                ps[m] = (u32)(vv & posmask);
                m++;
            }
            if (over) {
                if (a.flags[0] == 0) atomicOr((unsigned long long*)&a.flags[0], 1ull);
                continue;
            }
            for (int j = 0; j < m; j++) e[j] = fetch(ps[j]);
            for (int j = 1; j < m; j++) {  // stable: positions ascend within the run
                const u64 ke = e[j];
                const u32 kp = ps[j];
                int q = j - 1;
                while (q >= 0 && e[q] > ke) {
                    e[q + 1] = e[q];
                    ps[q + 1] = ps[q];
                    q--;
                }
                e[q + 1] = ke;
                ps[q + 1] = kp;
            }
            for (int j = 0; j < m; j++) emit(i + j, ps[j], e[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------- host-side planning
struct Plan {
    int npass;            // 0: identity
    int bits[MAXP];
    int shift[MAXP];      // inside the packed word
    int lo, pb, live;     // live = number of key bits carried
    int prefix;           // 1: the keys have more live bits than fit; bits below `lo` are ignored (fix-up needed)
    u64 kmask, enc_const;
};

inline int ceil_log2_u64(u64 x) {  // smallest b with 2^b >= x
    int b = 0;
    while (b < 64 && (1ull << b) < x) b++;
    return b;
}

// diff: OR of enc ^ enc0.  Chooses the key field [lo, lo+live) and the digit widths.
inline Plan make_plan(u64 diff, u64 enc0, u64 n) {
    Plan p{};
    p.pb = ceil_log2_u64(n);
    if (p.pb < 1) p.pb = 1;
    p.enc_const = enc0;
    if (diff == 0) {
        p.npass = 0;
        return p;
    }
    const int hi = 63 - __builtin_clzll(diff);
    int lo = __builtin_ctzll(diff);
    int live = hi - lo + 1;
    const int room = 64 - p.pb;
    if (live > room) {
        p.prefix = 1;
        live = room;
        lo = hi + 1 - live;
    }
    p.lo = lo;
    p.live = live;
    p.kmask = live >= 64 ? ~0ull : ((1ull << live) - 1ull);
    p.enc_const = enc0 & ~(p.kmask << lo);
    int np = (live + MAXBITS - 1) / MAXBITS;
    if (np < 1) np = 1;
    p.npass = np;
    int left = live, pos = 0;
    for (int i = 0; i < np; i++) {
        int w = (left + (np - i) - 1) / (np - i);
        if (w < MINBITS) w = MINBITS;
        p.bits[i] = w;
        p.shift[i] = p.pb + pos;
        pos += w;
        left -= w;
        if (left < 0) left = 0;
    }
    return p;
}


// ------------------------------------------------------------------------------------------- host-side driver
// geometry of the pass (tuned on a B200, profiles/r02_sort_lab_*.log)
#ifndef KGB_PK_CT
#define KGB_PK_CT 512
#define KGB_PK_KPT 16
#define KGB_PK_LBT 128
#define KGB_PK_OCC 1
#endif

struct Ws {
    u64* ctl;      // [0] diff, [1] enc0, [2] fix-up flags
    u32* hist;     // [MAXP][MAXBINS]
    u32* tickets;  // [MAXP] (+ pad)
    u64* states;   // [tiles][MAXBINS]
    size_t zero_bytes;  // ctl .. tickets are contiguous (zeroed before the analysis)
    size_t total;
};
template <int T> inline Ws carve_ws(void* ws, u64 n) {
    Ws w;
    char* p = (char*)ws;
    size_t o = 0;
    w.ctl = (u64*)(p + o);
    o += 256;
    w.hist = (u32*)(p + o);
    o += (size_t)MAXP * MAXBINS * 4;
    w.tickets = (u32*)(p + o);
    o += 256;
    w.zero_bytes = o;
    w.states = (u64*)(p + o);
    o += align_up((size_t)((n + T - 1) / T) * MAXBINS * 8, 256);
    w.total = o;
    return w;
}

template <int CT, int KPT, int LBT, int OCC> struct Launcher {
    template <int BITS, int ROLE, int KW> static cudaError_t go1(const PassArgs& a, int sms, cudaStream_t st) {
        typedef PassSmem<BITS, CT, KPT, LBT> S;
        auto kern = pass_kernel<BITS, CT, KPT, LBT, ROLE, KW, OCC>;
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S));
            if (e != cudaSuccess) return e;
            attr_set[dev] = true;
        }
        const u32 ntiles = (a.n + (u32)(CT * KPT) - 1) / (u32)(CT * KPT);
        const u32 cap = (u32)sms * OCC;
        kern<<<ntiles < cap ? ntiles : cap, CT + LBT, sizeof(S), st>>>(a);
        return cudaGetLastError();
    }
    template <int BITS> static cudaError_t go_bits(int role, int kw, const PassArgs& a, int sms, cudaStream_t st) {
        switch (role) {
            case 0: return go1<BITS, 0, 8>(a, sms, st);
            case ROLE_LAST: return go1<BITS, ROLE_LAST, 8>(a, sms, st);
            case ROLE_FIRST: return kw == 8 ? go1<BITS, ROLE_FIRST, 8>(a, sms, st) : go1<BITS, ROLE_FIRST, 4>(a, sms, st);
            default: return kw == 8 ? go1<BITS, ROLE_FIRST | ROLE_LAST, 8>(a, sms, st) : go1<BITS, ROLE_FIRST | ROLE_LAST, 4>(a, sms, st);
        }
    }
    static cudaError_t go(int bits, int role, int kw, const PassArgs& a, int sms, cudaStream_t st) {
        switch (bits) {
            case 7: return go_bits<7>(role, kw, a, sms, st);
            case 8: return go_bits<8>(role, kw, a, sms, st);
            case 9: return go_bits<9>(role, kw, a, sms, st);
            case 10: return go_bits<10>(role, kw, a, sms, st);
        }
        return cudaErrorInvalidValue;
    }
};

inline int key_width(int dtype) { return (dtype == KGB_F32 || dtype == KGB_I32 || dtype == KGB_F32_RAW) ? 4 : 8; }

// Step 1: zero the control block and find the varying bits (the caller synchronises and reads ctl[0..1] back).
template <int T> inline cudaError_t packed_analyze(const void* keys, int dtype, u64 n, const Ws& w, int sms, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(w.ctl, 0, w.zero_bytes, st);
    if (e != cudaSuccess) return e;
    long long blocks = (long long)((n + 2047) / 2048);
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    if (key_width(dtype) == 8)
        analyze_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(keys, (u32)n, dtype, w.ctl);
    else
        analyze_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(keys, (u32)n, dtype, w.ctl);
    return cudaGetLastError();
}

// Step 2: histograms, digit bases and the passes of plan `p`.  buf0/buf1: 8n bytes each.  When `final_packed` is set the
// last pass is run as a middle pass and leaves the sorted packed words in *final_packed (the prefix fix-up reads them).
template <int CT, int KPT, int LBT, int OCC>
inline cudaError_t packed_passes(const void* keys, int dtype, u64 n, const Plan& p, int descending, i64* idx_out,
                                 void* keys_sorted_out, u64* enc_sorted_out, const Ws& w, u64* buf0, u64* buf1,
                                 u64** final_packed, int sms, cudaStream_t st, cudaEvent_t* lab_events = nullptr) {
    PassArgs a{};
    a.in_keys = keys;
    a.idx_out = idx_out;
    a.keys_sorted_out = keys_sorted_out;
    a.enc_sorted_out = enc_sorted_out;
    a.states = w.states;
    a.n = (u32)n;
    a.lo = p.lo;
    a.pb = p.pb;
    a.kmask = p.kmask;
    a.enc_const = p.enc_const;
    a.dtype = dtype;
    a.descending = descending;
    if (p.npass == 0) {
        long long ib = (long long)((n + 255) / 256);
        if (ib > (long long)sms * 8) ib = (long long)sms * 8;
        identity_kernel<<<(unsigned)ib, 256, 0, st>>>(a);
        return cudaGetLastError();
    }
    HistArgs h{};
    h.npass = p.npass;
    for (int i = 0; i < p.npass; i++) {
        h.shift[i] = p.shift[i] - p.pb;
        h.bits[i] = p.bits[i];
    }
    h.lo = p.lo;
    h.kmask = p.kmask;
    h.dtype = dtype;
    long long hb = (long long)((n + 1023) / 1024);
    if (hb > (long long)sms * 8) hb = (long long)sms * 8;
    if (key_width(dtype) == 8)
        hist_kernel<8><<<(unsigned)hb, 256, 0, st>>>(keys, (u32)n, h, w.hist);
    else
        hist_kernel<4><<<(unsigned)hb, 256, 0, st>>>(keys, (u32)n, h, w.hist);
    base_kernel<<<p.npass, MAXBINS, 0, st>>>(w.hist);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    {   // look-back words: one zeroing serves every pass (the epoch in each word tells the passes apart)
        int maxbits = 0;
        for (int i = 0; i < p.npass; i++) maxbits = p.bits[i] > maxbits ? p.bits[i] : maxbits;
        const size_t tiles = (size_t)((n + (u64)(CT * KPT) - 1) / (u64)(CT * KPT));
        e = cudaMemsetAsync(w.states, 0, tiles * ((size_t)8 << maxbits), st);
        if (e != cudaSuccess) return e;
    }
    if (lab_events) cudaEventRecord(lab_events[0], st);
    u64* bufs[2] = {buf0, buf1};
    for (int i = 0; i < p.npass; i++) {
        int role = 0;
        if (i == 0) role |= ROLE_FIRST;
        if (i == p.npass - 1 && !final_packed) role |= ROLE_LAST;
        a.src = bufs[(i + 1) & 1];
        a.dst = bufs[i & 1];
        a.digit_base = w.hist + (size_t)i * MAXBINS;
        a.ticket = w.tickets + i;
        a.shift = p.shift[i];
        a.epoch = (u32)i + 1;
        e = Launcher<CT, KPT, LBT, OCC>::go(p.bits[i], role, key_width(dtype), a, sms, st);
        if (e != cudaSuccess) return e;
        if (lab_events) cudaEventRecord(lab_events[i + 1], st);
    }
    if (final_packed) *final_packed = bufs[(p.npass - 1) & 1];
    return cudaSuccess;
}

}  // namespace pk
}  // namespace kgb
