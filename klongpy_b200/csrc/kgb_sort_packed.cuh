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
#include <cstdlib>

#include "kgb_internal.h"

namespace kgb {
namespace pk {

constexpr int MAXP = 6;  // passes of one packed sort (<= 54 key bits at 9 bits per pass; the planner never needs more)
constexpr int MINBITS = 7, MAXBITS = 10;
constexpr int PLAN_MAXBITS = 10;  // widest digit the planner uses (KGB200_SORT_MAXBITS overrides; see make_plan)
constexpr int PLAN_MAXP = 5;  // passes a plan may use (n >= 2^14 leaves <= 50 key bits)
constexpr int MAXBINS = 1 << MAXBITS;
constexpr int ROLE_FIRST = 1, ROLE_LAST = 2;
constexpr u32 NO_TILE = 0xffffffffu;
constexpr int LB_PAD = 32;  // readable rows in front of tile 0's look-back words (the walk loads without bounds tests)

#ifndef KGB_PK_ATOM
#define KGB_PK_ATOM 0  // 1: warp digit counters are u32 updated with ATOMS.ADD; 0: u16 counters, leader LDS + STS
#endif
#ifndef KGB_PK_STATS
#define KGB_PK_STATS 0  // 1 (lab only): count look-back rounds / steps and the compute threads' waits into PassArgs::stats
#endif
#ifndef KGB_PK_NOLB
#define KGB_PK_NOLB 0
#endif
#ifndef KGB_PK_LB
#define KGB_PK_LB 16   // look-back states fetched per chain and round (capped so that a thread keeps <= 32 loads in flight)
#endif

struct PassArgs {
    const void* in_keys;    // FIRST: the caller's typed keys
    const u64* src;         // packed input of a later pass
    u64* dst;               // packed output (not LAST)
    i64* idx_out;           // LAST: the permutation
    void* keys_sorted_out;  // LAST, optional: typed sorted keys
    u64* enc_sorted_out;    // LAST, optional: order-encoded sorted keys
    const u32* hist;        // [1 << BITS] digit counts of this pass (the look-back group scans them into digit bases)
    u32* next_hist;         // next pass's digit counts, accumulated by THIS pass while the keys are in registers
    int next_shift;         //   (digit position of the next pass inside the packed word)
    u32 next_mask;          //   (1 << next bits) - 1, or 0: nothing to count (last pass, or the histograms came up front)
    const u64* invalid;     // device flag: non-zero = the guessed key range was wrong, every pass returns at once
    u32* states;            // [tiles][1 << BITS] look-back words of THIS pass (zeroed): [31:30] flag, [29:0] count
    u32* ticket;
    u32 n;
    int shift;              // bit position of this pass's digit inside the packed word
    int lo;                 // lowest live bit of the encoded key
    int pb;                 // position bits
    u64 kmask;              // mask of the live key field (after >> lo)
    u64 enc_base;           // smallest encoded key: the key field is (enc - enc_base) >> lo
    u32 stagger_ns;         // CTA b starts b * stagger_ns late (keeps the CTAs out of lockstep, see the look-back group)
    int dtype;              // KGB_* (+ _RAW) code of the caller's keys
    int descending;
    u64* stats;             // lab builds only (KGB_PK_STATS)
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
// look-back word of one (tile, digit): n < 2^30, so flag and count share 32 bits and a thread can keep 32 of them in flight
constexpr u32 LB_AGG = 1u << 30, LB_PREFIX = 2u << 30, LB_COUNT = (1u << 30) - 1u;
__device__ __forceinline__ u32 ld_relaxed_u32(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32* p, u32 v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int N> __device__ __forceinline__ void ld_relaxed_vec(const u32* p, u32 (&w)[N]) {
    static_assert(N == 1 || N == 2 || N == 4, "vector width");
    if (N == 1)
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
    else if (N == 2)
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[N > 1 ? 1 : 0]) : "l"(p) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0]), "=r"(w[N > 1 ? 1 : 0]), "=r"(w[N > 2 ? 2 : 0]), "=r"(w[N > 3 ? 3 : 0]) : "l"(p) : "memory");
}
template <int N> __device__ __forceinline__ void st_relaxed_vec(u32* p, const u32* w) {
    if (N == 1)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
    else if (N == 2)
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(w[0]), "r"(w[N > 1 ? 1 : 0]) : "memory");
    else
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w[0]), "r"(w[N > 1 ? 1 : 0]),
                     "r"(w[N > 2 ? 2 : 0]), "r"(w[N > 3 ? 3 : 0]) : "memory");
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
    static constexpr int GS = ELB < 4 ? ELB : 4;            // digits whose look-back words travel in one vector access
    static constexpr int NG = ELB / GS;
    static constexpr int LBN = (32 / ELB) < KGB_PK_LB ? (32 / ELB > 0 ? 32 / ELB : 1) : KGB_PK_LB;  // tiles per group and round
    static_assert(T <= 32768, "tile offsets are 16-bit");
    static_assert(ELB == 1 || ELB == 2 || ELB == 4 || ELB == 8, "look-back digits per thread");
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
    u32 lbw[32];            // look-back group: warp totals of the digit-count scan
    u32 nhist[MAXBINS];     // next pass's digit counts of this CTA
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
    if (a.invalid && *a.invalid) return;  // (uniform: every thread reads the same word before any barrier)

    for (int q = tid; q < MAXBINS; q += CT + LBT) sm.nhist[q] = 0;
    if (tid == 0) {
        // Out of lockstep: if all CTAs publish their tile counts at the same moment, tile t must walk back over ~grid/2
        // aggregates to find an inclusive prefix.  Staggered by tile_period / grid, predecessor t-k published k * stagger
        // earlier and a walk ends within a round or two.
        if (a.stagger_ns && blockIdx.x) {
            unsigned long long t0, now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            const unsigned long long wait = (unsigned long long)blockIdx.x * a.stagger_ns;
            do {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            } while (now - t0 < wait);
        }
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
        // Thread lt owns ELB ADJACENT digits, in groups of GS <= 4 whose words travel in one vector load / store.  A round
        // fetches the LBN nearest unvisited predecessor tiles of every unfinished group (independent loads: one L2 round
        // trip) and folds them front to back: a tile whose words are not all there yet ends the round (poll again), a
        // tile carrying inclusive prefixes ends the walk.  A look-back word is never 0 once written (flag 01 or 10 on
        // top), an aggregate is >= 2^30 and a prefix >= 2^31, so "all present" / "all prefixes" are one unsigned min / max
        // each; the words of a group are written together, and a reader that catches them half-way (some prefixes, some
        // aggregates) simply polls again.  The raw words are summed, flags included, and the flags' contribution
        // ((tiles folded + 1) << 30) comes off at the end — everything is arithmetic modulo 2^32 and every true count is
        // below 2^30.  The rows in FRONT of tile 0 are readable padding (never interpreted: tile 0 always carries
        // prefixes), so the loads need no bounds test.
        constexpr int GS = C::GS, NG = C::NG;
        const int lt = tid - CT;
        const int d0 = lt * ELB;             // first digit of this thread
        const bool owner = d0 < BINS;
        // exclusive scan of the pass's digit counts -> global digit bases (was a separate launch per sort)
        u32 dbase[ELB];
        {
            u32 tot = 0;
#pragma unroll
            for (int e = 0; e < ELB; e++) {
                const u32 c = owner ? a.hist[d0 + e] : 0u;
                dbase[e] = tot;
                tot += c;
            }
            u32 inc = tot;
            const int ll = lt & 31;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 v = __shfl_up_sync(0xffffffffu, inc, o);
                if (ll >= o) inc += v;
            }
            if (ll == 31) sm.lbw[lt >> 5] = inc;
            asm volatile("bar.sync 2, %0;" ::"n"(LBT) : "memory");
            u32 pre = inc - tot;
            for (int w = 0; w < (lt >> 5); w++) pre += sm.lbw[w];
#pragma unroll
            for (int e = 0; e < ELB; e++) dbase[e] += pre;
        }
#if KGB_PK_STATS
        unsigned long long st_rounds = 0, st_steps = 0, st_tiles = 0;
#endif
        for (u32 i = 0;; i++) {
            const u32 s = i & 1u;
            mbar_wait(&sm.cnt_bar[s], (i >> 1) & 1u, true);
            const u32 tile = sm.tile_of[s];
            if (tile == NO_TILE) break;
            u32 excl[ELB];
#pragma unroll
            for (int e = 0; e < ELB; e++) excl[e] = 0;
            if (owner && tile > 0) {
                const u32* sp[NG];   // group g's words of the nearest unvisited predecessor
                u32 adv[NG];
                bool go[NG];
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    sp[g] = a.states + (size_t)tile * BINS + d0 + g * GS - BINS;
                    adv[g] = 0;
                    go[g] = !KGB_PK_NOLB;
                }
                u32 spins = 0;
                for (;;) {
                    bool any = false;
#pragma unroll
                    for (int g = 0; g < NG; g++) any = any || go[g];
                    if (!any) break;
                    u32 w[NG][LBN][GS];
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        if (go[g]) {
#pragma unroll
                            for (int j = 0; j < LBN; j++) ld_relaxed_vec<GS>(sp[g] - (size_t)j * BINS, w[g][j]);
                        }
                    }
                    u32 progress = 0;
#pragma unroll
                    for (int g = 0; g < NG; g++) {
                        if (go[g]) {
                            bool walk = true;
                            u32 a_g = 0;
#pragma unroll
                            for (int j = 0; j < LBN; j++) {
                                u32 mn = w[g][j][0], mx = w[g][j][0];
#pragma unroll
                                for (int e = 1; e < GS; e++) {
                                    mn = min(mn, w[g][j][e]);
                                    mx = max(mx, w[g][j][e]);
                                }
                                const bool all_pf = mn >= LB_PREFIX;
                                const bool ready = mn >= LB_AGG && (all_pf || mx < LB_PREFIX);
                                const bool take = walk && ready;
#pragma unroll
                                for (int e = 0; e < GS; e++) excl[g * GS + e] += take ? w[g][j][e] : 0u;
                                a_g += take ? 1u : 0u;
                                if (take && all_pf) go[g] = false;
                                walk = take && !all_pf;
                            }
                            sp[g] -= (size_t)a_g * BINS;
                            adv[g] += a_g;
                            progress |= a_g;
                        }
                    }
#if KGB_PK_STATS
                    if (lt == 0) {
                        st_rounds++;
                    }
#endif
                    if (!progress) {
                        __nanosleep(100);
                        if (++spins > (1u << 24)) __trap();  // a predecessor never published: fail, do not hang
                    }
                }
#pragma unroll
                for (int g = 0; g < NG; g++) {
#pragma unroll
                    for (int e = 0; e < GS; e++) excl[g * GS + e] -= (adv[g] + 1u) << 30;
                }
#if KGB_PK_STATS
                if (lt == 0) st_steps += adv[0];
#endif
#if KGB_PK_NOLB
#pragma unroll
                for (int e = 0; e < ELB; e++) excl[e] = 0;
#endif
            }
#if KGB_PK_STATS
            if (lt == 0) st_tiles++;
#endif
            if (owner) {
                u32 pf[ELB];
#pragma unroll
                for (int e = 0; e < ELB; e++) {
                    const u32 cd = sm.cntd[s][d0 + e];
                    pf[e] = LB_PREFIX | ((excl[e] + (cd >> 16)) & LB_COUNT);
                    sm.gbase[s][d0 + e] = dbase[e] + excl[e] - (cd & 0xffffu);
                }
                if (tile > 0) {
#pragma unroll
                    for (int g = 0; g < NG; g++) st_relaxed_vec<GS>(&a.states[(size_t)tile * BINS + d0 + g * GS], &pf[g * GS]);
                }
            }
            mbar_arrive(&sm.gb_bar[s]);
        }
#if KGB_PK_STATS
        if (lt == 0 && a.stats) {
            atomicAdd((unsigned long long*)&a.stats[0], st_rounds);
            atomicAdd((unsigned long long*)&a.stats[1], st_steps);
            atomicAdd((unsigned long long*)&a.stats[2], st_tiles);
        }
#endif
        return;
    }

    // ================================================================ compute threads
    const u64 posmask = (a.pb >= 64) ? ~0ull : ((1ull << a.pb) - 1ull);
    u64 key[KPT];
    const int wslot = warp * (KPT * 32);
    const u32 lt_mask = (1u << lane) - 1u;
#if KGB_PK_STATS
    unsigned long long st_wait = 0;
#endif

    // request tile t's keys from HBM into `key` (raw for the FIRST pass; finished by `finish_load`)
    auto load_tile = [&](u32 t) {
        if (t >= ntiles) return;
        const u32 base = t * (u32)T;
        const bool full = (n - base) >= (u32)T;
        if (FIRST) {
            if (KW == 8) {
                const u64* p = reinterpret_cast<const u64*>(a.in_keys) + base + wslot + lane;
                if (full) {
#pragma unroll
                    for (int k = 0; k < KPT; k++) key[k] = __ldcs(p + k * 32);
                } else {
#pragma unroll
                    for (int k = 0; k < KPT; k++) key[k] = (base + wslot + k * 32 + lane < n) ? __ldcs(p + k * 32) : 0ull;
                }
            } else {
                const u32* p = reinterpret_cast<const u32*>(a.in_keys) + base + wslot + lane;
#pragma unroll
                for (int k = 0; k < KPT; k++) key[k] = (full || base + wslot + k * 32 + lane < n) ? (u64)__ldcs(p + k * 32) : 0ull;
            }
        } else {
            const u64* p = a.src + base + wslot + lane;
            if (full) {
#pragma unroll
                for (int k = 0; k < KPT; k++) key[k] = __ldcs(p + k * 32);
            } else {
#pragma unroll
                for (int k = 0; k < KPT; k++) key[k] = (base + wslot + k * 32 + lane < n) ? __ldcs(p + k * 32) : ~0ull;
            }
        }
    };
    auto finish_load = [&](u32 t) {
        if (FIRST) {
            const u32 base = t * (u32)T;
            const u64 ebase = a.enc_base, km = a.kmask;
            const int lo = a.lo, pb = a.pb;
            auto pack_all = [&](auto enc) {  // one instantiation per key type: the dtype test runs once per tile, not per key
#pragma unroll
                for (int k = 0; k < KPT; k++) {
                    const u32 g = base + wslot + k * 32 + lane;
                    const u64 e = enc(key[k]) - ebase;
                    key[k] = (g < n) ? ((((e >> lo) & km) << pb) | (u64)g) : ~0ull;
                }
            };
            switch (a.dtype) {
                case KGB_I64: pack_all([](u64 r) { return r ^ 0x8000000000000000ull; }); break;
                case KGB_F64: pack_all([](u64 r) { return KeyCodec<KGB_F64>::enc(__longlong_as_double((i64)r)); }); break;
                case KGB_F64_RAW: pack_all([](u64 r) { return KeyCodec<KGB_F64_RAW>::enc(__longlong_as_double((i64)r)); }); break;
                case KGB_I32: pack_all([](u64 r) { return (u64)((u32)r ^ 0x80000000u); }); break;
                case KGB_F32: pack_all([](u64 r) { return KeyCodec<KGB_F32>::enc(__int_as_float((int)(u32)r)); }); break;
                default: pack_all([](u64 r) { return KeyCodec<KGB_F32_RAW>::enc(__int_as_float((int)(u32)r)); }); break;
            }
        }
    };
    // tile parked in stage st -> global, as digit-contiguous runs.  Interior tiles and the plain roles (middle pass;
    // ascending permutation without key outputs) take loops without per-key tests.
    auto write_out = [&](u32 li, u32 t) {
        const u32 st = li & 1u;
        {
            const u32 bar = smem_u32(&sm.gb_bar[st]), parity = (li >> 1) & 1u;
            u32 done = 0, spins = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#if KGB_PK_STATS
                if (!done) st_wait++;
#endif
                if (!done && ++spins > (1u << 26)) __trap();  // the look-back group never answered: fail, do not hang
            }
        }
        const u32 base = t * (u32)T;
        const u32 tile_n = (n - base) < (u32)T ? (n - base) : (u32)T;
        const u64* stg = sm.stage[st];
        const u32* gb = sm.gbase[st];
        const bool plain = !LAST || (!a.descending && !a.keys_sorted_out && !a.enc_sorted_out);
        if (tile_n == (u32)T && plain) {
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const u32 p = j * CT + tid;
                const u64 v = stg[p];
                const u32 g = gb[(u32)(v >> shift) & (BINS - 1)] + p;
                if (LAST)
                    a.idx_out[g] = (i64)(v & posmask);
                else
                    a.dst[g] = v;
            }
        } else if (LAST && tile_n == (u32)T) {  // grade-down and/or sorted keys wanted
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const u32 p = j * CT + tid;
                const u64 v = stg[p];
                const u32 g = gb[(u32)(v >> shift) & (BINS - 1)] + p;
                const u32 op = a.descending ? (n - 1u - g) : g;
                a.idx_out[op] = (i64)(v & posmask);
                if (a.keys_sorted_out || a.enc_sorted_out) {
                    const u64 enc = (((v >> a.pb) & a.kmask) << a.lo) + a.enc_base;
                    if (a.enc_sorted_out) a.enc_sorted_out[op] = enc;
                    if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, enc);
                }
            }
        } else {  // the ragged last tile
#pragma unroll 1
            for (int j = 0; j < KPT; j++) {
                const u32 p = j * CT + tid;
                if (p < tile_n) {
                    const u64 v = stg[p];
                    const u32 g = gb[(u32)(v >> shift) & (BINS - 1)] + p;
                    if (LAST) {
                        const u32 op = a.descending ? (n - 1u - g) : g;
                        a.idx_out[op] = (i64)(v & posmask);
                        if (a.keys_sorted_out || a.enc_sorted_out) {
                            const u64 enc = (((v >> a.pb) & a.kmask) << a.lo) + a.enc_base;
                            if (a.enc_sorted_out) a.enc_sorted_out[op] = enc;
                            if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, enc);
                        }
                    } else {
                        a.dst[g] = v;
                    }
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
        // Row k: BITS ballots find the lanes with the same digit; EVERY lane then reads the warp's running count of its
        // digit and writes back count + (number of peers) — peers read and write the same values, so no leader election,
        // no branch and no shuffle is needed; the rank inside the tile's digit group follows after the totals step.
        u32 rank[KPT];
        if (a.next_mask) {
            // the NEXT pass's digit histogram, taken here while the keys sit in registers (one ATOMS.POPC.INC per key: lanes
            // that hit the same bin fold in hardware) instead of another read of all keys before the sort
            const u32 base = tile * (u32)T;
            if ((n - base) >= (u32)T) {
#pragma unroll
                for (int k = 0; k < KPT; k++) atomicAdd(&sm.nhist[(u32)(key[k] >> a.next_shift) & a.next_mask], 1u);
            } else {
#pragma unroll
                for (int k = 0; k < KPT; k++)
                    if (base + wslot + k * 32 + lane < n) atomicAdd(&sm.nhist[(u32)(key[k] >> a.next_shift) & a.next_mask], 1u);
            }
        }
        {
            wcount_t* wh = &sm.whist[warp][0];
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const u32 d = (u32)(key[k] >> shift) & (BINS - 1);
                u32 peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < BITS; b++) {
                    // test bit b, vote, keep the lanes that agree with this lane: peers &= ballot ^ (bit ? 0 : ~0)
                    asm("{\n\t.reg .pred p;\n\t.reg .b32 t, m, nb;\n\t"
                        "and.b32 t, %1, %2;\n\t"
                        "setp.ne.u32 p, t, 0;\n\t"
                        "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
                        "selp.b32 nb, 0, -1, p;\n\t"
                        "lop3.b32 %0, %0, m, nb, 0x60;\n\t}"
                        : "+r"(peers) : "r"(d), "r"(1u << b));
                }
                const u32 c = wh[d];
                rank[k] = c + (u32)__popc(peers & lt_mask);
                wh[d] = (wcount_t)(c + (u32)__popc(peers));
                __syncwarp();  // the next row reads what this row wrote (other lanes may hold the digit then)
            }
        }
        bar_named<CT>(1);  // B: warp histograms complete
        // ---- digit totals: exclusive offsets across warps, tile count, publish, exclusive digit offsets in the tile
        u32 cnt[E], inc = 0, tsum = 0;
        u32 wcnt[E == 1 ? W : 1];   // E == 1: the warps' counts stay in registers across the digit scan
        if (tid < NA) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int d = tid * E + e;
                u32 run = 0;
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const u32 c = sm.whist[w][d];
                    if (E == 1)
                        wcnt[w] = c;
                    else
                        sm.whist[w][d] = (wcount_t)run;
                    run += c;
                }
                cnt[e] = run;
                tsum += run;
                st_relaxed_u32(&a.states[(size_t)tile * BINS + d], (tile == 0 ? LB_PREFIX : LB_AGG) | run);
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
                if (E == 1) {
                    // slot of warp w's first key with this digit = digit offset in the tile + keys of lower warps: ONE table
                    // for the scatter (it used to read the digit offset and the warp offset separately)
                    u32 at = run;
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        sm.whist[w][tid] = (wcount_t)at;
                        at += wcnt[w];
                    }
                }
                run += cnt[e];
            }
        }
        bar_named<CT>(1);  // C: counts + digit offsets of the tile are in place
        if (tid == 0) {
            sm.tile_of[st] = tile;
            mbar_arrive(&sm.cnt_bar[st]);  // hand the tile to the look-back group
        }
        // ---- scatter into tile-sorted order
        {
            const wcount_t* wh = &sm.whist[warp][0];
            const u32* cd = sm.cntd[st];
            u64* stg = sm.stage[st];
#pragma unroll
            for (int k = 0; k < KPT; k++) {
                const u32 d = (u32)(key[k] >> shift) & (BINS - 1);
                if (E == 1)
                    stg[(u32)wh[d] + rank[k]] = key[k];
                else
                    stg[(cd[d] & 0xffffu) + wh[d] + rank[k]] = key[k];
            }
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
    if (a.next_mask) {
        for (u32 q = tid; q <= a.next_mask; q += CT) {
            const u32 c = sm.nhist[q];
            if (c) atomicAdd(&a.next_hist[q], c);
        }
    }
    if (tid == 0) {
        sm.tile_of[i & 1u] = NO_TILE;  // release the look-back group
        mbar_arrive(&sm.cnt_bar[i & 1u]);
#if KGB_PK_STATS
        if (a.stats) atomicAdd((unsigned long long*)&a.stats[3], st_wait);
#endif
    }
}

// ------------------------------------------------------------------------------------------- analysis + histograms
// Control block (u64 words, zeroed before each attempt):
//   [0] OR of enc ^ enc(key[0])   [1] ~min(enc)   [2] max(enc)  (exact, analyze_kernel)   [3] prefix fix-up overflow
//   [4] guessed plan invalid
//   [8..10] the same three statistics of the SAMPLE
constexpr int CTL_DIFF = 0, CTL_NMIN = 1, CTL_MAX = 2, CTL_FIXOVER = 3, CTL_INVALID = 4, CTL_SAMPLE = 8, CTL_WORDS = 16;
constexpr int FULL_MAX_BITS = 15;   // key fields up to 2^15 values get a full histogram in shared memory (128 KB)
constexpr u32 SAMPLE_KEYS = 1u << 16;

template <int KW> __device__ __forceinline__ u64 load_key(const void* keys, u64 i) {
    return KW == 8 ? __ldg(reinterpret_cast<const u64*>(keys) + i) : (u64)__ldg(reinterpret_cast<const u32*>(keys) + i);
}

// min / max / varying bits of a strided sample of the keys: enough to GUESS the key field (verified by hist_verify_kernel)
template <int KW>
__global__ void __launch_bounds__(1024) sample_kernel(const void* __restrict__ keys, u32 n, int dtype, u64* __restrict__ ctl) {
    const u64 first = enc_rt(dtype, load_key<KW>(keys, 0));
    const u32 m = n < SAMPLE_KEYS ? n : SAMPLE_KEYS;
    const u64 step = n / m;
    u64 diff = 0, mn = first, mx = first;
    for (u32 j = blockIdx.x * 1024 + threadIdx.x; j < m; j += gridDim.x * 1024) {
        const u64 e = enc_rt(dtype, load_key<KW>(keys, (u64)j * step));
        diff |= e ^ first;
        mn = e < mn ? e : mn;
        mx = e > mx ? e : mx;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        diff |= __shfl_xor_sync(0xffffffffu, diff, o);
        const u64 a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (diff) atomicOr((unsigned long long*)&ctl[CTL_SAMPLE + CTL_DIFF], diff);
        atomicMax((unsigned long long*)&ctl[CTL_SAMPLE + CTL_NMIN], ~mn);
        atomicMax((unsigned long long*)&ctl[CTL_SAMPLE + CTL_MAX], mx);
    }
}

struct HvArgs {
    const void* keys;
    u32 n;
    int dtype;
    u64 enc_base;   // key field = (enc - enc_base) >> lo, `live` bits wide
    int lo, live, prefix;
    int shift0;     // first pass's digit inside the key field (digit mode)
    u32 mask0;
    u64* ctl;
    u32* hist;      // digit mode: first pass's digit counts; full mode: the whole field's histogram [1 << live]
};

// ONE read of the keys: the check that every key fits the guessed field, and a histogram — of the first pass's digit, or
// (FULL, fields up to 2^15 values) of the whole key field, from which every pass's digit counts AND the group boundaries
// of `=a` follow without touching the keys again.  Shared-memory atomicAdd(.., 1) compiles to ATOMS.POPC.INC (lanes on
// one bin fold in hardware).  A key outside the field raises ctl[CTL_INVALID]; the caller then takes the exact
// statistics (analyze_kernel) and plans again.
template <int KW, bool FULL>
__global__ void __launch_bounds__(1024) hist_verify_kernel(const HvArgs h) {
    extern __shared__ u32 sh_hist[];
    const u32 bins = FULL ? (1u << h.live) : (h.mask0 + 1u);
    for (u32 q = threadIdx.x; q < bins; q += 1024) sh_hist[q] = 0;
    __syncthreads();
    const u64 lowmask = (h.prefix || h.lo == 0) ? 0ull : ((1ull << h.lo) - 1ull);
    // fits <=> enc >= base, field < 2^live, shared low bits equal the base's: one unsigned compare on rel = enc - base
    // (enc < base wraps to a huge rel) plus the low-bit test
    const u64 limit = (h.live + h.lo >= 64) ? ~0ull : ((1ull << (h.live + h.lo)) - 1ull);
    bool bad = false;
    constexpr int U = 8;
    auto sweep = [&](auto enc) {  // one instantiation per key type: the dtype is tested once per kernel, not per key
        for (u64 base = (u64)blockIdx.x * (1024 * U); base < h.n; base += (u64)gridDim.x * (1024 * U)) {
            u64 r[U];
            bool valid[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 i = base + u * 1024 + threadIdx.x;
                valid[u] = i < h.n;
                r[u] = valid[u] ? (KW == 8 ? __ldcs(reinterpret_cast<const u64*>(h.keys) + i) : (u64)__ldcs(reinterpret_cast<const u32*>(h.keys) + i)) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 e = enc(r[u]);
                const u64 rel = e - h.enc_base;
                const bool fits = e >= h.enc_base && rel <= limit && (rel & lowmask) == 0;
                const u64 f = rel >> h.lo;
                if (valid[u]) {
                    if (fits)
                        atomicAdd(&sh_hist[FULL ? (u32)f : ((u32)(f >> h.shift0) & h.mask0)], 1u);
                    else
                        bad = true;
                }
            }
        }
    };
    switch (h.dtype) {
        case KGB_I64: sweep([](u64 r) { return r ^ 0x8000000000000000ull; }); break;
        case KGB_F64: sweep([](u64 r) { return KeyCodec<KGB_F64>::enc(__longlong_as_double((i64)r)); }); break;
        case KGB_F64_RAW: sweep([](u64 r) { return KeyCodec<KGB_F64_RAW>::enc(__longlong_as_double((i64)r)); }); break;
        case KGB_I32: sweep([](u64 r) { return (u64)((u32)r ^ 0x80000000u); }); break;
        case KGB_F32: sweep([](u64 r) { return KeyCodec<KGB_F32>::enc(__int_as_float((int)(u32)r)); }); break;
        default: sweep([](u64 r) { return KeyCodec<KGB_F32_RAW>::enc(__int_as_float((int)(u32)r)); }); break;
    }
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr((unsigned long long*)&h.ctl[CTL_INVALID], 1ull);
    __syncthreads();
    for (u32 q = threadIdx.x; q < bins; q += 1024) {
        const u32 c = sh_hist[q];
        if (c) atomicAdd(&h.hist[q], c);
    }
}

// exact statistics of ALL keys (only after a wrong guess): ctl[CTL_DIFF] = OR of enc ^ enc(key[0]), ctl[CTL_NMIN] = ~min,
// ctl[CTL_MAX] = max (zero-initialised by the caller, hence the complement)
template <int KW>
__global__ void __launch_bounds__(256) analyze_kernel(const void* __restrict__ keys, u32 n, int dtype, u64* __restrict__ ctl) {
    const u64 first = enc_rt(dtype, load_key<KW>(keys, 0));
    u64 diff = 0, mn = first, mx = first;
    constexpr int U = 8;
    for (u64 base = (u64)blockIdx.x * (256 * U); base < n; base += (u64)gridDim.x * (256 * U)) {
        u64 r[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const u64 i = base + u * 256 + threadIdx.x;
            r[u] = load_key<KW>(keys, i < n ? i : 0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const u64 e = enc_rt(dtype, r[u]);
            diff |= e ^ first;
            mn = e < mn ? e : mn;
            mx = e > mx ? e : mx;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        diff |= __shfl_xor_sync(0xffffffffu, diff, o);
        const u64 a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (diff) atomicOr((unsigned long long*)&ctl[CTL_DIFF], diff);
        atomicMax((unsigned long long*)&ctl[CTL_NMIN], ~mn);
        atomicMax((unsigned long long*)&ctl[CTL_MAX], mx);
    }
}

struct DeriveArgs {
    const u32* full;   // [1 << live]
    int live, npass;
    int shift[MAXP];   // digit positions inside the key field
    int bits[MAXP];
    u32* hist;         // [npass][MAXBINS] out
    // group boundaries (optional): starts of the runs of equal keys in the sorted order, ascending key
    i64* starts_out;
    i64* ngroups_out;
    u32 n;
};

// From the full key-field histogram: every pass's digit counts, and — for `=a` — the group start offsets (exclusive
// prefix sums of the non-empty bins) and their number.  One CTA; at most 2^15 bins.
static __global__ void __launch_bounds__(1024) derive_kernel(const DeriveArgs d) {
    __shared__ u32 sh[MAXP * MAXBINS];
    __shared__ u32 wtot[32], wcnt[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q < d.npass * MAXBINS; q += 1024) sh[q] = 0;
    __syncthreads();
    const u32 bins = 1u << d.live;
    const u32 per = (bins + 1023) / 1024;          // consecutive bins per thread (<= 32)
    const u32 b0 = tid * per;
    u32 sum = 0, cnt = 0;
    for (u32 q = b0; q < b0 + per && q < bins; q++) {
        const u32 c = d.full[q];
        if (c) {
            for (int p = 0; p < d.npass; p++) atomicAdd(&sh[p * MAXBINS + ((q >> d.shift[p]) & ((1u << d.bits[p]) - 1u))], c);
            sum += c;
            cnt++;
        }
    }
    __syncthreads();
    for (int q = tid; q < d.npass * MAXBINS; q += 1024) d.hist[q] = sh[q];
    if (!d.starts_out) return;
    // block-wide exclusive scan of (sum, cnt) over the threads, then each thread walks its bins again
    u32 isum = sum, icnt = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, isum, o), b = __shfl_up_sync(0xffffffffu, icnt, o);
        if (lane >= o) {
            isum += a;
            icnt += b;
        }
    }
    if (lane == 31) {
        wtot[warp] = isum;
        wcnt[warp] = icnt;
    }
    __syncthreads();
    u32 ps = isum - sum, pc = icnt - cnt;
    for (int w = 0; w < warp; w++) {
        ps += wtot[w];
        pc += wcnt[w];
    }
    for (u32 q = b0; q < b0 + per && q < bins; q++) {
        const u32 c = d.full[q];
        if (c) {
            d.starts_out[pc++] = (i64)ps;
            ps += c;
        }
    }
    if (tid == 1023) {
        d.starts_out[pc] = (i64)d.n;
        *d.ngroups_out = (i64)pc;
    }
}

// no live bit at all (or n == 1): the stable permutation is the identity
static __global__ void __launch_bounds__(256) identity_kernel(const PassArgs a) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const u64 op = a.descending ? (a.n - 1 - i) : i;
        a.idx_out[op] = (i64)i;
        if (a.enc_sorted_out) a.enc_sorted_out[op] = a.enc_base;
        if (a.keys_sorted_out) dec_store_rt(a.dtype, a.keys_sorted_out, op, a.enc_base);
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
    const u64* invalid;   // the guessed plan was wrong: nothing to fix up
    u32 n;
    int pb, dtype, descending;
};
template <int KW>
__global__ void __launch_bounds__(256) fixup_kernel(const FixArgs a) {
    if (a.invalid && *a.invalid) return;
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
    int npass;            // 0: identity; < 0: not sortable by the packed pass (the caller uses the general one)
    int bits[MAXP];
    int shift[MAXP];      // inside the packed word
    int lo, pb, live;     // live = number of key bits carried
    int prefix;           // 1: the keys have more live bits than fit; bits below `lo` are ignored (fix-up needed)
    u64 kmask, enc_base;
};

inline int ceil_log2_u64(u64 x) {  // smallest b with 2^b >= x
    int b = 0;
    while (b < 64 && (1ull << b) < x) b++;
    return b;
}

// Key field (enc - mn) >> lo for keys in [mn, mx] whose bits below `lo` all equal mn's; digit widths: as few passes as
// 10-bit digits allow, widths evened out (27 live bits -> 9+9+9, 14 -> 7+7).
inline Plan plan_from_range(u64 mn, u64 mx, int lo, u64 n) {
    Plan p{};
    p.pb = ceil_log2_u64(n);
    if (p.pb < 1) p.pb = 1;
    p.enc_base = mn;
    const u64 range = mx - mn;
    const int hi = range ? 63 - __builtin_clzll(range) : 0;
    if (lo > hi) lo = hi;
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
    // Digit widths by measured cost (ms per pass and 1e8 keys on a B200, profiles/r02h_sort_widths.log: 7 bits 0.52,
    // 8 bits 0.63, 9 bits 0.66, 10 bits 0.88 — the 1024-bin pass pays for two digits per totals thread, twice the
    // look-back words and 32 KB of warp counters): the cheapest multiset of widths covering `live` bits, e.g. 27 -> 9+9+9,
    // 30 -> 9+7+7+7 (not 10+10+10), 14 -> 7+7, 37 -> 9+7+7+7+7.  KGB200_SORT_MAXBITS caps the widest digit (tuning).
    int maxbits = PLAN_MAXBITS;
    if (const char* ev = getenv("KGB200_SORT_MAXBITS")) maxbits = atoi(ev) < MINBITS ? MINBITS : (atoi(ev) > MAXBITS ? MAXBITS : atoi(ev));
    static const int cost[MAXBITS + 1] = {0, 0, 0, 0, 0, 0, 0, 52, 63, 66, 88};
    int best_cost = 1 << 30, best_np = -1, best_w[PLAN_MAXP] = {0};
    int w[PLAN_MAXP];
    for (int np = 1; np <= PLAN_MAXP; np++) {
        // non-increasing width sequences of length np (odometer)
        for (int i = 0; i < np; i++) w[i] = MINBITS;
        for (;;) {
            int sum = 0, c = 0;
            for (int i = 0; i < np; i++) {
                sum += w[i];
                c += cost[w[i]];
            }
            if (sum >= live && c < best_cost) {
                best_cost = c;
                best_np = np;
                for (int i = 0; i < np; i++) best_w[i] = w[i];
            }
            int i = np - 1;   // next non-increasing sequence
            while (i >= 0 && w[i] == (i == 0 ? maxbits : w[i - 1])) i--;
            if (i < 0) break;
            w[i]++;
            for (int j = i + 1; j < np; j++) w[j] = MINBITS;
        }
    }
    if (best_np < 0) {  // only tiny n (few position bits) with wide keys
        p.npass = -1;
        return p;
    }
    p.npass = best_np;
    int pos = 0;
    for (int i = 0; i < best_np; i++) {
        p.bits[i] = best_w[i];
        p.shift[i] = p.pb + pos;
        pos += best_w[i];
    }
    return p;
}

// exact statistics (diff = OR of enc ^ enc(key 0), extremes) -> plan; all keys equal -> identity
inline Plan make_plan(u64 diff, u64 mn, u64 mx, u64 n) {
    if (diff == 0 || mx == mn) {
        Plan p{};
        p.pb = ceil_log2_u64(n);
        if (p.pb < 1) p.pb = 1;
        p.enc_base = mn;
        return p;
    }
    return plan_from_range(mn, mx, __builtin_ctzll(diff), n);
}

// statistics of a 64 K-key SAMPLE -> a guess: the sampled range widened by 1/1024 of itself on either side (the
// extremes of uniform data lie range/65536 outside the sample's), aligned to the low bits the sampled keys share.
// hist_verify_kernel checks every key against it; a wrong guess is re-planned from the exact statistics.
inline Plan guess_plan(u64 sdiff, u64 smn, u64 smx, u64 n) {
    const int lo = sdiff ? __builtin_ctzll(sdiff) : 0;
    const u64 unit = 1ull << lo;
    const u64 r = smx - smn;
    u64 margin = r / 1024 + 1;
    margin = (margin + unit - 1) / unit * unit;   // a multiple of 2^lo: the base keeps the keys' shared low bits
    const u64 mn = smn >= margin ? smn - margin : (smn & (unit - 1));
    const u64 mx = smx <= ~0ull - margin ? smx + margin : smx + ((~0ull - smx) / unit * unit);
    return plan_from_range(mn, mx, lo, n);
}

// ------------------------------------------------------------------------------------------- host-side driver
// geometry of the pass (tuned on a B200, profiles/r02*_sort_lab_*.log)
#ifndef KGB_PK_CT
#define KGB_PK_CT 512
#define KGB_PK_KPT 16
#define KGB_PK_LBT 128
#define KGB_PK_OCC 1
#endif

struct Ws {
    u64* ctl;      // CTL_WORDS control words (see above)
    u32* hist;     // [MAXP][MAXBINS] digit counts per pass
    u32* tickets;  // [MAXP] (+ pad)
    u32* full;     // [1 << FULL_MAX_BITS] full key-field histogram
    size_t zero_bytes;  // ctl .. full are contiguous (zeroed before every attempt)
    u32* states;   // look-back words, one [LB_PAD + tiles][1 << bits] block per pass
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
    w.full = (u32*)(p + o);
    o += (size_t)4 << FULL_MAX_BITS;
    w.zero_bytes = o;
    w.states = (u32*)(p + o);
    o += align_up(((size_t)((n + T - 1) / T) + LB_PAD) * 4 * (PLAN_MAXP * MAXBINS), 256);
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

struct GroupOut {      // `=a`: filled when the key field is small enough for the full histogram (done = 1)
    i64* starts_out;   // [G + 1]
    i64* ngroups_out;
    int done;
};
struct LabInfo {       // development harness only
    cudaEvent_t ev[MAXP + 4];   // [0] start, [1] sample read back, [2] histograms done, [3 + i] pass i done, last: fix-up done
    Plan plan;
    int attempts;
    u64* stats;
};

#define KGB_PK_TRY(expr)                 \
    do {                                 \
        cudaError_t e_ = (expr);         \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

// One attempt with plan `p`: histograms (+ verification of the plan), passes, prefix fix-up.  No synchronisation.
template <int CT, int KPT, int LBT, int OCC>
inline cudaError_t packed_attempt(const void* keys, int dtype, u64 n, const Plan& p, int descending, i64* idx_out,
                                  void* keys_sorted_out, u64* enc_sorted_out, const Ws& w, u64* buf0, u64* buf1, int sms,
                                  cudaStream_t st, GroupOut* grp, LabInfo* lab) {
    const int kw = key_width(dtype);
    // the full key-field histogram (16 K random bins: no lane folding, 0.28 ms at 1e8 keys against 0.19 for one digit) pays
    // only when `=a` wants the group table out of it
    const bool full = grp != nullptr && !p.prefix && p.live <= FULL_MAX_BITS;
    KGB_PK_TRY(cudaMemsetAsync(w.ctl, 0, CTL_SAMPLE * 8, st));                         // keep the sample's words
    KGB_PK_TRY(cudaMemsetAsync(w.hist, 0, w.zero_bytes - ((char*)w.hist - (char*)w.ctl), st));
    const size_t tiles = (size_t)((n + (u64)(CT * KPT) - 1) / (u64)(CT * KPT));
    {   // look-back words of all passes: one memset
        size_t words = 0;
        for (int i = 0; i < p.npass; i++) words += (tiles + LB_PAD) << p.bits[i];
        KGB_PK_TRY(cudaMemsetAsync(w.states, 0, words * 4, st));
    }
    HvArgs h{};
    h.keys = keys;
    h.n = (u32)n;
    h.dtype = dtype;
    h.enc_base = p.enc_base;
    h.lo = p.lo;
    h.live = p.live;
    h.prefix = p.prefix;
    h.shift0 = p.shift[0] - p.pb;
    h.mask0 = (1u << p.bits[0]) - 1u;
    h.ctl = w.ctl;
    h.hist = full ? w.full : w.hist;
    {
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            KGB_PK_TRY(cudaFuncSetAttribute(hist_verify_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << FULL_MAX_BITS));
            KGB_PK_TRY(cudaFuncSetAttribute(hist_verify_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << FULL_MAX_BITS));
            attr_set[dev] = true;
        }
        long long blocks = (long long)((n + 8191) / 8192);
        // two resident CTAs per SM wherever their histograms fit (the whole-field one takes 4 << live bytes: 128 KB at
        // 15 bits): 0.274 -> 0.191 ms for 10 k keys in 1e8 rows; three are slower again (KGB200_HV_CTAS: measurements)
        static const int hv_ctas = getenv("KGB200_HV_CTAS") ? atoi(getenv("KGB200_HV_CTAS")) : 0;
        const long long cap = (long long)sms * (hv_ctas > 0 ? hv_ctas : ((full && p.live > 14) ? 1 : 2));
        if (blocks > cap) blocks = cap;
        const size_t smem = full ? ((size_t)4 << p.live) : ((size_t)4 << p.bits[0]);
        if (full) {
            if (kw == 8) hist_verify_kernel<8, true><<<(unsigned)blocks, 1024, smem, st>>>(h);
            else hist_verify_kernel<4, true><<<(unsigned)blocks, 1024, smem, st>>>(h);
        } else {
            if (kw == 8) hist_verify_kernel<8, false><<<(unsigned)blocks, 1024, smem, st>>>(h);
            else hist_verify_kernel<4, false><<<(unsigned)blocks, 1024, smem, st>>>(h);
        }
        KGB_PK_TRY(cudaGetLastError());
    }
    if (grp) grp->done = 0;
    if (full) {
        DeriveArgs d{};
        d.full = w.full;
        d.live = p.live;
        d.npass = p.npass;
        for (int i = 0; i < p.npass; i++) {
            d.shift[i] = p.shift[i] - p.pb;
            d.bits[i] = p.bits[i];
        }
        d.hist = w.hist;
        d.n = (u32)n;
        if (grp) {
            d.starts_out = grp->starts_out;
            d.ngroups_out = grp->ngroups_out;
            grp->done = 1;
        }
        derive_kernel<<<1, 1024, 0, st>>>(d);
        KGB_PK_TRY(cudaGetLastError());
    }
    if (lab) cudaEventRecord(lab->ev[2], st);
    PassArgs a{};
    a.in_keys = keys;
    a.idx_out = idx_out;
    a.keys_sorted_out = keys_sorted_out;
    a.enc_sorted_out = (grp && full) ? nullptr : enc_sorted_out;  // `=a` wanted the sorted keys only to find the boundaries
    a.n = (u32)n;
    a.lo = p.lo;
    a.pb = p.pb;
    a.kmask = p.kmask;
    a.enc_base = p.enc_base;
    a.dtype = dtype;
    a.descending = descending;
    a.invalid = w.ctl + CTL_INVALID;
    a.stats = lab ? lab->stats : nullptr;
    {
        const char* ev = getenv("KGB200_SORT_STAGGER_NS");  // tuning override (0 disables)
        a.stagger_ns = ev ? (u32)atoi(ev) : 30u;
    }
    u64* bufs[2] = {buf0, buf1};
    u32* states = w.states;
    for (int i = 0; i < p.npass; i++) {
        int role = 0;
        if (i == 0) role |= ROLE_FIRST;
        if (i == p.npass - 1 && !p.prefix) role |= ROLE_LAST;
        a.src = bufs[(i + 1) & 1];
        a.dst = bufs[i & 1];
        a.hist = w.hist + (size_t)i * MAXBINS;
        a.ticket = w.tickets + i;
        a.shift = p.shift[i];
        a.states = states + ((size_t)LB_PAD << p.bits[i]);
        states += (tiles + LB_PAD) << p.bits[i];
        if (!full && i + 1 < p.npass) {
            a.next_hist = w.hist + (size_t)(i + 1) * MAXBINS;
            a.next_shift = p.shift[i + 1];
            a.next_mask = (1u << p.bits[i + 1]) - 1u;
        } else {
            a.next_hist = nullptr;
            a.next_mask = 0;
        }
        KGB_PK_TRY((Launcher<CT, KPT, LBT, OCC>::go(p.bits[i], role, kw, a, sms, st)));
        if (lab) {
            cudaEventRecord(lab->ev[3 + i], st);
            if (a.stats) a.stats += 4;
        }
    }
    if (p.prefix) {
        FixArgs f{};
        f.sorted = bufs[(p.npass - 1) & 1];
        f.keys = keys;
        f.idx_out = idx_out;
        f.keys_sorted_out = keys_sorted_out;
        f.enc_sorted_out = enc_sorted_out;
        f.flags = w.ctl + CTL_FIXOVER;
        f.invalid = w.ctl + CTL_INVALID;
        f.n = (u32)n;
        f.pb = p.pb;
        f.dtype = dtype;
        f.descending = descending;
        if (kw == 8) fixup_kernel<8><<<sms * 8, 256, 0, st>>>(f);
        else fixup_kernel<4><<<sms * 8, 256, 0, st>>>(f);
        KGB_PK_TRY(cudaGetLastError());
    }
    return cudaSuccess;
}

// The whole packed sort.  Returns 1 when idx_out (and the optional outputs) are complete, 0 when the caller must run
// the general pass (keys too wide for this n, or a prefix tie too long for the fix-up), a negative cudaError otherwise.
// Synchronises `st` twice: after the 64 K-key sample (to guess the key field) and at the end (was the guess right?).
template <int CT, int KPT, int LBT, int OCC>
inline int packed_sort_run(const void* keys, int dtype, u64 n, int descending, i64* idx_out, void* keys_sorted_out,
                           u64* enc_sorted_out, const Ws& w, u64* buf0, u64* buf1, int sms, cudaStream_t st,
                           GroupOut* grp = nullptr, LabInfo* lab = nullptr) {
#define KGB_PK_RUN(expr)                        \
    do {                                        \
        cudaError_t e_ = (expr);                \
        if (e_ != cudaSuccess) return -(int)e_; \
    } while (0)
    const int kw = key_width(dtype);
    if (lab) cudaEventRecord(lab->ev[0], st);
    KGB_PK_RUN(cudaMemsetAsync(w.ctl, 0, CTL_WORDS * 8, st));
    {
        const unsigned sb = n >= SAMPLE_KEYS ? 16u : 1u;
        if (kw == 8) sample_kernel<8><<<sb, 1024, 0, st>>>(keys, (u32)n, dtype, w.ctl);
        else sample_kernel<4><<<sb, 1024, 0, st>>>(keys, (u32)n, dtype, w.ctl);
        KGB_PK_RUN(cudaGetLastError());
    }
    u64 hs[3];
    KGB_PK_RUN(cudaMemcpyAsync(hs, w.ctl + CTL_SAMPLE, sizeof hs, cudaMemcpyDeviceToHost, st));
    KGB_PK_RUN(cudaStreamSynchronize(st));
    if (lab) cudaEventRecord(lab->ev[1], st);
    Plan p = guess_plan(hs[CTL_DIFF], ~hs[CTL_NMIN], hs[CTL_MAX], n);
    if (grp) grp->done = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (lab) {
            lab->plan = p;
            lab->attempts = attempt + 1;
        }
        if (p.npass < 0) return 0;
        if (p.npass == 0) {  // (second attempt only: the exact statistics say every key is equal)
            PassArgs a{};
            a.idx_out = idx_out;
            a.keys_sorted_out = keys_sorted_out;
            a.enc_sorted_out = enc_sorted_out;
            a.n = (u32)n;
            a.enc_base = p.enc_base;
            a.dtype = dtype;
            a.descending = descending;
            long long ib = (long long)((n + 255) / 256);
            if (ib > (long long)sms * 8) ib = (long long)sms * 8;
            identity_kernel<<<(unsigned)ib, 256, 0, st>>>(a);
            KGB_PK_RUN(cudaGetLastError());
            if (grp) grp->done = 0;
            return 1;
        }
        KGB_PK_RUN((packed_attempt<CT, KPT, LBT, OCC>(keys, dtype, n, p, descending, idx_out, keys_sorted_out, enc_sorted_out, w,
                                                     buf0, buf1, sms, st, grp, lab)));
        if (lab) cudaEventRecord(lab->ev[3 + p.npass], st);
        u64 hc[5];
        KGB_PK_RUN(cudaMemcpyAsync(hc, w.ctl, sizeof hc, cudaMemcpyDeviceToHost, st));
        KGB_PK_RUN(cudaStreamSynchronize(st));
        if (!hc[CTL_INVALID]) return hc[CTL_FIXOVER] ? 0 : 1;
        // the sample misled (an outlier, a rare low bit): exact statistics from one more read of the keys, then once more
        KGB_PK_RUN(cudaMemsetAsync(w.ctl, 0, CTL_SAMPLE * 8, st));
        long long blocks = (long long)((n + 2047) / 2048);
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        if (kw == 8) analyze_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(keys, (u32)n, dtype, w.ctl);
        else analyze_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(keys, (u32)n, dtype, w.ctl);
        KGB_PK_RUN(cudaGetLastError());
        KGB_PK_RUN(cudaMemcpyAsync(hc, w.ctl, sizeof hc, cudaMemcpyDeviceToHost, st));
        KGB_PK_RUN(cudaStreamSynchronize(st));
        p = make_plan(hc[CTL_DIFF], ~hc[CTL_NMIN], hc[CTL_MAX], n);
    }
    return -(int)cudaErrorUnknown;  // unreachable: a plan made from the exact statistics cannot be invalid
#undef KGB_PK_RUN
}

}  // namespace pk
}  // namespace kgb
