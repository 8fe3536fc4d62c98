// 1brc text ingest: `name;[-]d[d].d\n` lines -> (dense station id, float64 temperature) columns + station dictionary.
// Replaces the per-line Python loop of the reference's examples/1brc/par.py:57-88 and the string grouping of
// worker.kg:5.  Byte work, HBM-bound by design; round 2 reads the text ONCE:
//
//   brc_onepass_kernel  one CTA per 65536-byte tile, tiles handed out by an atomic ticket.  The tile (plus a 512-byte
//                     halo in front, for the line that straddles the tile edge) is staged in shared memory by one
//                     cp.async.bulk (TMA engine, mbarrier completion); of every 16-byte granule the newline flags are taken
//                     (exact zero-byte test of x ^ 0x0a0a0a0a, one multiply gathers four flags: 24 instructions per
//                     granule) and paired into 32-byte mask words in shared memory — no mask array in HBM.  A block
//                     scan of the popcounts numbers the tile's lines and expands the masks into a list of newline
//                     positions; the tile's line count is published in a decoupled look-back chain (one 64-bit word
//                     per tile, kgb_device.cuh) by an otherwise idle warp, which returns the tile's first ROW NUMBER
//                     while the other warps expand.  Then line i goes to thread i mod T, so a warp reads ~5
//                     consecutive 128-byte lines of text and writes its 16 bytes per row coalesced.  Per line: the
//                     measurement is parsed backwards from the newline in fixed point; the name's first 24 bytes are
//                     loaded unconditionally (four 64-bit shared loads, funnel-shifted, masked by clamped shifts),
//                     hashed multilinearly (one IMAD.WIDE per word) and resolved in an L2-resident open-addressing
//                     table of 32-byte slots {tag | published | id | len, 24 name bytes}: a known name of <= 24
//                     bytes costs ONE 256-bit load and six compares.  Longer names continue against the dictionary
//                     row of their id.  A new name takes the slot with one CAS, draws the next dense id, stores its
//                     bytes and publishes.  Traffic: the text once + 16 B per row.
//   brc_count_kernel  newline count of a byte range (the exact row count for callers that want exactly sized outputs,
//                     or of three sample windows for an estimate: kgb_brc_estimate).
//
//   History: v1 (one thread per 128-byte segment, byte-wise FNV-1a and compare) 18.8 ms per 1e8 rows; round 1
//   (mask pass + parse pass, 64-bit multiply hash, 16-byte slots + dictionary compare) 2.03 ms; this file: see
//   DESIGN.md §3.7.
#include <cstdlib>

#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

constexpr int BRC_LOG2_SLOTS = 20;  // 32 MB; only the sectors of live names are ever hot (10k names: 320 KB)
constexpr u64 BRC_SLOTS = 1ull << BRC_LOG2_SLOTS;
constexpr int BRC_NAME_BYTES = 256;                 // dictionary row: a name of <= 255 bytes as zero-extended 32-bit words
constexpr int BRC_NAME_WORDS = BRC_NAME_BYTES / 4;
#ifndef KGB_BRC_MASK_LUT
#define KGB_BRC_MASK_LUT 0  // 1: name byte masks from a 25-row shared table (two LDS per line) instead of clamped shifts
#endif
#ifndef KGB_BRC_BULK
#define KGB_BRC_BULK 1  // 0: stage interior tiles through registers (LDG + STS) instead of cp.async.bulk
#endif
constexpr int BRC_HALO = 512;                       // bytes before the tile kept with it: a line may start there
constexpr int BRC_MIN_TILE = 16384;                 // the look-back states are sized for the smallest tile
template <int TILE>
struct BrcGeom {  // shared-memory geometry of one tile
    static constexpr int STAGE = BRC_HALO + TILE + 64;  // + misalignment of the tile (< 16) + 28 bytes of over-read
    static constexpr int GRAN = STAGE / 16;             // 16-byte granules staged per tile
    static constexpr int WORDS = STAGE / 32;            // mask words (one per 32 staged bytes)
    static constexpr int MAX_LINES = TILE / 6 + 1;      // the shortest valid line, "a;0.0\n", is 6 bytes
    static constexpr int POS_SLOTS = (MAX_LINES + 7) & ~7;
    static constexpr int OFF_MLUT = STAGE;                               // 25 rows x 8 words: byte masks of a name of len 0..24
    static constexpr int OFF_MASK = OFF_MLUT + 25 * 32;                  // newline mask words of the staged bytes
    static constexpr int OFF_POS = OFF_MASK + ((WORDS + 15) & ~15) * 4;  // newline positions (u16, relative to the tile)
    static constexpr int OFF_WCNT = OFF_POS + 16 + POS_SLOTS * 2;        // 16: s_pos[-1] is addressable (never used)
    static constexpr int OFF_MISC = OFF_WCNT + 32 * 4;
    static constexpr int SMEM = OFF_MISC + 64;
    static_assert(STAGE % 32 == 0, "mask words cover whole 32-byte groups");
    static_assert(OFF_MLUT % 16 == 0 && OFF_MASK % 16 == 0 && OFF_POS % 16 == 0 && OFF_WCNT % 8 == 0, "alignment");
};
constexpr int BRC_INSLOT = 24;                      // name bytes kept in the slot
constexpr int BRC_NO_START = -(1 << 30);
constexpr u64 BRC_MIX = 0xd6e8feb86659fd93ull;

// slot word 0: [63:29] tag (35 hash bits, never 0)  [28] published  [27:8] dense id  [7:0] name length
constexpr u64 BRC_PUB = 1ull << 28;
struct __align__(32) BrcSlot {  // one 32-byte sector: a lookup is ONE 256-bit load
    u64 w0;          // 0 = free; tag set by the claimant's CAS; id / len / published set by its second store
    u64 n0, n1, n2;  // the name's first 24 bytes (zero-extended): names of <= 24 bytes never touch the dictionary rows
};
struct Slot4 {
    u64 w0, n0, n1, n2;
};
__device__ __forceinline__ Slot4 ld_slot(const BrcSlot* p) {
    Slot4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.w0), "=l"(v.n0), "=l"(v.n1), "=l"(v.n2)
                 : "l"(p)
                 : "memory");
    return v;
}
struct BrcCtl {
    unsigned long long n_names;
    unsigned long long rows_total;   // written by the last tile
    unsigned long long sample_count; // brc_count_kernel's accumulator
    unsigned int ticket;
    int err;  // 1 malformed row, 2 dictionary full / name too long, 3 a slot was never published, 4 no final newline
};

// workspace: [BrcCtl 256 B] [look-back states, one u64 per tile] [table]
struct BrcWs {
    BrcCtl* ctl;
    u64* states;
    BrcSlot* tab;
    size_t zero_bytes;  // ctl + states + table are zeroed before a pass
    size_t total;
};
static i64 brc_tiles(i64 n, int tile) { return (n + tile - 1) / tile; }
static BrcWs carve_brc(void* ws, i64 n) {
    BrcWs w;
    char* p = (char*)ws;
    size_t o = 0;
    w.ctl = (BrcCtl*)(p + o);
    o += 256;
    w.states = (u64*)(p + o);
    o += align_up((size_t)(brc_tiles(n, BRC_MIN_TILE) + 1) * 8, 256);
    w.tab = (BrcSlot*)(p + o);
    o += (size_t)BRC_SLOTS * sizeof(BrcSlot);
    w.zero_bytes = o;
    w.total = o;
    return w;
}

// 4 bytes -> their newline flags as the TOP nibble of the result (bit 28 + i = byte i is '\n'); five instructions.
// Exact zero-byte test of x ^ 0x0a0a0a0a (no borrow between bytes), then one multiply gathers the four 0x80 flags:
// flag i sits at bit 8i + 7 and the multiplier's term 2^(21 - 7i) moves it to bit 28 + i; every other partial product
// lands on a distinct bit below 28 or overflows, so nothing carries into the nibble.
__device__ __forceinline__ u32 nl4_top(u32 x) {
    const u32 t = x ^ 0x0a0a0a0au;
    const u32 m = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);
    return m * 0x00204081u;
}
// 16 bytes -> 16 mask bits
__device__ __forceinline__ u32 nl16(uint4 q) {
    u32 r = nl4_top(q.w) >> 28;
    r = __funnelshift_l(nl4_top(q.z), r, 4);
    r = __funnelshift_l(nl4_top(q.y), r, 4);
    return __funnelshift_l(nl4_top(q.x), r, 4);
}
// bits [a, b) of a 32-bit word (a, b any integers)
__device__ __forceinline__ u32 bit_range(int a, int b) {
    const u32 below_b = b >= 32 ? 0xffffffffu : b <= 0 ? 0u : (1u << b) - 1u;
    const u32 below_a = a >= 32 ? 0xffffffffu : a <= 0 ? 0u : (1u << a) - 1u;
    return below_b & ~below_a;
}

// ---- newline count of [text, text + n): 4 x 128-bit loads in flight per lane over the aligned body, bytes at the ends
__global__ void __launch_bounds__(256) brc_count_kernel(const unsigned char* __restrict__ text, i64 n,
                                                        unsigned long long* __restrict__ out) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(text);
    const uintptr_t body0 = (a0 + 15) & ~(uintptr_t)15, body1 = (a0 + (uintptr_t)n) & ~(uintptr_t)15;
    u32 cnt = 0;
    const i64 gtid = (i64)blockIdx.x * 256 + threadIdx.x;
    if (body1 <= body0) {  // shorter than one aligned granule
        if (gtid == 0)
            for (i64 j = 0; j < n; j++) cnt += text[j] == '\n';
    } else {
        if (gtid == 0) {
            for (uintptr_t p = a0; p < body0; p++) cnt += *reinterpret_cast<const unsigned char*>(p) == '\n';
            for (uintptr_t p = body1; p < a0 + (uintptr_t)n; p++) cnt += *reinterpret_cast<const unsigned char*>(p) == '\n';
        }
        const uint4* g = reinterpret_cast<const uint4*>(body0);
        const i64 ng = (i64)((body1 - body0) >> 4), stride = (i64)gridDim.x * 256;
        i64 c = gtid;
        for (; c + 3 * stride < ng; c += 4 * stride) {
            const uint4 q0 = __ldcs(g + c), q1 = __ldcs(g + c + stride), q2 = __ldcs(g + c + 2 * stride), q3 = __ldcs(g + c + 3 * stride);
            cnt += __popc(nl16(q0)) + __popc(nl16(q1)) + __popc(nl16(q2)) + __popc(nl16(q3));
        }
        for (; c < ng; c += stride) cnt += __popc(nl16(__ldcs(g + c)));
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ u32 s_c[8];
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < 8; w++) t += s_c[w];
        if (t) atomicAdd(out, (unsigned long long)t);
    }
}

// ---- one-pass kernel
// Bytes [bo, bo + 16) of the staged text (any alignment) as two little-endian 64-bit words, zero from byte `len` on:
// five aligned 32-bit shared-memory loads, neighbours funnel-shifted.  (Only names longer than 24 bytes come here.)
__device__ __forceinline__ void load16(const u32* __restrict__ s_w, int bo, int len, u64& lo, u64& hi) {
    const int wi = bo >> 2;
    const u32 sh = (u32)(bo & 3) * 8;
    const u32 a0 = s_w[wi], a1 = s_w[wi + 1], a2 = s_w[wi + 2], a3 = s_w[wi + 3], a4 = s_w[wi + 4];
    lo = (u64)__funnelshift_r(a0, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
    hi = (u64)__funnelshift_r(a2, a3, sh) | ((u64)__funnelshift_r(a3, a4, sh) << 32);
    if (len < 8) lo &= (1ull << (8 * len)) - 1ull;
    if (len < 16) hi = len <= 8 ? 0ull : hi & ((1ull << (8 * (len - 8))) - 1ull);
}

template <int TILE, int T, int CTAS>
__global__ void __launch_bounds__(T, CTAS) brc_onepass_kernel(const unsigned char* __restrict__ text, i64 n, i64 ntiles,
                                                           u64* __restrict__ states, BrcSlot* __restrict__ tab,
                                                           BrcCtl* __restrict__ ctl, i64 rows_capacity,
                                                           i64* __restrict__ id_out, double* __restrict__ temp_out,
                                                           u32* __restrict__ dict_names, int* __restrict__ dict_len,
                                                           i64 dict_capacity) {
    extern __shared__ uint4 brc_smem[];
    unsigned char* s_text = reinterpret_cast<unsigned char*>(brc_smem);
    const u32* s_w = reinterpret_cast<const u32*>(brc_smem);
    typedef BrcGeom<TILE> G;
    u32* s_mlut = reinterpret_cast<u32*>(s_text + G::OFF_MLUT);
    u32* s_mask = reinterpret_cast<u32*>(s_text + G::OFF_MASK);
    unsigned short* s_pos = reinterpret_cast<unsigned short*>(s_text + G::OFF_POS + 16);
    u32* s_wcnt = reinterpret_cast<u32*>(s_text + G::OFF_WCNT);
    int* s_misc = reinterpret_cast<int*>(s_text + G::OFF_MISC);
    constexpr int WPT = (G::WORDS + T - 1) / T;
    constexpr int NW = T / 32;
    static_assert(WPT * (T - 32) >= G::WORDS, "the last warp owns no mask words: it runs the look-back");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // [0] tile  [1] start of the first line  [2..3] first row  [4..5] mbarrier of the bulk copy  [6] bulk copy issued
    const u32 bar = (u32)__cvta_generic_to_shared(s_misc + 4);
    if (tid == 0) {
        // interior tiles (every staged granule inside the text's granules) arrive by ONE bulk copy of the TMA engine
        // (cp.async.bulk, SASS UBLKCP), issued by this thread as soon as it holds the ticket — before the CTA's first
        // barrier; the tiles at the two ends of the text take the per-granule path below (bytes outside the text read
        // as zero there)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const i64 tk = (i64)atomicAdd(&ctl->ticket, 1u);
        const uintptr_t ta = reinterpret_cast<uintptr_t>(text) + (uintptr_t)(tk * TILE);
        const uintptr_t gg = (ta - BRC_HALO) & ~(uintptr_t)15;
        const uintptr_t lo_b = reinterpret_cast<uintptr_t>(text) & ~(uintptr_t)15;
        const uintptr_t hi_b = (reinterpret_cast<uintptr_t>(text) + (uintptr_t)n + 15) & ~(uintptr_t)15;
        const bool bulk = KGB_BRC_BULK && gg >= lo_b && gg + G::STAGE <= hi_b;
        if (bulk) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)G::STAGE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (u32)__cvta_generic_to_shared(brc_smem)),
                         "l"(gg), "r"((u32)G::STAGE), "r"(bar)
                         : "memory");
        }
        s_misc[0] = (int)tk;
        s_misc[6] = bulk ? 1 : 0;
    }
#if KGB_BRC_MASK_LUT
    if (tid < (BRC_INSLOT + 1) * 8) {
        const int rem = (tid >> 3) - 4 * (tid & 7);
        s_mlut[tid] = rem >= 4 ? 0xffffffffu : rem <= 0 ? 0u : (1u << (8 * rem)) - 1u;
    }
#else
    (void)s_mlut;
#endif
    __syncthreads();
    const i64 tile = s_misc[0];
    const i64 byte0 = tile * TILE;
    const int tlen = (int)((n - byte0) < (i64)TILE ? (n - byte0) : (i64)TILE);
    // ---- stage the text: shared byte j <- the byte at address g0 + j, g0 = (tile start - halo) rounded down to 16;
    //      newline flags of every staged granule -> mask word j / 32
    const uintptr_t tile_addr = reinterpret_cast<uintptr_t>(text) + (uintptr_t)byte0;
    const uintptr_t g0 = (tile_addr - BRC_HALO) & ~(uintptr_t)15;
    const int so = (int)(tile_addr - g0);  // shared offset of the tile's byte 0: halo + misalignment
    {
        const uintptr_t lo_b = reinterpret_cast<uintptr_t>(text) & ~(uintptr_t)15;
        const uintptr_t hi_b = (reinterpret_cast<uintptr_t>(text) + (uintptr_t)n + 15) & ~(uintptr_t)15;
        const uint4* gsrc = reinterpret_cast<const uint4*>(g0);
        if (s_misc[6]) {  // the bulk copy is on its way: wait for its bytes, then take the newline flags from shared memory
            u32 done = 0;
            while (!done)
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done)
                             : "r"(bar)
                             : "memory");
#pragma unroll 3
            for (int c0 = 0; c0 < G::GRAN; c0 += T) {
                const int c = c0 + tid;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (c < G::GRAN) q = brc_smem[c];
                const u32 h = nl16(q);
                const u32 other = __shfl_xor_sync(0xffffffffu, h, 1);
                if (!(tid & 1) && c < G::GRAN) s_mask[c >> 1] = h | (other << 16);
            }
        } else if (g0 >= lo_b && g0 + G::STAGE <= hi_b) {  // (KGB_BRC_BULK=0: the same tiles through registers)
#pragma unroll 3
            for (int c0 = 0; c0 < G::GRAN; c0 += T) {
                const int c = c0 + tid;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (c < G::GRAN) {
                    q = __ldcs(gsrc + c);
                    brc_smem[c] = q;
                }
                const u32 h = nl16(q);
                const u32 other = __shfl_xor_sync(0xffffffffu, h, 1);
                if (!(tid & 1) && c < G::GRAN) s_mask[c >> 1] = h | (other << 16);
            }
        } else {
#pragma unroll 1
            for (int c0 = 0; c0 < G::GRAN; c0 += T) {
                const int c = c0 + tid;
                const uintptr_t ga = g0 + (uintptr_t)c * 16;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (c < G::GRAN && ga >= lo_b && ga < hi_b) q = __ldcs(reinterpret_cast<const uint4*>(ga));
                if (c < G::GRAN) brc_smem[c] = q;
                const u32 h = nl16(q);
                const u32 other = __shfl_xor_sync(0xffffffffu, h, 1);
                if (!(tid & 1) && c < G::GRAN) s_mask[c >> 1] = h | (other << 16);
            }
        }
    }
    __syncthreads();
    // ---- this thread's mask words, restricted to the tile's own bytes; line numbers by a block scan of popcounts
    const int w0 = tid * WPT;
    u32 m[WPT];
    u32 cnt = 0;
    {
        const int jhi = so + tlen - 1;  // last byte of the tile
        const int wlo = so >> 5, whi = jhi >> 5;
        const u32 mlo = 0xffffffffu << (so & 31), mhi = 0xffffffffu >> (31 - (jhi & 31));
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            const int w = w0 + j;
            u32 v = (w >= wlo && w <= whi) ? s_mask[w] : 0u;
            if (w == wlo) v &= mlo;
            if (w == whi) v &= mhi;
            m[j] = v;
            cnt += __popc(v);
        }
    }
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wcnt[warp] = inc;
    if (tid == 32) {  // start of the line that ends first in this tile, relative to the tile: walk the halo's masks back
        const int jmin = so - (int)(byte0 < (i64)BRC_HALO ? byte0 : (i64)BRC_HALO);  // never before the text's first byte
        int rel = BRC_NO_START;
        for (int w = (so - 1) >> 5; w >= 0; --w) {
            const u32 v = s_mask[w] & bit_range(jmin - 32 * w, so - 32 * w);
            if (v) {
                rel = 32 * w + (31 - __clz(v)) + 1 - so;
                break;
            }
            if (32 * w <= jmin) break;
        }
        if (rel == BRC_NO_START && byte0 <= (i64)BRC_HALO) rel = (int)(0 - byte0);  // the text begins inside the halo
        s_misc[1] = rel;
    }
    if (tid == 64 && tile == ntiles - 1 && s_text[so + tlen - 1] != '\n') ctl->err = 4;
    __syncthreads();
    u32 k = inc - cnt, nlines;
    {  // warp totals -> this warp's offset and the tile's line count: one load + a shuffle scan instead of NW broadcast loads
        u32 wi = lane < NW ? s_wcnt[lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += v;
        }
        nlines = __shfl_sync(0xffffffffu, wi, NW - 1);
        const u32 before = __shfl_sync(0xffffffffu, wi, warp ? warp - 1 : 0);
        if (warp) k += before;
    }
    if (warp == NW - 1) {  // this warp owns no mask words: publish the tile's line count, fetch the tile's first row
        if (lane == 0) st_relaxed_u64(&states[tile], lb_pack(tile == 0 ? KGB_LB_PREFIX : KGB_LB_AGG, 0, nlines));
        const u64 excl = tile == 0 ? 0ull : lb_warp_lookback(states, tile, 0);
        if (lane == 0) {
            if (tile) st_relaxed_u64(&states[tile], lb_pack(KGB_LB_PREFIX, 0, excl + nlines));
            *reinterpret_cast<u64*>(s_misc + 2) = excl;
            if (tile == ntiles - 1) ctl->rows_total = excl + nlines;
        }
    }
    const bool too_many = nlines > (u32)G::MAX_LINES;  // lines shorter than 6 bytes: malformed
    if (!too_many) {
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            u32 mm = m[j];
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1;
                s_pos[k++] = (unsigned short)((w0 + j) * 32 + b - so);
            }
        }
    }
    __syncthreads();
    if (too_many) {
        if (tid == 0) ctl->err = 1;
        return;
    }
    const i64 base = (i64) * reinterpret_cast<const u64*>(s_misc + 2);
    const int start0 = s_misc[1];
    // rows of this tile that fit the caller's arrays; this tile's slices of the outputs
    const u32 nroom = base >= rows_capacity ? 0u : (rows_capacity - base < (i64)nlines ? (u32)(rows_capacity - base) : nlines);
    double* __restrict__ tout = temp_out + base;
    i64* __restrict__ iout = id_out + base;
    int bad = 0;  // first error seen by this thread (reported once, after the loop)
#pragma unroll 1
    for (u32 i = tid; i < nlines; i += T) {
        const int pe = s_pos[i];  // this line's '\n'
        const int prev = s_pos[(int)i - 1];
        const int st = i ? prev + 1 : start0;
        // ---- measurement, backwards from the newline: [-]d[d].d.  The eight bytes before the newline from three
        //      aligned words (the halo precedes the tile, so the address is never negative);
        //      tail_hi = bytes pe-4 .. pe-1 = {d2 | '-' | ';', d1, '.', f}
        const int qb = so + pe - 8;
        const uint2* q8 = reinterpret_cast<const uint2*>(s_text + (qb & ~7));
        const uint2 t0 = q8[0], t1 = q8[1];
        const bool tup = (qb & 4) != 0;
        const u32 q0 = tup ? t0.y : t0.x, q1 = tup ? t1.x : t0.y, q2 = tup ? t1.y : t1.x;
        const u32 tail_lo = __funnelshift_r(q0, q1, (u32)(qb & 3) * 8), tail_hi = __funnelshift_r(q1, q2, (u32)(qb & 3) * 8);
        const u32 x = tail_hi ^ 0x302e3030u;   // valid: byte3 = f, byte2 = 0, byte1 = d1, byte0 = d2 if the integer part has two digits
        const u32 y = x & 0xffffff00u;
        // bytes 1 and 3 must be <= 9, byte 2 must be 0 (7-bit adds: no carry between bytes)
        const u32 wrong = (((y & 0x7f007f00u) + 0x76007600u) | y) & 0x80ff8000u;
        const bool two = (x & 255u) < 10u;
        const u32 c = two ? (tail_lo >> 24) : (tail_hi & 255u);
        const bool neg = c == '-';
        const int back = 4 + (int)two + (int)neg;  // distance from the newline to the ';'
        const u32 semi = __byte_perm(tail_lo, tail_hi, (u32)(8 - back)) & 255u;
        const int len = pe - back - st;
        if (st == BRC_NO_START || len > BRC_NAME_BYTES - 1) {  // longer than the halo / than a dictionary row
            bad = bad ? bad : 2;
            continue;
        }
        if (wrong != 0u || semi != ';' || len <= 0) {
            bad = bad ? bad : 1;
            continue;
        }
        // tenths = 100 d2 + 10 d1 + f as one byte-wise dot product.  float("d.d") is the correctly rounded tenths / 10:
        // tenths * 0.1 corrected once by its exact residual (verified for all 1000 values, tests/test_ingest.py)
        const u32 tenths = __dp4a(two ? x : y, 0x01000a64u, 0u);
        const double tf = (double)(int)tenths;
        const double q = tf * 0.1;
        const double v = __fma_rn(__fma_rn(-q, 10.0, tf), 0.1, q);
        const u64 vbits = (u64)__double_as_longlong(v) | ((u64)neg << 63);
        if (i < nroom) tout[i] = __longlong_as_double((i64)vbits);
        // ---- station: the first 24 bytes travel in six registers; hash = multilinear over the words, one multiply to mix
        const int nb = so + st;  // shared byte offset of the name
        // 32 bytes from the 8-byte boundary below the name, as four 64-bit loads (the shared-memory pipe is this kernel's
        // bound: a 32-bit load of 32 scattered lines costs ~4.5 wavefronts, a 64-bit one ~5 for twice the bytes)
        const uint2* p8 = reinterpret_cast<const uint2*>(s_text + (nb & ~7));
        const uint2 p0 = p8[0], p1 = p8[1], p2 = p8[2], p3 = p8[3];
        const bool up = (nb & 4) != 0;  // the name starts in the upper word of the first pair
        const u32 a0 = up ? p0.y : p0.x, a1 = up ? p1.x : p0.y, a2 = up ? p1.y : p1.x, a3 = up ? p2.x : p1.y;
        const u32 a4 = up ? p2.y : p2.x, a5 = up ? p3.x : p2.y, a6 = up ? p3.y : p3.x;
        const u32 sh = (u32)(nb & 3) * 8;
#if KGB_BRC_MASK_LUT
        const int lc = len < BRC_INSLOT ? len : BRC_INSLOT;
        const uint4 ma = *reinterpret_cast<const uint4*>(s_mlut + lc * 8);
        const uint2 mb = *reinterpret_cast<const uint2*>(s_mlut + lc * 8 + 4);
        const u32 x0 = __funnelshift_r(a0, a1, sh) & ma.x, x1 = __funnelshift_r(a1, a2, sh) & ma.y;
        const u32 x2 = __funnelshift_r(a2, a3, sh) & ma.z, x3 = __funnelshift_r(a3, a4, sh) & ma.w;
        const u32 x4 = __funnelshift_r(a4, a5, sh) & mb.x, x5 = __funnelshift_r(a5, a6, sh) & mb.y;
#else
        // byte masks of a name of `len` bytes without a table: word k keeps its low min(max(len - 4k, 0), 4) bytes =
        // all-ones shifted right by max(32 (k + 1) - 8 len, 0) bits, the clamped funnel shift giving 0 from 32 bits on
        const int t8 = -8 * len;
#define KGB_BRC_M(k) __funnelshift_rc(0xffffffffu, 0u, (u32)max(t8 + 32 * ((k) + 1), 0))
        const u32 x0 = __funnelshift_r(a0, a1, sh) & KGB_BRC_M(0), x1 = __funnelshift_r(a1, a2, sh) & KGB_BRC_M(1);
        const u32 x2 = __funnelshift_r(a2, a3, sh) & KGB_BRC_M(2), x3 = __funnelshift_r(a3, a4, sh) & KGB_BRC_M(3);
        const u32 x4 = __funnelshift_r(a4, a5, sh) & KGB_BRC_M(4), x5 = __funnelshift_r(a5, a6, sh) & KGB_BRC_M(5);
#undef KGB_BRC_M
#endif
        u64 h = (u64)(u32)len * 0x9e3779b1ull + 0x7f4a7c15f39cc060ull;
        h += (u64)x0 * 0x85ebca6bull;
        h += (u64)x1 * 0xc2b2ae35ull;
        h += (u64)x2 * 0x27d4eb2full;
        h += (u64)x3 * 0x165667b1ull;
        h += (u64)x4 * 0xd3a2646dull;
        h += (u64)x5 * 0xfd7046c5ull;
        if (len > BRC_INSLOT) {  // rare for station names
            for (int cb = BRC_INSLOT; cb < len; cb += 16) {
                u64 lo, hi;
                load16(s_w, nb + cb, len - cb, lo, hi);
                h = (h ^ (h >> 32) ^ lo) * BRC_MIX;
                h = (h ^ (h >> 32) ^ hi) * BRC_MIX;
            }
        }
        h ^= h >> 32;
        h *= BRC_MIX;
        h ^= h >> 29;
        const u64 n0 = (u64)x0 | ((u64)x1 << 32), n1 = (u64)x2 | ((u64)x3 << 32), n2 = (u64)x4 | ((u64)x5 << 32);
        const u64 tag = ((h >> 8) & 0x7ffffffffull) | 1ull;  // 35 bits, never 0
        const u64 expect = (tag << 29) | BRC_PUB | (u64)(u32)len;  // a published slot of this name, but for its id
        u32 slot = (u32)(h >> (64 - BRC_LOG2_SLOTS));
        i64 id = -1;
        for (u32 probe = 0; probe < (u32)BRC_SLOTS; probe++) {
            Slot4 sv = ld_slot(&tab[slot]);
            // the common case in one test: tag, published, length and 24 name bytes all equal
            u64 d = ((sv.w0 ^ expect) & ~(0xfffffull << 8)) | (sv.n0 ^ n0) | (sv.n1 ^ n1) | (sv.n2 ^ n2);
            if (d == 0 && len <= BRC_INSLOT) {
                id = (i64)((sv.w0 >> 8) & 0xfffffull);
                break;
            }
            if (sv.w0 == 0) {
                const u64 old = atomicCAS((unsigned long long*)&tab[slot].w0, 0ull, (unsigned long long)(tag << 29));
                if (old == 0) {  // ours: draw the next dense id, store the name, then publish {id, len}
                    const u64 nid = atomicAdd(&ctl->n_names, 1ull);
                    st_relaxed_u64(&tab[slot].n0, n0);
                    st_relaxed_u64(&tab[slot].n1, n1);
                    st_relaxed_u64(&tab[slot].n2, n2);
                    if ((i64)nid < dict_capacity) {
                        uint4* rowp = reinterpret_cast<uint4*>(dict_names + nid * BRC_NAME_WORDS);
                        for (int cb = 0; cb < len; cb += 16) {  // zero-filled to the next multiple of 16 bytes
                            u64 lo, hi;
                            load16(s_w, nb + cb, len - cb, lo, hi);
                            __stcg(rowp + (cb >> 4), make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)));
                        }
                        dict_len[nid] = len;
                    } else {
                        bad = bad ? bad : 2;
                    }
                    __threadfence();
                    st_relaxed_u64(&tab[slot].w0, expect | ((nid & 0xfffffull) << 8));
                    id = (i64)nid;
                    break;
                }
                sv.w0 = old & ~BRC_PUB;  // somebody else's claim: if it is our tag, read the slot again below
            }
            if ((sv.w0 >> 29) == tag) {
                for (u32 spins = 0; !(sv.w0 & BRC_PUB); spins++) {  // the claimant publishes right after its CAS
                    if (spins) __nanosleep(32);
                    sv = ld_slot(&tab[slot]);
                    if (spins > (1u << 22)) break;
                }
                if (!(sv.w0 & BRC_PUB)) {
                    bad = bad ? bad : 3;
                    break;
                }
                d = ((sv.w0 ^ expect) & ~(0xfffffull << 8)) | (sv.n0 ^ n0) | (sv.n1 ^ n1) | (sv.n2 ^ n2);
                if (d != 0) {  // published slots are final: a second read rules out a torn first one
                    sv = ld_slot(&tab[slot]);
                    d = ((sv.w0 ^ expect) & ~(0xfffffull << 8)) | (sv.n0 ^ n0) | (sv.n1 ^ n1) | (sv.n2 ^ n2);
                }
                const i64 sid = (i64)((sv.w0 >> 8) & 0xfffffull);
                if (d == 0 && sid < dict_capacity) {
                    u64 diff = 0;
                    if (len > BRC_INSLOT) {  // the rest of the name against the dictionary row, 16 bytes per load
                        // no fence on this side: the row's address is computed from the id just read, the loads go to
                        // L2 (the point of coherence) and the writer fenced between the row and the id
                        const uint4* rowp = reinterpret_cast<const uint4*>(dict_names + sid * BRC_NAME_WORDS);
                        for (int cb = 16; cb < len; cb += 16) {
                            const uint4 r = __ldcg(rowp + (cb >> 4));
                            u64 lo, hi;
                            load16(s_w, nb + cb, len - cb, lo, hi);
                            diff |= (lo ^ ((u64)r.x | ((u64)r.y << 32))) | (hi ^ ((u64)r.z | ((u64)r.w << 32)));
                        }
                    }
                    if (diff == 0) {
                        id = sid;
                        break;
                    }
                }
            }
            slot = (slot + 1) & (u32)(BRC_SLOTS - 1);
        }
        if (id < 0) bad = bad ? bad : 2;
        if (i < nroom) iout[i] = id;
    }
    if (bad) ctl->err = bad;
}

template <int TILE, int T, int CTAS>
static cudaError_t launch_onepass(cudaStream_t st, const unsigned char* text, i64 n, u64* states, BrcSlot* tab, BrcCtl* ctl,
                                  i64 rows_capacity, i64* id_out, double* temp_out, u32* dict_names, int* dict_len,
                                  i64 dict_capacity) {
    static bool opted[64] = {};
    const int dev = current_device();
    if (dev >= 0 && dev < 64 && !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(brc_onepass_kernel<TILE, T, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             BrcGeom<TILE>::SMEM);
        if (e != cudaSuccess) return e;
        opted[dev] = true;
    }
    const i64 ntiles = brc_tiles(n, TILE);
    brc_onepass_kernel<TILE, T, CTAS><<<(unsigned)ntiles, T, BrcGeom<TILE>::SMEM, st>>>(
        text, n, ntiles, states, tab, ctl, rows_capacity, id_out, temp_out, dict_names, dict_len, dict_capacity);
    return cudaGetLastError();
}

static cudaError_t launch_count(const unsigned char* text, i64 n, unsigned long long* out, cudaStream_t st) {
    long long blocks = (n / 16 + 256 * 4 - 1) / (256 * 4);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    brc_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(text, n, out);
    return cudaGetLastError();
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_brc_workspace_bytes(int64_t n_bytes, size_t* out_bytes) {
    KGB_REQUIRE(out_bytes && n_bytes >= 0, "kgb_brc_workspace_bytes: bad arguments");
    *out_bytes = carve_brc(nullptr, n_bytes).total;
    return KGB_OK;
}

// Shared front of kgb_brc_scan / kgb_brc_estimate: newline count of the whole text (exact) or of three windows.
static int brc_count(const uint8_t* text, int64_t n_bytes, bool exact, int64_t* out_host, void* workspace, size_t workspace_bytes,
                     void* stream, const char* who) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(out_host && n_bytes >= 0, "%s: bad arguments", who);
    *out_host = 0;
    if (n_bytes == 0) return KGB_OK;
    KGB_REQUIRE(text && workspace && workspace_bytes >= 256, "%s: null pointer or workspace too small", who);
    cudaStream_t st = (cudaStream_t)stream;
    BrcWs w = carve_brc(workspace, n_bytes);
    KGB_CUDA_CHECK(cudaMemsetAsync(w.ctl, 0, 256, st));
    const int64_t WIN = 4ll << 20;
    int64_t sampled = n_bytes;
    if (exact || n_bytes <= 3 * WIN) {
        KGB_CUDA_CHECK(launch_count(text, n_bytes, &w.ctl->sample_count, st));
    } else {
        sampled = 3 * WIN;
        const int64_t offs[3] = {0, (n_bytes - WIN) / 2, n_bytes - WIN};
        for (int k = 0; k < 3; k++) KGB_CUDA_CHECK(launch_count(text + offs[k], WIN, &w.ctl->sample_count, st));
    }
    unsigned long long cnt = 0;
    unsigned char last = 0;
    KGB_CUDA_CHECK(cudaMemcpyAsync(&cnt, &w.ctl->sample_count, 8, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaMemcpyAsync(&last, text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    KGB_REQUIRE(last == '\n', "%s: the text must end with a newline", who);
    if (sampled == n_bytes) {
        *out_host = (int64_t)cnt;
    } else {  // rows per byte of the sample, 2 % and 4096 rows of slack; never more than one row per byte
        const double est = (double)n_bytes * ((double)cnt / (double)sampled) * 1.02 + 4096.0;
        *out_host = est < (double)n_bytes ? (int64_t)est : n_bytes;
    }
    return KGB_OK;
}

extern "C" int kgb_brc_scan(const uint8_t* text, int64_t n_bytes, int64_t* rows_out_host, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return brc_count(text, n_bytes, true, rows_out_host, workspace, workspace_bytes, stream, "kgb_brc_scan");
}

extern "C" int kgb_brc_estimate(const uint8_t* text, int64_t n_bytes, int64_t* rows_estimate_out_host, void* workspace,
                                size_t workspace_bytes, void* stream) {
    return brc_count(text, n_bytes, false, rows_estimate_out_host, workspace, workspace_bytes, stream, "kgb_brc_estimate");
}

extern "C" int kgb_brc_parse(const uint8_t* text, int64_t n_bytes, int64_t rows_capacity, int64_t* id_out, double* temp_out,
                             uint8_t* dict_names_out, int32_t* dict_len_out, int64_t dict_capacity, int64_t* rows_out_host,
                             int64_t* n_names_out_host, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(rows_out_host && n_names_out_host && n_bytes >= 0 && rows_capacity >= 0 && dict_capacity >= 0,
                "kgb_brc_parse: bad arguments");
    *n_names_out_host = 0;
    *rows_out_host = 0;
    if (n_bytes == 0) return KGB_OK;
    KGB_REQUIRE(text && (rows_capacity == 0 || (id_out && temp_out)) && dict_names_out && dict_len_out && workspace &&
                    workspace_bytes >= carve_brc(nullptr, n_bytes).total,
                "kgb_brc_parse: null pointer or workspace too small");
    KGB_REQUIRE(dict_capacity <= (int64_t)(BRC_SLOTS / 2), "kgb_brc_parse: dictionary capacity above %lld",
                (long long)(BRC_SLOTS / 2));
    KGB_REQUIRE(n_bytes < (1ll << 46), "kgb_brc_parse: text too long");
    cudaStream_t st = (cudaStream_t)stream;
    BrcWs w = carve_brc(workspace, n_bytes);
    KGB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
    // tile bytes x CTA threads x CTAs per SM (KGB200_BRC_VARIANT: tuning)
    static const int variant = getenv("KGB200_BRC_VARIANT") ? atoi(getenv("KGB200_BRC_VARIANT")) : 0;
#define KGB_BRC_GO(TILE, T, CTAS)                                                                                              \
    KGB_CUDA_CHECK((launch_onepass<TILE, T, CTAS>(st, text, n_bytes, w.states, w.tab, w.ctl, rows_capacity, (i64*)id_out, temp_out, \
                                                  (u32*)dict_names_out, dict_len_out, dict_capacity)))
    switch (variant) {
        case 1: KGB_BRC_GO(32768, 384, 4); break;
        case 2: KGB_BRC_GO(49152, 512, 3); break;
        case 3: KGB_BRC_GO(32768, 512, 3); break;
        case 4: KGB_BRC_GO(16384, 256, 6); break;
        case 5: KGB_BRC_GO(24576, 384, 4); break;
        case 6: KGB_BRC_GO(65536, 512, 2); break;
        case 7: KGB_BRC_GO(65536, 1024, 1); break;
        default: KGB_BRC_GO(65536, 768, 2); break;
    }
#undef KGB_BRC_GO
    BrcCtl h;
    KGB_CUDA_CHECK(cudaMemcpyAsync(&h, w.ctl, sizeof(BrcCtl), cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    KGB_REQUIRE(h.err != 4, "kgb_brc_parse: the text must end with a newline");
    KGB_REQUIRE(h.err != 1, "kgb_brc_parse: malformed row (expected name;[-]d[d].d)");
    KGB_REQUIRE(h.err != 2, "kgb_brc_parse: more than %lld distinct names (or a name longer than 255 bytes)",
                (long long)dict_capacity);
    KGB_REQUIRE(h.err == 0, "kgb_brc_parse: dictionary slot was never published");
    *rows_out_host = (int64_t)h.rows_total;
    *n_names_out_host = (int64_t)h.n_names;
    return KGB_OK;
}
