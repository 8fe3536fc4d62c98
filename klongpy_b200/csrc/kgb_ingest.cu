// 1brc text ingest: `name;[-]d[d].d\n` lines -> (dense station id, float64 temperature) columns + station dictionary.
// Replaces the per-line Python loop of the reference's examples/1brc/par.py:57-88 and the string grouping of
// worker.kg:5.  Byte work, HBM-bound by design:
//
//   brc_mask_kernel   streams the text once with 2 x 128-bit loads per lane (32 bytes -> one mask word, bit j = byte j
//                     is '\n'), one coalesced 128-byte store of mask words per warp and per 1024 bytes, plus (when a
//                     warp saw a newline) one RED of its count into the counter of its 65536-byte tile — the same
//                     mask / tile-counter layout as the stream compaction in kgb_select.cuh, whose offsets kernel
//                     is reused.  Traffic: the text once + n/8 bytes.
//   brc_parse_kernel  one CTA per 65536-byte tile with newlines: the tile's mask words are expanded into a list of
//                     newline positions in shared memory (block scan of popcounts = row numbers), then line i goes
//                     to thread i mod 512, so a warp reads ~5 consecutive 128-byte lines of text and writes its 16
//                     bytes per row coalesced.  Per line: the measurement is parsed backwards from the newline in
//                     fixed point; the name is hashed four bytes at a time (aligned loads, funnel shift) and
//                     resolved in an L2-resident open-addressing table of 16-byte slots {hash, id<<8|len}.  A known
//                     name costs one slot read plus a word-wise compare against the dictionary row of its id; a
//                     new name takes the slot with one CAS, draws the next dense id, stores its bytes into the
//                     dictionary row and publishes {id, len}.  Traffic: the text a second time + 16 B per row.
//   v1 (one thread per 128-byte segment, byte-wise FNV-1a and compare): 1e8 rows in 18.8 ms end to end.
#include <cstdlib>

#include "kgb_select.cuh"

namespace kgb {

constexpr int BRC_LOG2_SLOTS = 20;  // 32 MB; only the sectors of live names are ever hot (10k names: 320 KB)
constexpr u64 BRC_SLOTS = 1ull << BRC_LOG2_SLOTS;
constexpr u64 BRC_UNSET = ~0ull;
constexpr int BRC_NAME_BYTES = 256;                 // dictionary row: a name of <= 255 bytes as zero-extended 32-bit words
constexpr int BRC_NAME_WORDS = BRC_NAME_BYTES / 4;
constexpr int BRC_MAX_LINES = (int)(SEL_TILE / 6) + 1;  // the shortest valid line, "a;0.0\n", is 6 bytes
struct __align__(32) BrcSlot {  // one 32-byte sector: a lookup is ONE 256-bit load
    u64 hash;   // 0 = free; set once by the claimant's CAS
    u64 idlen;  // (dense id in first-arrival order << 8) | name length; BRC_UNSET until the name is published
    u64 n0, n1; // the name's first 16 bytes (zero-extended): names of <= 16 bytes never touch the dictionary rows
};
struct Slot4 {
    u64 hash, idlen, n0, n1;
};
__device__ __forceinline__ Slot4 ld_slot(const BrcSlot* p) {
    Slot4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v.hash), "=l"(v.idlen), "=l"(v.n0), "=l"(v.n1)
                 : "l"(p)
                 : "memory");
    return v;
}
struct BrcCtl {
    unsigned long long n_names;
    int err;  // 1 malformed row, 2 dictionary full / name too long, 3 a slot was never published
};

// workspace: [select state (tile counters)] [mask words] [BrcCtl 256 B] [table]
struct BrcWs {
    unsigned long long* tilecnt;
    u32* masks;
    BrcCtl* ctl;
    BrcSlot* tab;
    size_t total;
};
static BrcWs carve_brc(void* ws, i64 n) {
    BrcWs w;
    char* p = (char*)ws;
    size_t o = 0;
    w.tilecnt = (unsigned long long*)(p + 64);
    o += select_state_bytes(n);
    w.masks = (u32*)(p + o);
    o += align_up(select_mask_words(n) * 4, 256);
    w.ctl = (BrcCtl*)(p + o);
    o += 256;
    w.tab = (BrcSlot*)(p + o);
    o += (size_t)BRC_SLOTS * sizeof(BrcSlot);
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

constexpr int BRC_MASK_ROWS = 8;                         // 512-byte rows per warp step: 8 x 128-bit loads in flight per lane
constexpr int BRC_MASK_STEP = BRC_MASK_ROWS * 512;       // bytes per warp step (divides the 65536-byte tile)

__global__ void __launch_bounds__(256) brc_mask_kernel(const unsigned char* __restrict__ text, i64 n, u32* __restrict__ masks,
                                                        unsigned long long* __restrict__ tilecnt) {
    const int lane = threadIdx.x & 31;
    const i64 nsteps = (n + BRC_MASK_STEP - 1) / BRC_MASK_STEP;
    const i64 wstride = (i64)gridDim.x * 8;
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    for (i64 step = (i64)blockIdx.x * 8 + (threadIdx.x >> 5); step < nsteps; step += wstride) {
        const i64 s0 = step * BRC_MASK_STEP;
        u32 h[BRC_MASK_ROWS];  // row r: 16 mask bits for bytes s0 + 512 r + 16 lane ..
        if (aligned && s0 + BRC_MASK_STEP <= n) {
            uint4 q[BRC_MASK_ROWS];
#pragma unroll
            for (int r = 0; r < BRC_MASK_ROWS; r++)  // volatile: all eight loads are issued before the first compare
                asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(q[r].x), "=r"(q[r].y), "=r"(q[r].z), "=r"(q[r].w)
                             : "l"(text + s0 + r * 512 + lane * 16));
#pragma unroll
            for (int r = 0; r < BRC_MASK_ROWS; r++) h[r] = nl16(q[r]);
        } else {
#pragma unroll 1
            for (int r = 0; r < BRC_MASK_ROWS; r++) {
                const i64 b0 = s0 + r * 512 + lane * 16;
                u32 w = 0;
                for (int j = 0; j < 16; j++)
                    if (b0 + j < n && text[b0 + j] == '\n') w |= 1u << j;
                h[r] = w;
            }
        }
        // a mask word covers 32 bytes = the halves of lanes 2k and 2k+1 of one row.  Rows are paired so that one
        // exchange serves both: the even lane assembles word k of row r, the odd lane word k of row r + 1.
        u32 hits = 0;
        const i64 wbase = s0 / 32;
        const i64 nwords = (n + SEL_WARP_ELEMS - 1) / SEL_WARP_ELEMS * 32;  // select_mask_words(n)
#pragma unroll
        for (int r = 0; r < BRC_MASK_ROWS; r += 2) {
            const u32 send = (lane & 1) ? h[r] : h[r + 1];
            const u32 recv = __shfl_xor_sync(0xffffffffu, send, 1);
            const u32 word = (lane & 1) ? (recv | (h[r + 1] << 16)) : (h[r] | (recv << 16));
            const i64 wi = wbase + (r + (lane & 1)) * 16 + (lane >> 1);
            if (wi < nwords) masks[wi] = word;
            hits += __popc(word);
        }
        hits = __reduce_add_sync(0xffffffffu, hits);
        if (lane == 0 && hits) atomicAdd(&tilecnt[s0 / SEL_TILE], (unsigned long long)hits);
    }
}

// ---- parse kernel: shared-memory layout (dynamic, ~94 KB: two CTAs per SM)
constexpr int BRC_HALO = 512;                                   // bytes before the tile kept with it: a line may start there
constexpr int BRC_TEXT_SMEM = BRC_HALO + (int)SEL_TILE + 64;    // + misalignment of the tile (< 16) + 20 bytes of over-read
constexpr int BRC_POS_SLOTS = (BRC_MAX_LINES + 7) & ~7;
constexpr size_t BRC_SMEM_BYTES = (size_t)BRC_TEXT_SMEM + 1000 * 8 + (size_t)BRC_POS_SLOTS * 2 + 32 * 4 + 16;
constexpr int BRC_NO_START = -(1 << 30);
constexpr u64 BRC_MIX = 0xd6e8feb86659fd93ull;

// Bytes [bo, bo + 16) of the staged text (any alignment) as two little-endian 64-bit words, zero from byte `len` on:
// five aligned 32-bit shared-memory loads, neighbours funnel-shifted.
__device__ __forceinline__ void load16(const u32* __restrict__ s_w, int bo, int len, u64& lo, u64& hi) {
    const int wi = bo >> 2;
    const u32 sh = (u32)(bo & 3) * 8;
    const u32 a0 = s_w[wi], a1 = s_w[wi + 1], a2 = s_w[wi + 2], a3 = s_w[wi + 3], a4 = s_w[wi + 4];
    lo = (u64)__funnelshift_r(a0, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
    hi = (u64)__funnelshift_r(a2, a3, sh) | ((u64)__funnelshift_r(a3, a4, sh) << 32);
    if (len < 8) lo &= (1ull << (8 * len)) - 1ull;
    if (len < 16) hi = len <= 8 ? 0ull : hi & ((1ull << (8 * (len - 8))) - 1ull);
}

__global__ void __launch_bounds__(256) brc_init_kernel(BrcSlot* __restrict__ tab, BrcCtl* __restrict__ ctl) {
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i < BRC_SLOTS) *reinterpret_cast<uint4*>(&tab[i]) = make_uint4(0u, 0u, 0xffffffffu, 0xffffffffu);  // {hash 0, idlen UNSET}
    if (i == 0) {
        ctl->n_names = 0;
        ctl->err = 0;
    }
}

// One CTA per 65536-byte tile.  The tile (plus a 512-byte halo in front, for the line that straddles the tile edge) is
// staged in shared memory with coalesced 128-bit loads, the tile's newline positions are expanded from the mask words
// into a second shared array, then LINE i of the tile goes to thread i mod T: all byte work is 32-bit-addressed shared
// memory traffic (neighbouring lanes parse neighbouring lines: few bank conflicts) and the 16 bytes of output per row
// are coalesced stores.
template <int T, int MIN_CTAS>
__global__ void __launch_bounds__(T, MIN_CTAS) brc_parse_kernel(const unsigned char* __restrict__ text, i64 n,
                                                                const u32* __restrict__ masks, i64 nwords,
                                                                const unsigned long long* __restrict__ offs,
                                                                BrcSlot* __restrict__ tab, BrcCtl* __restrict__ ctl,
                                                                i64* __restrict__ id_out, double* __restrict__ temp_out,
                                                                u32* __restrict__ dict_names, int* __restrict__ dict_len,
                                                                i64 dict_capacity) {
    extern __shared__ uint4 brc_smem[];
    unsigned char* s_text = reinterpret_cast<unsigned char*>(brc_smem);
    const u32* s_w = reinterpret_cast<const u32*>(brc_smem);
    double* s_tenth = reinterpret_cast<double*>(s_text + BRC_TEXT_SMEM);  // tenths -> tenths / 10.0, correctly rounded
    unsigned short* s_pos = reinterpret_cast<unsigned short*>(s_tenth + 1000);
    u32* s_wcnt = reinterpret_cast<u32*>(s_pos + BRC_POS_SLOTS);
    int* s_start0 = reinterpret_cast<int*>(s_wcnt + 32);
    constexpr int WPT = (SEL_TILE_WORDS + T - 1) / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 tile = blockIdx.x;
    const unsigned long long base = offs[tile];
    const u32 nlines = (u32)(offs[tile + 1] - base);
    if (nlines == 0) return;  // no line ends in this tile
    if (nlines > (u32)BRC_MAX_LINES) {  // lines shorter than 6 bytes: malformed
        if (tid == 0) ctl->err = 1;
        return;
    }
    const i64 byte0 = tile * SEL_TILE;
    // ---- stage the text: shared byte j <- the byte at address g0 + j, g0 = (tile start - halo) rounded down to 16
    const uintptr_t tile_addr = reinterpret_cast<uintptr_t>(text) + (uintptr_t)byte0;
    const uintptr_t g0 = (tile_addr - BRC_HALO) & ~(uintptr_t)15;
    const int so = (int)(tile_addr - g0);  // shared offset of the tile's byte 0: halo + misalignment
    {
        const uintptr_t lo_b = reinterpret_cast<uintptr_t>(text) & ~(uintptr_t)15;
        const uintptr_t hi_b = (reinterpret_cast<uintptr_t>(text) + (uintptr_t)n + 15) & ~(uintptr_t)15;
#pragma unroll 4
        for (int c = tid; c < BRC_TEXT_SMEM / 16; c += T) {
            const uintptr_t ga = g0 + (uintptr_t)c * 16;
            if (ga >= lo_b && ga < hi_b) brc_smem[c] = __ldcs(reinterpret_cast<const uint4*>(ga));
        }
    }
    if (tid == 0) {  // start of the line that ends first in this tile, relative to the tile: walk the mask words back
        int rel = BRC_NO_START;
        i64 w = byte0 >> 5;
        for (int back = 0; back <= BRC_HALO / 32; back++) {
            if (w == 0) {
                rel = (int)(0 - byte0);
                break;
            }
            const u32 m = masks[--w];
            if (m) {
                rel = (int)(w * 32 + (31 - __clz(m)) + 1 - byte0);
                break;
            }
        }
        if (rel < -BRC_HALO) rel = BRC_NO_START;
        *s_start0 = rel;
    }
    for (int t = tid; t < 1000; t += T) s_tenth[t] = (double)t / 10.0;
    // ---- newline positions of the tile, ascending
    const int lw0 = tid * WPT;
    u32 m[WPT];
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        const i64 gw = tile * SEL_TILE_WORDS + lw0 + j;
        m[j] = (lw0 + j < SEL_TILE_WORDS && gw < nwords) ? masks[gw] : 0u;
    }
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < WPT; j++) cnt += __popc(m[j]);
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wcnt[warp] = inc;
    __syncthreads();
    u32 k = inc - cnt;
#pragma unroll
    for (int w = 0; w < T / 32; w++)
        if (w < warp) k += s_wcnt[w];
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        u32 mm = m[j];
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1;
            s_pos[k++] = (unsigned short)((lw0 + j) * 32 + b);
        }
    }
    __syncthreads();
#pragma unroll 1
    for (u32 i = tid; i < nlines; i += T) {
        const int pe = s_pos[i];  // this line's '\n'
        const int st = i ? (int)s_pos[i - 1] + 1 : *s_start0;
        if (st == BRC_NO_START) {  // longer than the halo: no name of <= 255 bytes fits
            ctl->err = 2;
            continue;
        }
        // ---- measurement, backwards from the newline: [-]d[d].d
        if (pe - st < 5) {  // shortest line: "a;0.0"
            ctl->err = 1;
            continue;
        }
        // the eight bytes before the newline from three aligned words (every shared load of a warp costs ~4 wavefronts:
        // 32 lines span ~550 bytes, four words per bank)
        const int qb = so + pe - 8;  // >= 0: the halo precedes the tile
        const u32 q0 = s_w[qb >> 2], q1 = s_w[(qb >> 2) + 1], q2 = s_w[(qb >> 2) + 2];
        const u32 tail_lo = __funnelshift_r(q0, q1, (u32)(qb & 3) * 8), tail_hi = __funnelshift_r(q1, q2, (u32)(qb & 3) * 8);
        const int f = (int)(tail_hi >> 24) - '0', dot = (int)((tail_hi >> 16) & 255u), d1 = (int)((tail_hi >> 8) & 255u) - '0';
        const int c4 = (int)(tail_hi & 255u), c5 = (int)(tail_lo >> 24), c6 = pe - 6 >= st ? (int)((tail_lo >> 16) & 255u) : 0;
        int d2 = 0, back = 4;  // `back`: distance from the newline to the ';'
        int c = c4;
        if (c >= '0' && c <= '9') {
            d2 = c - '0';
            back = 5;
            c = c5;
        }
        bool neg = false;
        if (c == '-') {
            neg = true;
            back++;
            c = back == 5 ? c5 : c6;
        }
        const int len = pe - back - st;
        if ((unsigned)f > 9u || (unsigned)d1 > 9u || dot != '.' || c != ';' || len <= 0) {
            ctl->err = 1;
            continue;
        }
        const double v = s_tenth[(d2 * 10 + d1) * 10 + f];  // correctly rounded quotient of two exact integers == float("d.d")
        temp_out[base + i] = neg ? -v : v;
        if (len > BRC_NAME_BYTES - 1) {
            ctl->err = 2;
            continue;
        }
        // ---- station: the first 16 bytes travel in two registers, the hash takes 16 bytes at a time beyond them
        const int nb = so + st;  // shared byte offset of the name
        u64 n0, n1;
        load16(s_w, nb, len, n0, n1);
        u64 h = 0x9e3779b97f4a7c15ull ^ (u64)len;
        h = (h ^ n0) * BRC_MIX;
        h = (h ^ (h >> 32) ^ n1) * BRC_MIX;
        for (int cb = 16; cb < len; cb += 16) {  // rare for station names
            u64 lo, hi;
            load16(s_w, nb + cb, len - cb, lo, hi);
            h = (h ^ (h >> 32) ^ lo) * BRC_MIX;
            h = (h ^ (h >> 32) ^ hi) * BRC_MIX;
        }
        h ^= h >> 32;
        h *= BRC_MIX;
        h ^= h >> 29;
        if (h == 0) h = 1;  // 0 marks a free slot
        u32 slot = (u32)(h >> (64 - BRC_LOG2_SLOTS));
        u64 id = BRC_UNSET;
        for (u32 probe = 0; probe < (u32)BRC_SLOTS && id == BRC_UNSET; probe++) {
            Slot4 sv = ld_slot(&tab[slot]);
            if (sv.hash == 0) {
                sv.hash = atomicCAS((unsigned long long*)&tab[slot].hash, 0ull, (unsigned long long)h);
                if (sv.hash == 0) {  // ours: draw the next dense id, store the name, then publish {id, len}
                    const u64 nid = atomicAdd(&ctl->n_names, 1ull);
                    st_relaxed_u64(&tab[slot].n0, n0);
                    st_relaxed_u64(&tab[slot].n1, n1);
                    if ((i64)nid < dict_capacity) {
                        uint4* row = reinterpret_cast<uint4*>(dict_names + nid * BRC_NAME_WORDS);
                        for (int cb = 0; cb < len; cb += 16) {  // zero-filled to the next multiple of 16 bytes
                            u64 lo, hi;
                            load16(s_w, nb + cb, len - cb, lo, hi);
                            __stcg(row + (cb >> 4), make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32)));
                        }
                        dict_len[nid] = len;
                    } else {
                        ctl->err = 2;
                    }
                    __threadfence();
                    st_relaxed_u64(&tab[slot].idlen, (nid << 8) | (u64)len);
                    id = nid;
                    break;
                }
                sv.idlen = BRC_UNSET;  // somebody else's claim: if it is our hash, read the slot again below
            }
            if (sv.hash == h) {
                for (u32 spins = 0; sv.idlen == BRC_UNSET; spins++) {  // the claimant publishes right after its CAS
                    if (spins) __nanosleep(32);
                    sv = ld_slot(&tab[slot]);
                    if (spins > (1u << 22)) {
                        ctl->err = 3;
                        break;
                    }
                }
                if (sv.idlen == BRC_UNSET) break;
                if (sv.n0 != n0 || sv.n1 != n1) sv = ld_slot(&tab[slot]);  // published slots are final: rules out a torn first read
                if ((int)(sv.idlen & 255) == len && sv.n0 == n0 && sv.n1 == n1 && (i64)(sv.idlen >> 8) < dict_capacity) {
                    u64 diff = 0;
                    if (len > 16) {  // the rest of the name against the dictionary row, 16 bytes per load
                        // no fence on this side: the row's address is computed from the id just read, the loads go to
                        // L2 (the point of coherence) and the writer fenced between the row and the id
                        const uint4* row = reinterpret_cast<const uint4*>(dict_names + (sv.idlen >> 8) * BRC_NAME_WORDS);
                        for (int cb = 16; cb < len; cb += 16) {
                            const uint4 r = __ldcg(row + (cb >> 4));
                            u64 lo, hi;
                            load16(s_w, nb + cb, len - cb, lo, hi);
                            diff |= (lo ^ ((u64)r.x | ((u64)r.y << 32))) | (hi ^ ((u64)r.z | ((u64)r.w << 32)));
                        }
                    }
                    if (diff == 0) id = sv.idlen >> 8;
                }
            }
            slot = (slot + 1) & (u32)(BRC_SLOTS - 1);
        }
        if (id == BRC_UNSET && ctl->err == 0) ctl->err = 2;
        id_out[base + i] = (i64)id;
    }
}

template <int T, int MIN_CTAS>
static cudaError_t launch_parse(i64 ntiles, cudaStream_t st, const unsigned char* text, i64 n, const u32* masks, i64 nwords,
                                const unsigned long long* offs, BrcSlot* tab, BrcCtl* ctl, i64* id_out, double* temp_out,
                                u32* dict_names, int* dict_len, i64 dict_capacity) {
    static bool opted[64] = {};
    const int dev = current_device();
    if (dev >= 0 && dev < 64 && !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(brc_parse_kernel<T, MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)BRC_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        opted[dev] = true;
    }
    brc_parse_kernel<T, MIN_CTAS><<<(unsigned)ntiles, T, BRC_SMEM_BYTES, st>>>(text, n, masks, nwords, offs, tab, ctl, id_out,
                                                                                temp_out, dict_names, dict_len, dict_capacity);
    return cudaGetLastError();
}

}  // namespace kgb

using namespace kgb;

extern "C" int kgb_brc_workspace_bytes(int64_t n_bytes, size_t* out_bytes) {
    KGB_REQUIRE(out_bytes && n_bytes >= 0, "kgb_brc_workspace_bytes: bad arguments");
    *out_bytes = carve_brc(nullptr, n_bytes).total;
    return KGB_OK;
}

extern "C" int kgb_brc_scan(const uint8_t* text, int64_t n_bytes, int64_t* rows_out_host, void* workspace,
                            size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(rows_out_host && n_bytes >= 0, "kgb_brc_scan: bad arguments");
    *rows_out_host = 0;
    if (n_bytes == 0) return KGB_OK;
    KGB_REQUIRE(text && workspace && workspace_bytes >= carve_brc(nullptr, n_bytes).total,
                "kgb_brc_scan: null pointer or workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    BrcWs w = carve_brc(workspace, n_bytes);
    KGB_CUDA_CHECK(cudaMemsetAsync(workspace, 0, select_state_bytes(n_bytes), st));
    const long long steps = (long long)((n_bytes + BRC_MASK_STEP - 1) / BRC_MASK_STEP);
    long long mb = (steps + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (mb > cap) mb = cap;
    brc_mask_kernel<<<(unsigned)mb, 256, 0, st>>>(text, n_bytes, w.masks, w.tilecnt);
    // the count lands in the tile table's last slot; a scratch word inside the control block receives the copy
    i64* dcount = (i64*)((char*)w.ctl + 128);
    const i64 ntiles = (i64)select_tiles(n_bytes);
    select_offsets_kernel<<<1, 1024, 0, st>>>(w.tilecnt, ntiles, dcount);
    KGB_CUDA_CHECK(cudaGetLastError());
    unsigned char last = 0;
    KGB_CUDA_CHECK(cudaMemcpyAsync(rows_out_host, dcount, 8, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaMemcpyAsync(&last, text + n_bytes - 1, 1, cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    KGB_REQUIRE(last == '\n', "kgb_brc_scan: the text must end with a newline");
    return KGB_OK;
}

extern "C" int kgb_brc_parse(const uint8_t* text, int64_t n_bytes, int64_t rows, int64_t* id_out, double* temp_out,
                             uint8_t* dict_names_out, int32_t* dict_len_out, int64_t dict_capacity,
                             int64_t* n_names_out_host, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_names_out_host && n_bytes >= 0 && rows >= 0 && dict_capacity >= 0, "kgb_brc_parse: bad arguments");
    *n_names_out_host = 0;
    if (n_bytes == 0 || rows == 0) return KGB_OK;
    KGB_REQUIRE(text && id_out && temp_out && dict_names_out && dict_len_out && workspace &&
                    workspace_bytes >= carve_brc(nullptr, n_bytes).total,
                "kgb_brc_parse: null pointer or workspace too small");
    KGB_REQUIRE(dict_capacity <= (int64_t)(BRC_SLOTS / 2), "kgb_brc_parse: dictionary capacity above %lld",
                (long long)(BRC_SLOTS / 2));
    cudaStream_t st = (cudaStream_t)stream;
    BrcWs w = carve_brc(workspace, n_bytes);
    brc_init_kernel<<<(unsigned)(BRC_SLOTS / 256), 256, 0, st>>>(w.tab, w.ctl);
    const i64 ntiles = (i64)select_tiles(n_bytes);
    // CTA size, two CTAs per SM: 512 -> 1.70 ms, 768 -> 1.63 ms, 1024 (32 registers, spills) -> 1.80 ms at 1e8 rows
    static const int threads = getenv("KGB200_BRC_THREADS") ? atoi(getenv("KGB200_BRC_THREADS")) : 768;
    const i64 nwords = (i64)select_mask_words(n_bytes);
    if (threads == 768)
        KGB_CUDA_CHECK((launch_parse<768, 2>(ntiles, st, text, n_bytes, w.masks, nwords, w.tilecnt, w.tab, w.ctl, (i64*)id_out,
                                             temp_out, (u32*)dict_names_out, dict_len_out, dict_capacity)));
    else if (threads == 1024)
        KGB_CUDA_CHECK((launch_parse<1024, 2>(ntiles, st, text, n_bytes, w.masks, nwords, w.tilecnt, w.tab, w.ctl, (i64*)id_out,
                                              temp_out, (u32*)dict_names_out, dict_len_out, dict_capacity)));
    else
        KGB_CUDA_CHECK((launch_parse<512, 2>(ntiles, st, text, n_bytes, w.masks, nwords, w.tilecnt, w.tab, w.ctl, (i64*)id_out,
                                             temp_out, (u32*)dict_names_out, dict_len_out, dict_capacity)));
    BrcCtl h;
    KGB_CUDA_CHECK(cudaMemcpyAsync(&h, w.ctl, sizeof(BrcCtl), cudaMemcpyDeviceToHost, st));
    KGB_CUDA_CHECK(cudaStreamSynchronize(st));
    KGB_REQUIRE(h.err != 1, "kgb_brc_parse: malformed row (expected name;[-]d[d].d)");
    KGB_REQUIRE(h.err != 2, "kgb_brc_parse: more than %lld distinct names (or a name longer than 255 bytes)",
                (long long)dict_capacity);
    KGB_REQUIRE(h.err == 0, "kgb_brc_parse: dictionary slot was never published");
    *n_names_out_host = (int64_t)h.n_names;
    return KGB_OK;
}
