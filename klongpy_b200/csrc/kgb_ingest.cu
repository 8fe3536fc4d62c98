// 1brc text ingest: `name;[-]d[d].d\n` lines -> (dense station id, float64 temperature) columns + station dictionary.
// Replaces the per-line Python loop of the reference's examples/1brc/par.py:57-88 and the string grouping of
// worker.kg:5.  Byte work, HBM-bound by design:
//
//   brc_mask_kernel   streams the text once with 2 x 128-bit loads per lane (32 bytes -> one mask word, bit j = byte j
//                     is '\n'), one coalesced 128-byte store of mask words per warp and per 1024 bytes, plus (when a
//                     warp saw a newline) one RED of its count into the counter of its 65536-byte tile — the same
//                     mask / tile-counter layout as the stream compaction in kgb_select.cuh, whose offsets kernel
//                     is reused.  Traffic: the text once + n/8 bytes.
//   brc_parse_kernel  one CTA per 65536-byte tile with newlines: block scan of the per-thread newline counts gives
//                     every line its row number; the thread owning a line's '\n' finds the line start through the
//                     mask (no byte search), parses the measurement backwards from the newline in fixed point,
//                     hashes the name (FNV-1a, 64 bit) and resolves it in an L2-resident open-addressing table.
//                     A name already present costs one 32-byte slot read and a byte compare against the stored
//                     occurrence; a new name takes the slot with one CAS, draws the next dense id and publishes
//                     {span, id}.  Traffic: the text a second time (L2/L1 friendly: neighbouring threads share
//                     sectors) + 16 bytes of output per row.
#include <cstdlib>

#include "kgb_select.cuh"

namespace kgb {

constexpr int BRC_LOG2_SLOTS = 18;
constexpr u64 BRC_SLOTS = 1ull << BRC_LOG2_SLOTS;
constexpr u64 BRC_UNSET = ~0ull;
struct __align__(32) BrcSlot {
    u64 hash;  // 0 = free
    u64 span;  // (byte offset << 8) | length of one occurrence of the name; BRC_UNSET until published
    u64 id;    // dense id in first-arrival order; BRC_UNSET until published
    u64 pad;
};
struct BrcCtl {
    unsigned long long n_names;
    int err;  // 1 malformed row, 2 dictionary full, 3 a slot was never published
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

// 4 bytes -> 4 mask bits (bit i = byte i == '\n')
__device__ __forceinline__ u32 nl4(u32 x) {
    const u32 m = __vcmpeq4(x, 0x0a0a0a0au) & 0x80808080u;  // 0x80 per matching byte
    return ((m >> 7) | (m >> 14) | (m >> 21) | (m >> 28)) & 0xfu;
}

__global__ void __launch_bounds__(256) brc_mask_kernel(const unsigned char* __restrict__ text, i64 n, u32* __restrict__ masks,
                                                        unsigned long long* __restrict__ tilecnt) {
    const int lane = threadIdx.x & 31;
    const i64 nsteps = (n + SEL_WARP_ELEMS - 1) / SEL_WARP_ELEMS;
    const i64 wstride = (i64)gridDim.x * 8;
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    for (i64 step = (i64)blockIdx.x * 8 + (threadIdx.x >> 5); step < nsteps; step += wstride) {
        const i64 b0 = step * SEL_WARP_ELEMS + lane * 32;  // this lane's 32 bytes -> mask word `lane` of the step
        u32 word = 0;
        if (aligned && b0 + 32 <= n) {
            const uint4 q0 = __ldcs(reinterpret_cast<const uint4*>(text + b0));
            const uint4 q1 = __ldcs(reinterpret_cast<const uint4*>(text + b0 + 16));
            word = nl4(q0.x) | (nl4(q0.y) << 4) | (nl4(q0.z) << 8) | (nl4(q0.w) << 12) | (nl4(q1.x) << 16) |
                   (nl4(q1.y) << 20) | (nl4(q1.z) << 24) | (nl4(q1.w) << 28);
        } else {
            for (int j = 0; j < 32; j++)
                if (b0 + j < n && text[b0 + j] == '\n') word |= 1u << j;
        }
        masks[step * 32 + lane] = word;
        const u32 hits = __reduce_add_sync(0xffffffffu, (u32)__popc(word));
        if (lane == 0 && hits) atomicAdd(&tilecnt[step / SEL_TILE_STEPS], (unsigned long long)hits);
    }
}

__device__ __forceinline__ u64 fnv1a(const unsigned char* __restrict__ p, int len) {
    u64 h = 0xcbf29ce484222325ull;
    for (int j = 0; j < len; j++) {
        h ^= (u64)p[j];
        h *= 0x100000001b3ull;
    }
    return h ? h : 1ull;  // 0 marks a free slot
}

// position of the newest '\n' strictly before byte position p (or -1): the line containing p starts right after it
__device__ __forceinline__ i64 prev_newline(const u32* __restrict__ masks, i64 p) {
    i64 w = p >> 5;
    const int b = (int)(p & 31);
    u32 m = b ? (masks[w] & ((1u << b) - 1u)) : 0u;
    while (m == 0) {
        if (w == 0) return -1;
        m = masks[--w];
    }
    return w * 32 + (31 - __clz(m));
}

__global__ void __launch_bounds__(SEL_THREADS) brc_parse_kernel(const unsigned char* __restrict__ text, i64 n,
                                                                 const u32* __restrict__ masks, i64 nwords,
                                                                 const unsigned long long* __restrict__ offs,
                                                                 BrcSlot* __restrict__ tab, BrcCtl* __restrict__ ctl,
                                                                 i64* __restrict__ id_out, double* __restrict__ temp_out,
                                                                 i64* __restrict__ dict_off, int* __restrict__ dict_len,
                                                                 i64 dict_capacity) {
    __shared__ u32 s_wcnt[SEL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 tile = blockIdx.x;
    const unsigned long long base = offs[tile];
    if (offs[tile + 1] == base) return;  // no line ends in this tile
    const i64 w0 = tile * SEL_TILE_WORDS + (i64)tid * SEL_WPT;
    u32 m[SEL_WPT];
#pragma unroll
    for (int j = 0; j < SEL_WPT; j++) m[j] = (w0 + j < nwords) ? masks[w0 + j] : 0u;
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < SEL_WPT; j++) cnt += __popc(m[j]);
    u32 inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_wcnt[warp] = inc;
    __syncthreads();
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; w++)
        if (w < warp) woff += s_wcnt[w];
    u64 row = base + woff + (inc - cnt);
    i64 start = -2;  // start of the next line this thread handles; -2 = not known yet
#pragma unroll 1
    for (int j = 0; j < SEL_WPT; j++) {
        u32 mm = m[j];
        while (mm) {
            const int b = __ffs(mm) - 1;
            mm &= mm - 1;
            const i64 p = (w0 + j) * 32 + b;  // this line's '\n'
            if (start == -2) start = prev_newline(masks, p) + 1;
            // ---- measurement, backwards from the newline: [-]d[d].d
            bool bad = p - start < 5;  // shortest line: "a;0.0"
            int tenths = 0;
            i64 semi = start;
            bool neg = false;
            if (!bad) {
                const int f = text[p - 1] - '0', dot = text[p - 2], d1 = text[p - 3] - '0';
                i64 q = p - 4;
                int d2 = 0;
                int c = text[q];
                if (c >= '0' && c <= '9') {
                    d2 = c - '0';
                    q--;
                    c = q >= start ? text[q] : 0;
                }
                if (c == '-') {
                    neg = true;
                    q--;
                    c = q >= start ? text[q] : 0;
                }
                bad = (unsigned)f > 9u || (unsigned)d1 > 9u || dot != '.' || c != ';' || q <= start;
                tenths = (d2 * 10 + d1) * 10 + f;
                semi = q;
            }
            if (bad) {
                ctl->err = 1;
            } else {
                double v = (double)tenths / 10.0;  // correctly rounded quotient of two exact integers == float("d.d")
                temp_out[row] = neg ? -v : v;
                // ---- station: hash the name, find or claim its slot
                const int len = (int)(semi - start);
                const unsigned char* name = text + start;
                const u64 h = fnv1a(name, len);
                const u64 span = ((u64)start << 8) | (u64)(len & 255);
                u64 slot = (h ^ (h >> 29)) & (BRC_SLOTS - 1);
                u64 id = BRC_UNSET;
                for (u64 probe = 0; probe < BRC_SLOTS && id == BRC_UNSET; probe++) {
                    u64 cur = ld_relaxed_u64(&tab[slot].hash);
                    if (cur == 0) {
                        cur = atomicCAS((unsigned long long*)&tab[slot].hash, 0ull, (unsigned long long)h);
                        if (cur == 0) {  // ours: draw the next dense id and publish {span, id}
                            const u64 nid = atomicAdd(&ctl->n_names, 1ull);
                            if ((i64)nid < dict_capacity && len <= 255) {
                                dict_off[nid] = start;
                                dict_len[nid] = len;
                            } else {
                                ctl->err = 2;
                            }
                            st_relaxed_u64(&tab[slot].span, span);
                            __threadfence();
                            st_relaxed_u64(&tab[slot].id, nid);
                            id = nid;
                            break;
                        }
                    }
                    if (cur == h) {
                        u64 sid = ld_relaxed_u64(&tab[slot].id);
                        for (u32 spins = 0; sid == BRC_UNSET; spins++) {  // the claimant publishes right after its CAS
                            __nanosleep(32);
                            sid = ld_relaxed_u64(&tab[slot].id);
                            if (spins > (1u << 22)) {
                                ctl->err = 3;
                                break;
                            }
                        }
                        if (sid == BRC_UNSET) break;
                        __threadfence();
                        const u64 sp = ld_relaxed_u64(&tab[slot].span);
                        if ((int)(sp & 255) == len) {
                            const unsigned char* other = text + (sp >> 8);
                            bool same = true;
                            for (int t = 0; t < len; t++) same &= other[t] == name[t];
                            if (same) id = sid;
                        }
                    }
                    slot = (slot + 1) & (BRC_SLOTS - 1);
                }
                if (id == BRC_UNSET && ctl->err == 0) ctl->err = 2;
                id_out[row] = (i64)id;
            }
            row++;
            start = p + 1;
        }
    }
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
    const long long steps = (long long)((n_bytes + SEL_WARP_ELEMS - 1) / SEL_WARP_ELEMS);
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
                             int64_t* dict_off_out, int32_t* dict_len_out, int64_t dict_capacity,
                             int64_t* n_names_out_host, void* workspace, size_t workspace_bytes, void* stream) {
    KGB_REQUIRE(current_device() >= 0, "kgb_init() has not been called on this thread");
    KGB_REQUIRE(n_names_out_host && n_bytes >= 0 && rows >= 0 && dict_capacity >= 0, "kgb_brc_parse: bad arguments");
    *n_names_out_host = 0;
    if (n_bytes == 0 || rows == 0) return KGB_OK;
    KGB_REQUIRE(text && id_out && temp_out && dict_off_out && dict_len_out && workspace &&
                    workspace_bytes >= carve_brc(nullptr, n_bytes).total,
                "kgb_brc_parse: null pointer or workspace too small");
    KGB_REQUIRE(dict_capacity <= (int64_t)(BRC_SLOTS / 2), "kgb_brc_parse: dictionary capacity above %lld",
                (long long)(BRC_SLOTS / 2));
    cudaStream_t st = (cudaStream_t)stream;
    BrcWs w = carve_brc(workspace, n_bytes);
    KGB_CUDA_CHECK(cudaMemsetAsync(w.ctl, 0, 128, st));
    KGB_CUDA_CHECK(cudaMemsetAsync(w.tab, 0, (size_t)BRC_SLOTS * sizeof(BrcSlot), st));
    // span / id words start as BRC_UNSET: one strided memset would do, a second full one is simpler and 8 MB is 2 us
    KGB_CUDA_CHECK(cudaMemset2DAsync((char*)w.tab + 8, sizeof(BrcSlot), 0xff, 16, (size_t)BRC_SLOTS, st));
    const i64 ntiles = (i64)select_tiles(n_bytes);
    brc_parse_kernel<<<(unsigned)ntiles, SEL_THREADS, 0, st>>>(text, n_bytes, w.masks, (i64)select_mask_words(n_bytes),
                                                             w.tilecnt, w.tab, w.ctl, (i64*)id_out, temp_out,
                                                             (i64*)dict_off_out, dict_len_out, dict_capacity);
    KGB_CUDA_CHECK(cudaGetLastError());
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
