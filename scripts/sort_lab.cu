// Development harness for the packed radix pass (klongpy_b200/csrc/kgb_sort_packed.cuh): runs the whole packed sort on
// the three key sets of BASELINE config 3 at n keys, checks the permutation against cub::DeviceRadixSort (stable) and
// prints per-kernel times.  Geometry comes from -DKGB_PK_CT/-DKGB_PK_KPT/-DKGB_PK_LBT/-DKGB_PK_OCC, so one binary per
// variant (scripts/build_sort_lab.sh builds them in parallel; scripts/run_sort_lab.sh runs them on the GPU box).
//   sort_lab [n=100000000] [reps=5]
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "kgb_sort_packed.cuh"

namespace kgb {  // the library's error plumbing is not linked into the lab
void set_error(const char*, ...) {}
int current_device() { return 0; }
int sm_count() { return 148; }
}  // namespace kgb

using namespace kgb;
using namespace kgb::pk;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int CT = KGB_PK_CT, KPT = KGB_PK_KPT, LBT = KGB_PK_LBT, OCC = KGB_PK_OCC, T = CT * KPT;

__global__ void enc_iota(const i64* k, u64* enc, u32* pos, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        enc[i] = key_from_i64(k[i]);
        pos[i] = (u32)i;
    }
}
__global__ void compare(const i64* got, const u32* ref, u64 n, unsigned long long* bad, unsigned long long* first_bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (got[i] != (i64)ref[i]) {
            atomicAdd(bad, 1ull);
            atomicMin(first_bad, (unsigned long long)i);
        }
    }
}

__host__ __device__ inline u64 mix64(u64 x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// pseudo-random permutation of [0, n): 4-round Feistel network on 2*hb bits, cycle-walked into range
__device__ inline u64 feistel(u64 x, int hb, u64 seed) {
    const u64 m = (1ull << hb) - 1;
    u64 l = x >> hb, r = x & m;
    for (int k = 0; k < 4; k++) {
        const u64 t = l ^ (mix64(r + seed * 977 + k) & m);
        l = r;
        r = t;
    }
    return (l << hb) | r;
}
__global__ void gen_keys(i64* k, u64 n, int ds, int hb) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        i64 v;
        if (ds == 0) {
            u64 x = i;
            do x = feistel(x, hb, 7); while (x >= n);
            v = (i64)x;
        } else if (ds == 1) {
            v = (i64)(mix64(i) % 10000);
        } else if (ds == 2) {
            v = (i64)mix64(i);
        } else if (ds == 3) {
            v = (i64)(mix64(i) >> 24) - (1ll << 39);
        } else {
            v = (i64)(mix64(i) % 3) * 1000003ll - 7;
        }
        k[i] = v;
    }
}

struct Times {
    float total, sample, hist, pass[MAXP], fix;
};

static u64* g_stats = nullptr;  // device, MAXP * 4 words
static int run_sort(const i64* d_keys, u64 n, i64* d_idx, void* ws, u64* buf0, u64* buf1, int sms, Times* tm, Plan* plan_out,
                    int* attempts, GroupOut* grp) {
    cudaStream_t st = 0;
    Ws w = carve_ws<T>(ws, n);
    LabInfo lab{};
    for (auto& e : lab.ev) CK(cudaEventCreate(&e));
    lab.stats = KGB_PK_STATS ? g_stats : nullptr;
    const int rc = packed_sort_run<CT, KPT, LBT, OCC>(d_keys, KGB_I64, n, 0, d_idx, nullptr, nullptr, w, buf0, buf1, sms, st, grp, &lab);
    if (rc < 0) {
        printf("packed_sort_run: %s\n", cudaGetErrorString((cudaError_t)(-rc)));
        exit(2);
    }
    cudaEvent_t end;
    CK(cudaEventCreate(&end));
    CK(cudaEventRecord(end, st));
    CK(cudaEventSynchronize(end));
    const Plan& p = lab.plan;
    *plan_out = p;
    *attempts = lab.attempts;
    CK(cudaEventElapsedTime(&tm->total, lab.ev[0], end));
    if (lab.attempts == 1 && p.npass > 0) {
        CK(cudaEventElapsedTime(&tm->sample, lab.ev[0], lab.ev[1]));
        CK(cudaEventElapsedTime(&tm->hist, lab.ev[1], lab.ev[2]));
        for (int i = 0; i < p.npass; i++) CK(cudaEventElapsedTime(&tm->pass[i], lab.ev[2 + i], lab.ev[3 + i]));
        CK(cudaEventElapsedTime(&tm->fix, lab.ev[2 + p.npass], lab.ev[3 + p.npass]));
    }
    for (auto& e : lab.ev) CK(cudaEventDestroy(e));
    CK(cudaEventDestroy(end));
    return rc == 1 ? 0 : 1;
}

int main(int argc, char** argv) {
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 100000000ull;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("sort_lab CT=%d KPT=%d LBT=%d OCC=%d ATOM=%d LB=%d STATS=%d NOLB=%d  T=%d  n=%llu sms=%d\n", CT, KPT, LBT, OCC, KGB_PK_ATOM, KGB_PK_LB, KGB_PK_STATS, KGB_PK_NOLB, T,
           (unsigned long long)n, sms);
    {
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, pass_kernel<9, CT, KPT, LBT, 0, 8, OCC>));
        printf("pass<9,mid>: %d regs, %zu B local, smem %zu B\n", fa.numRegs, fa.localSizeBytes, sizeof(PassSmem<9, CT, KPT, LBT>));
    }
    i64 *d_keys, *d_idx;
    u64 *d_enc, *d_enc2, *buf0, *buf1;
    u32 *d_pos, *d_ref;
    void *ws, *cub_ws = nullptr;
    unsigned long long* d_bad;
    CK(cudaMalloc(&d_keys, n * 8));
    CK(cudaMalloc(&d_idx, n * 8));
    CK(cudaMalloc(&d_enc, n * 8));
    CK(cudaMalloc(&d_enc2, n * 8));
    CK(cudaMalloc(&buf0, n * 8));
    CK(cudaMalloc(&buf1, n * 8));
    CK(cudaMalloc(&d_pos, n * 4));
    CK(cudaMalloc(&d_ref, n * 4));
    CK(cudaMalloc(&d_bad, 16));
    CK(cudaMalloc(&g_stats, MAXP * 4 * 8));
    const size_t wsb = carve_ws<T>(nullptr, n).total;
    CK(cudaMalloc(&ws, wsb));
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, d_enc, d_enc2, d_pos, d_ref, (long long)n);
    CK(cudaMalloc(&cub_ws, cub_bytes));
    const char* names[] = {"perm", "10k", "rand64", "rand40", "dup3"};
    for (int ds = 0; ds < 5; ds++) {
        {
            int hb = 1;
            while ((1ull << (2 * hb)) < n) hb++;
            gen_keys<<<1184, 256>>>(d_keys, n, ds, hb);
            CK(cudaGetLastError());
        }
        enc_iota<<<1184, 256>>>(d_keys, d_enc, d_pos, n);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float cub_ms = 1e9f;
        for (int r = 0; r < 2; r++) {
            CK(cudaEventRecord(e0));
            cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, d_enc, d_enc2, d_pos, d_ref, (long long)n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            cub_ms = std::min(cub_ms, ms);
        }
        Times best{};
        best.total = 1e9f;
        Plan p{};
        int over = 0, attempts = 0;
        i64 *d_starts, *d_ng;
        CK(cudaMalloc(&d_starts, (n + 1) * 8));
        CK(cudaMalloc(&d_ng, 8));
        GroupOut grp{d_starts, d_ng, 0};
        for (int r = 0; r < reps + 1; r++) {
            Times tm{};
            CK(cudaMemset(g_stats, 0, MAXP * 4 * 8));
            over = run_sort(d_keys, n, d_idx, ws, buf0, buf1, sms, &tm, &p, &attempts, ds == 1 ? &grp : nullptr);
            if (r > 0 && tm.total < best.total) best = tm;
        }
        unsigned long long bad[2] = {0, ~0ull};
        CK(cudaMemcpy(d_bad, bad, 16, cudaMemcpyHostToDevice));
        compare<<<1184, 256>>>(d_idx, d_ref, n, d_bad, d_bad + 1);
        CK(cudaMemcpy(bad, d_bad, 16, cudaMemcpyDeviceToHost));
        printf("%-7s npass=%d bits=[", names[ds], p.npass);
        for (int i = 0; i < p.npass; i++) printf("%d%s", p.bits[i], i + 1 < p.npass ? "," : "");
        printf("] lo=%d live=%d pb=%d prefix=%d attempts=%d | total %.3f ms (sample+sync %.3f, hist %.3f, passes", p.lo, p.live, p.pb,
               p.prefix, attempts, best.total, best.sample, best.hist);
        for (int i = 0; i < p.npass; i++) printf(" %.3f", best.pass[i]);
        printf(", fixup+sync %.3f) cub %.3f ms | mismatches %llu first %lld fallback=%d", best.fix, cub_ms, bad[0],
               bad[0] ? (long long)bad[1] : -1ll, over);
        if (ds == 1) {
            i64 ng = -1, first3[3] = {-1, -1, -1}, lastv = -1;
            CK(cudaMemcpy(&ng, d_ng, 8, cudaMemcpyDeviceToHost));
            if (grp.done && ng > 2) {
                CK(cudaMemcpy(first3, d_starts, 24, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&lastv, d_starts + ng, 8, cudaMemcpyDeviceToHost));
            }
            printf(" | group done=%d G=%lld starts[0..2]=%lld,%lld,%lld starts[G]=%lld", grp.done, (long long)ng, (long long)first3[0],
                   (long long)first3[1], (long long)first3[2], (long long)lastv);
        }
        printf("\n");
        CK(cudaFree(d_starts));
        CK(cudaFree(d_ng));
#if KGB_PK_STATS
        {
            u64 hs[MAXP * 4];
            CK(cudaMemcpy(hs, g_stats, sizeof hs, cudaMemcpyDeviceToHost));
            for (int i = 0; i < p.npass; i++)
                printf("   pass %d: look-back rounds/tile %.2f, steps/tile %.2f (chain 0 of thread 0), write-out wait polls/tile %.1f\n", i,
                       (double)hs[i * 4] / (double)(hs[i * 4 + 2] ? hs[i * 4 + 2] : 1), (double)hs[i * 4 + 1] / (double)(hs[i * 4 + 2] ? hs[i * 4 + 2] : 1),
                       (double)hs[i * 4 + 3] / (double)(hs[i * 4 + 2] ? hs[i * 4 + 2] : 1));
        }
#endif
        fflush(stdout);
    }
    return 0;
}
