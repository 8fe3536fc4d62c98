// Internal interface of the onesweep radix sort (kgb_sort.cu), used by grade, group and unique.
#pragma once
#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

size_t sort_ws_bytes(i64 n);

// Stable ascending LSD radix sort of (order-encoded key, position).  idx_out: int64[n] permutation
// (written reversed when `descending`).  keys_sorted_out: optional typed sorted keys.  enc_sorted_out:
// optional order-encoded (u64) sorted keys.  Synchronises the stream for n >= 2^21 (packed path, kgb_sort_packed.cuh).
// `grp` (optional, for `=a`): when the keys' value range is small enough for the packed path's full histogram, the group
// start offsets and their count are written as well and grp->done is set — the caller skips its boundary pass (and
// enc_sorted_out, wanted only for that pass, is then NOT written).
struct SortGroupOut {
    i64* starts_out;   // [G + 1], ascending key order, last entry n
    i64* ngroups_out;  // device int64[1]
    int done;
};
int sort_pairs(const void* keys, int dtype, i64 n, int descending, i64* idx_out, void* keys_sorted_out,
               u64* enc_sorted_out, void* ws, size_t ws_bytes, cudaStream_t st, SortGroupOut* grp = nullptr);

}  // namespace kgb
