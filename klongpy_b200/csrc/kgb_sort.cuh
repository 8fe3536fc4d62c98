// Internal interface of the onesweep radix sort (kgb_sort.cu), used by grade, group and unique.
#pragma once
#include "kgb_device.cuh"
#include "kgb_internal.h"

namespace kgb {

size_t sort_ws_bytes(i64 n);

// Stable ascending LSD radix sort of (order-encoded key, position).  idx_out: int64[n] permutation
// (written reversed when `descending`).  keys_sorted_out: optional typed sorted keys.  enc_sorted_out:
// optional order-encoded (u64) sorted keys.  No host synchronisation.
int sort_pairs(const void* keys, int dtype, i64 n, int descending, i64* idx_out, void* keys_sorted_out,
               u64* enc_sorted_out, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace kgb
