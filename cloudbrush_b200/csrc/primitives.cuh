// primitives.cuh -- device-wide exclusive scan and LSD radix sort (hand-written, sm_100a).
#pragma once
#include "common.cuh"

namespace cb {

struct LaunchCounter {
  int64_t n = 0;
};

// out[i] = sum_{j<i} in[j]  (in == out allowed).  Returns nothing; total = out[n-1] + in[n-1].
void exclusive_scan_u32(const u32* in, u32* out, size_t n, cudaStream_t s, LaunchCounter& lc);

// LSD radix sort on key bits [begin_bit, end_bit), 8-bit digits, stable.
// keys/vals are ping-ponged between the given buffers; the return value is 0 when the sorted
// data ends in (keys_a, vals_a) and 1 when it ends in (keys_b, vals_b).  vals_* may be null
// (keys only).  n < 2^32.
int radix_sort(u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, size_t n, int begin_bit, int end_bit,
               cudaStream_t s, LaunchCounter& lc);

// bytes moved per element by one pass of radix_sort (read for histogram + read + write)
inline size_t radix_pass_bytes(bool has_vals) { return 8 + (has_vals ? 24 : 16); }

}  // namespace cb
