// primitives.cuh -- device-wide exclusive scan and LSD radix sort (hand-written, sm_100a).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace cb {

// Counts kernel launches and, in profile mode, brackets the instrumented kernels with CUDA
// events on the launching stream (bench.py's roofline numbers come from these).
struct KernelStat {
  std::vector<int> ev;  // indices into the event pool, (begin, end) pairs
};
struct LaunchCounter {
  int64_t n = 0;
  bool profile = false;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  std::map<std::string, KernelStat> stats;
  int next_event() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return (int)used++;
  }
  void begin(const char* name) {
    if (!profile) return;
    int i = next_event();
    stats[name].ev.push_back(i);
    cudaEventRecord(pool[i], stream);
  }
  void end(const char* name) {
    if (!profile) return;
    int i = next_event();
    stats[name].ev.push_back(i);
    cudaEventRecord(pool[i], stream);
  }
  void reset() {
    stats.clear();
    used = 0;
  }
  // total milliseconds and launch count of one instrumented kernel (synchronises)
  bool query(const std::string& name, double& ms, int64_t& launches) {
    auto it = stats.find(name);
    if (it == stats.end()) return false;
    ms = 0;
    launches = 0;
    const std::vector<int>& ev = it->second.ev;
    for (size_t i = 0; i + 1 < ev.size(); i += 2) {
      cudaEventSynchronize(pool[ev[i + 1]]);
      float f = 0;
      cudaEventElapsedTime(&f, pool[ev[i]], pool[ev[i + 1]]);
      ms += f;
      launches++;
    }
    return true;
  }
  ~LaunchCounter() {
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
  }
};

// out[i] = sum_{j<i} in[j]  (in == out allowed).  Returns nothing; total = out[n-1] + in[n-1].
void exclusive_scan_u32(const u32* in, u32* out, size_t n, cudaStream_t s, LaunchCounter& lc);

// LSD radix sort on key bits [begin_bit, end_bit), 8-bit digits, stable.
// keys/vals are ping-ponged between the given buffers; the return value is 0 when the sorted
// data ends in (keys_a, vals_a) and 1 when it ends in (keys_b, vals_b).  vals_* may be null
// (keys only).  n < 2^32.
int radix_sort(u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, size_t n, int begin_bit, int end_bit,
               cudaStream_t s, LaunchCounter& lc, const char* tag = "misc");

// bytes moved per element by one pass of radix_sort (read for histogram + read + write)
inline size_t radix_pass_bytes(bool has_vals) { return 8 + (has_vals ? 24 : 16); }

}  // namespace cb
