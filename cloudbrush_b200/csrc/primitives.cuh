// primitives.cuh -- device-wide exclusive scan and LSD radix sort (hand-written, sm_100a).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace cb {

// Small results (counters, sizes, flags) come back to the host through a one-block kernel that stores them into
// mapped pinned host memory, not through the copy engine: a device-to-host copy queues behind the multi-megabyte
// transfers a pipelined caller keeps in flight (cb_prefetch_sfa_buffer, cb_export_*_begin) and stalled every
// stage group by the length of such a transfer (tools/e2e_probe.py).
struct ReadBack {
  static constexpr size_t kCap = 8192;
  char* host = nullptr;  // mapped pinned staging area
  char* dev = nullptr;   // the same memory as the device sees it
  size_t used = 0;
  int64_t* launches = nullptr;  // the owner's launch count (the read-back kernel is one of the library's kernels)
  struct Item {
    void* dst;
    size_t off, bytes;
  };
  std::vector<Item> items;
  // enqueue: dst[0..bytes) = dev_src[0..bytes) once sync() has returned
  void get(void* dst, const void* dev_src, size_t bytes, cudaStream_t s);
  // the other direction: dev_dst[0..bytes) = host_src[0..bytes) (read at the call), in stream order
  void put(void* dev_dst, const void* host_src, size_t bytes, cudaStream_t s);
  // waits for the stream and delivers everything enqueued
  void sync(cudaStream_t s);
  ~ReadBack();
};

// Counts kernel launches and, in profile mode, brackets the instrumented kernels with CUDA
// events on the launching stream (bench.py's roofline numbers come from these).
struct KernelStat {
  std::vector<int> ev;  // indices into the event pool, (begin, end) pairs
};
struct LaunchCounter {
  int64_t n = 0;
  bool profile = false;
  cudaStream_t stream = nullptr;
  ReadBack rb;  // (lives here because every primitive already receives the launch counter)
  LaunchCounter() { rb.launches = &n; }
  LaunchCounter(const LaunchCounter&) = delete;
  LaunchCounter& operator=(const LaunchCounter&) = delete;
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

// ---------------------------------------------------------------------------------------------
// Hash partition of k-mer tuples into buckets (the shuffle replacement of the overlap path).
// Grouping tuples by k-mer needs no total order: two scatter passes on the top bits of a hash of
// the canonical k-mer leave buckets of a few hundred tuples, which one CTA groups in shared
// memory (stage_overlap.cu: bucket_join_kernel).  The scatter is NOT stable: every tile reserves
// its output ranges with one atomic add per bin on the bin's global cursor, so tiles never wait
// for one another (the LSD passes above spend a third of their time in the decoupled look-back).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 kmer_hash(u64 x) {
  x *= 0x9E3779B97F4A7C15ull;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  return x;  // the TOP bits are the well mixed ones
}

struct BucketPlan {
  int b1 = 1, b2 = 1;   // digit bits of the two scatter passes
  int shift1 = 0, shift2 = 0, lshift = 0;  // digit = (hash >> shift) & mask; lshift: the 10-bit local digit
  u32 NB = 4;           // buckets = 2^(b1 + b2)
  bool fits = true;     // false: more tuples than two passes can cut into buckets the bucket kernel holds
};
constexpr int kLocalBits = 10;
// buckets of 320..640 tuples on average (the bucket kernel holds 1024)
BucketPlan make_bucket_plan(size_t n_tuples);
BucketPlan make_bucket_plan_bits(int nbits);  // at least 2^nbits buckets (nbits <= 20)

// Tuples (ka, va)[n_slots] -- slots whose key is ~0 are skipped -- end up bucketed in (ka, va) again
// ((kb, vb) is the intermediate buffer); bstart[NB + 1] receives the bucket offsets and *d_total the
// number of tuples kept.  d_cnt1 (2^b1 level-1 digit counts) may come from the producer of the tuples
// (keygen counts while it writes); nullptr = counted here with one more read of the keys.
void bucket_partition(u64* ka, u32* va, u64* kb, u32* vb, size_t n_slots, u64 kmask, const BucketPlan& pl,
                      const u32* d_cnt1, u32* bstart, u32* d_total, cudaStream_t s, LaunchCounter& lc);

// bytes moved per element by one pass of radix_sort (read for histogram + read + write)
inline size_t radix_pass_bytes(bool has_vals) { return 8 + (has_vals ? 24 : 16); }

}  // namespace cb
