// common.cuh -- shared host/device helpers of the CloudBrush B200 path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <stdexcept>
#include <unordered_map>

#include "../../include/cloudbrush_gpu.h"

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned long long ull;

namespace cb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CB_CUDA(expr)                                                                           \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      throw cb::Error(_e == cudaErrorMemoryAllocation ? CB_E_NOMEM : CB_E_CUDA,                 \
                      std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ +  \
                          ":" + std::to_string(__LINE__));                                      \
  } while (0)

// Per-context caching device allocator.  Every stage of the path allocates the same buffer
// sizes run after run, so freed blocks are kept in a size-ordered free list and handed out again
// (all work of a context is on one stream, so reuse right after free is stream-ordered and
// safe).  cudaMalloc/cudaFree are only paid while the cache warms up.  (The driver's
// cudaMallocAsync pool showed sporadic ~0.8 s remapping stalls on this workload.)
struct Arena {
  std::multimap<size_t, void*> free_;
  std::unordered_map<void*, size_t> size_;
  size_t reserved = 0, in_use = 0;
  static size_t round_up(size_t b) {
    if (b < 512) return 512;
    size_t g = b >= (8u << 20) ? (2u << 20) : 512;
    return (b + g - 1) / g * g;
  }
  void* alloc(size_t bytes) {
    bytes = round_up(bytes);
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= bytes + bytes / 8 + 4096) {
      void* p = it->second;
      in_use += it->first;
      free_.erase(it);
      return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      trim();
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw cb::Error(CB_E_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    size_[p] = bytes;
    reserved += bytes;
    in_use += bytes;
    return p;
  }
  void free(void* p) {
    auto it = size_.find(p);
    if (it == size_.end()) return;
    in_use -= it->second;
    free_.insert({it->second, p});
  }
  void trim() {  // give every cached block back to the driver
    for (auto& kv : free_) {
      cudaFree(kv.second);
      reserved -= kv.first;
      size_.erase(kv.second);
    }
    free_.clear();
  }
  ~Arena() {
    trim();
    for (auto& kv : size_) cudaFree(kv.first);
  }
};

// the arena of the context whose API call is running on this thread (set by CB_API_BEGIN)
inline Arena*& current_arena() {
  static thread_local Arena* a = nullptr;
  return a;
}

// Device buffer owned by the current context's arena.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  Arena* a = nullptr;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), a(o.a) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; a = o.a; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count, cudaStream_t) {
    release();
    a = current_arena();
    if (!a) throw cb::Error(CB_E_INVALID, "device allocation outside an API call");
    n = count;
    if (count) p = (T*)a->alloc(count * sizeof(T));
  }
  void borrow(T* ptr, size_t count) {  // a view of memory owned elsewhere (a == nullptr: never freed here)
    release();
    p = ptr;
    n = count;
    a = nullptr;
  }
  void release() {
    if (p && a) a->free(p);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

inline int div_up(size_t a, size_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// 2-bit DNA words.  A read of L bases is ceil(L/32) u64 words, base i in word i/32 at bits
// [63-2*(i%32)-1, 63-2*(i%32)] (MSB first), codes A=0 C=1 G=2 T=3, so that unsigned word-wise
// comparison equals String.compareTo on the ACGT strings the reference compares
// (e.g. MatchPrefix.java:122, GenNonContainedReads.java:193).  Unused low bits are zero.
// ---------------------------------------------------------------------------------------------

// reverse the order of the 32 2-bit groups of x and complement every base (A<->T, C<->G)
__host__ __device__ __forceinline__ u64 revcomp_word(u64 x) {
#ifdef __CUDA_ARCH__
  u64 y = __brevll(x);
#else
  u64 y = 0;
  for (int i = 0; i < 64; i++) y |= ((x >> i) & 1ull) << (63 - i);
#endif
  y = ((y & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((y & 0x5555555555555555ull) << 1);
  return ~y;
}

// shift an NW-word base string left by nb bases (0 <= nb <= 32*NW), zero fill
template <int NW>
__host__ __device__ __forceinline__ void shl_bases(u64 (&a)[NW], int nb) {
  // register arrays must only be indexed with compile-time constants (a dynamic index sends
  // the array to local memory): the word shift is a binary cascade of predicated moves
  const int ws = nb >> 5, bs = (nb & 31) * 2;
#pragma unroll
  for (int step = 1; step < 2 * NW; step <<= 1) {
    if (ws & step) {
#pragma unroll
      for (int i = 0; i < NW; i++) a[i] = (i + step < NW) ? a[(i + step < NW) ? i + step : 0] : 0ull;
    }
  }
  if (bs) {
#pragma unroll
    for (int i = 0; i < NW; i++) a[i] = (a[i] << bs) | ((i + 1 < NW) ? (a[(i + 1 < NW) ? i + 1 : 0] >> (64 - bs)) : 0ull);
  }
}

// out = reverse complement of the L-base string in `in`
template <int NW>
__host__ __device__ __forceinline__ void revcomp_read(const u64 (&in)[NW], u64 (&out)[NW], int L) {
#pragma unroll
  for (int i = 0; i < NW; i++) out[i] = revcomp_word(in[NW - 1 - i]);
  shl_bases<NW>(out, NW * 32 - L);
}

// keep the first nb bases, zero the rest
template <int NW>
__host__ __device__ __forceinline__ void keep_prefix(u64 (&a)[NW], int nb) {
#pragma unroll
  for (int i = 0; i < NW; i++) {
    int lo = i * 32;
    if (nb <= lo) a[i] = 0;
    else if (nb < lo + 32) a[i] &= ~0ull << (64 - 2 * (nb - lo));
  }
}

// do the first nb bases of a and b agree?
template <int NW>
__host__ __device__ __forceinline__ bool prefix_equal(const u64 (&a)[NW], const u64 (&b)[NW], int nb) {
  u64 diff = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    int lo = i * 32;
    u64 x = a[i] ^ b[i];
    if (nb <= lo) x = 0;
    else if (nb < lo + 32) x &= ~0ull << (64 - 2 * (nb - lo));
    diff |= x;
  }
  return diff == 0;
}

// the K-mer (K <= 31) starting at base `pos`, right aligned in 2K bits
template <int NW>
__host__ __device__ __forceinline__ u64 extract_kmer(const u64 (&a)[NW], int pos, int K) {
  int w = pos >> 5, off = (pos & 31) * 2;
  u64 hi = 0, lo = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    if (i == w) hi = a[i];
    if (i == w + 1) lo = a[i];
  }
  u64 top = off ? ((hi << off) | (lo >> (64 - off))) : hi;
  return top >> (64 - 2 * K);
}

__host__ __device__ __forceinline__ u64 revcomp_kmer(u64 kmer, int K) {
  return revcomp_word(kmer) >> (64 - 2 * K);
}

// lexicographic compare of NW-word strings (same length): -1/0/1
template <int NW>
__host__ __device__ __forceinline__ int cmp_words(const u64 (&a)[NW], const u64 (&b)[NW]) {
#pragma unroll
  for (int i = 0; i < NW; i++) {
    if (a[i] < b[i]) return -1;
    if (a[i] > b[i]) return 1;
  }
  return 0;
}

// one packed read row (NW even: NW/2 aligned 16-byte loads through the read-only path)
template <int NW>
__device__ __forceinline__ void load_read_row(const u64* __restrict__ reads, u32 idx, u64 (&r)[NW]) {
  const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    uint4 v = __ldg(row + i);
    r[2 * i] = (u64)v.x | ((u64)v.y << 32);
    r[2 * i + 1] = (u64)v.z | ((u64)v.w << 32);
  }
}

// Overhang storage: the first NWO 32-bit words (16 bases each, MSB first) of a left-aligned
// NW-word base string.  NWO <= 2*NW.
template <int NW>
__device__ __forceinline__ void store_overhang(u32* __restrict__ dst, const u64 (&h)[NW], int nwo) {
#pragma unroll
  for (int i = 0; i < 2 * NW; i++)
    if (i < nwo) dst[i] = (i & 1) ? (u32)h[i >> 1] : (u32)(h[i >> 1] >> 32);
}
template <int NW>
__device__ __forceinline__ void fetch_overhang(const u32* __restrict__ src, u64 (&h)[NW], int nwo) {
#pragma unroll
  for (int i = 0; i < NW; i++) {
    const u32 hi = (2 * i < nwo) ? src[2 * i] : 0u;
    const u32 lo = (2 * i + 1 < nwo) ? src[2 * i + 1] : 0u;
    h[i] = ((u64)hi << 32) | (u64)lo;
  }
}

__host__ __device__ __forceinline__ u64 mix64(u64 x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

// 16-byte id key: the id bytes big-endian, zero padded.  Unsigned (hi, lo) order equals
// String.compareTo for ASCII ids of <= 16 bytes (shorter-is-smaller on a common prefix).
struct IdKey {
  u64 hi, lo;
};
__host__ __device__ __forceinline__ bool idkey_less(const IdKey& a, const IdKey& b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}

}  // namespace cb
