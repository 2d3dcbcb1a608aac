// primitives.cu -- device-wide exclusive scan and LSD radix sort for the shuffle replacement.
//
// The Hadoop shuffle of MatchPrefix (reference: JobClient.runJob at MatchPrefix.java:468 --
// HashPartitioner + sort by Text key + group) is replaced by a stable LSD radix sort of
// (canonical k-mer key, payload) tuples.
//
// What ships ("onesweep", second half of this file): ONE histogram read of the keys per sort
// (rs_hist_all, all digit positions at once), then one kernel per digit pass (rs_onesweep) whose
// per-tile digit offsets come from a chained scan with decoupled look-back; digits are 8 or 9 bits
// wide (PassPlan).  HBM traffic per pass: 12 B read + 12 B written per tuple, the algorithmic minimum.
//
// Fallback (n >= 2^30 or CB_SORT=classic), one pass over an 8-bit digit is:
//   rs_upsweep   per-tile digit histogram (shared-memory staged, warp-aggregated atomics)
//   rs_chunk_sums / rs_chunk_scan / rs_apply   column-wise exclusive scan of the tile histograms
//   rs_scatter   stable in-tile ranking with __match_any_sync, shared-memory staged so that each
//                digit run leaves the SM as one contiguous, coalesced store
// HBM traffic per pass: 8 B (keys, histogram) + 12 B read + 12 B write per tuple.
#include <cstdlib>

#include <cstring>

#include "primitives.cuh"

namespace cb {

__global__ void __launch_bounds__(128) read_back_kernel(char* __restrict__ dst, const char* __restrict__ src, int bytes) {
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

void ReadBack::get(void* dst, const void* dev_src, size_t bytes, cudaStream_t s) {
  if (!host) {
    CB_CUDA(cudaHostAlloc((void**)&host, kCap, cudaHostAllocMapped));
    CB_CUDA(cudaHostGetDevicePointer((void**)&dev, host, 0));
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (used + need > kCap) {  // (never on the path: a handful of words per round trip)
    CB_CUDA(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, s));
    return;
  }
  read_back_kernel<<<1, 128, 0, s>>>(dev + used, (const char*)dev_src, (int)bytes);
  CB_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  items.push_back({dst, used, bytes});
  used += need;
}

void ReadBack::put(void* dev_dst, const void* host_src, size_t bytes, cudaStream_t s) {
  if (!host) {
    CB_CUDA(cudaHostAlloc((void**)&host, kCap, cudaHostAllocMapped));
    CB_CUDA(cudaHostGetDevicePointer((void**)&dev, host, 0));
  }
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (used + need > kCap) {
    CB_CUDA(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, s));
    return;
  }
  memcpy(host + used, host_src, bytes);  // (the slot is not reused before the next sync())
  read_back_kernel<<<1, 128, 0, s>>>((char*)dev_dst, dev + used, (int)bytes);
  CB_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  used += need;
}

void ReadBack::sync(cudaStream_t s) {
  CB_CUDA(cudaStreamSynchronize(s));
  for (const Item& it : items) memcpy(it.dst, host + it.off, it.bytes);
  items.clear();
  used = 0;
}

ReadBack::~ReadBack() {
  if (host) cudaFreeHost(host);
}

// ------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; `total` receives the block sum.
// tmp must hold (blockDim.x / 32) u32.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* tmp, u32& total) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  u32 inc = warp_incl_scan(v);
  if (lane == 31) tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < nw ? tmp[lane] : 0;
    u32 wi = warp_incl_scan(w);
    if (lane < nw) tmp[lane] = wi - w;
    if (lane == nw - 1) tmp[nw] = wi;
  }
  __syncthreads();
  u32 r = tmp[warp] + inc - v;
  total = tmp[nw];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SC_THREADS) scan_tile_kernel(const u32* __restrict__ in, u32* __restrict__ out,
                                                               u32* __restrict__ block_sums, size_t n) {
  __shared__ u32 tmp[SC_THREADS / 32 + 1];
  size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
  u32 v[SC_ITEMS];
  u32 sum = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    sum += v[i];
  }
  u32 total;
  u32 ex = block_excl_scan(sum, tmp, total);
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
  if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) scan_add_kernel(u32* __restrict__ out, const u32* __restrict__ block_offs,
                                                              size_t n) {
  size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
  u32 o = block_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SC_ITEMS; i++)
    if (base + i < n) out[base + i] += o;
}

void exclusive_scan_u32(const u32* in, u32* out, size_t n, cudaStream_t s, LaunchCounter& lc) {
  if (n == 0) return;
  size_t nb = (n + SC_TILE - 1) / SC_TILE;
  if (nb == 1) {
    scan_tile_kernel<<<1, SC_THREADS, 0, s>>>(in, out, nullptr, n);
    lc.n++;
    CB_CUDA(cudaGetLastError());
    return;
  }
  DevBuf<u32> sums;
  sums.alloc(nb, s);
  scan_tile_kernel<<<(unsigned)nb, SC_THREADS, 0, s>>>(in, out, sums.p, n);
  lc.n++;
  CB_CUDA(cudaGetLastError());
  exclusive_scan_u32(sums.p, sums.p, nb, s, lc);
  scan_add_kernel<<<(unsigned)nb, SC_THREADS, 0, s>>>(out, sums.p, n);
  lc.n++;
  CB_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------
constexpr int RS_UP_THREADS = 256;               // upsweep: one thread per digit bin
constexpr int RS_THREADS = 512;                  // scatter: 512 threads x 8 keys = 4096 elements ranked at a time
constexpr int RS_WARPS = RS_THREADS / 32;        // (ncu on 256 x 16: 128 registers, 25 % occupancy, long
constexpr int RS_ITEMS = 8;                      //  smem dependency chains; this shape halves the chain and
constexpr int RS_SUB = RS_THREADS * RS_ITEMS;    //  doubles the resident warps)
constexpr int RS_SUBTILES = 4;
constexpr int RS_TILE = RS_SUB * RS_SUBTILES;  // 16384 elements per block

// MATCH: aggregate equal digits of a warp with __match_any_sync before the shared atomic (narrow,
// skewed digits, e.g. the 2-bit top digit of a 42-bit key); otherwise plain shared atomics, whose
// rare bank conflicts the hardware serialises at a few cycles each.
template <bool MATCH>
__global__ void __launch_bounds__(RS_UP_THREADS) rs_upsweep(const u64* __restrict__ kin, u32* __restrict__ hist, size_t n,
                                                            int shift, u32 mask) {
  __shared__ u32 h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  size_t base = (size_t)blockIdx.x * RS_TILE;
  int lane = threadIdx.x & 31;
#pragma unroll 8
  for (int i = 0; i < RS_TILE / RS_UP_THREADS; i++) {
    size_t idx = base + (size_t)i * RS_UP_THREADS + threadIdx.x;
    bool valid = idx < n;
    u32 d = valid ? (u32)((kin[valid ? idx : 0] >> shift) & mask) : 0xffffffffu;
    if (MATCH) {
      u32 peers = __match_any_sync(0xffffffffu, d);
      if (valid && lane == __ffs(peers) - 1) atomicAdd(&h[d], __popc(peers));
    } else {
      if (valid) atomicAdd(&h[d], 1u);
    }
  }
  __syncthreads();
  hist[(size_t)blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
}

// column sums of hist rows [r0, r1) of chunk blockIdx.x
__global__ void __launch_bounds__(256) rs_chunk_sums(const u32* __restrict__ hist, u32* __restrict__ chunk_sum,
                                                     int tiles, int rows_per_chunk) {
  int r0 = blockIdx.x * rows_per_chunk, r1 = min(tiles, r0 + rows_per_chunk);
  u32 s = 0;
  for (int r = r0; r < r1; r++) s += hist[(size_t)r * 256 + threadIdx.x];
  chunk_sum[(size_t)blockIdx.x * 256 + threadIdx.x] = s;
}

// one block: chunk_sum[c][d] -> exclusive base of (digit d, chunk c) in the sorted output
__global__ void __launch_bounds__(256) rs_chunk_scan(u32* __restrict__ chunk_sum, int chunks) {
  __shared__ u32 tmp[256 / 32 + 1];
  u32 run = 0;
  for (int c = 0; c < chunks; c++) {
    u32 v = chunk_sum[(size_t)c * 256 + threadIdx.x];
    chunk_sum[(size_t)c * 256 + threadIdx.x] = run;
    run += v;
  }
  u32 total;
  u32 dbase = block_excl_scan(run, tmp, total);
  for (int c = 0; c < chunks; c++) chunk_sum[(size_t)c * 256 + threadIdx.x] += dbase;
}

__global__ void __launch_bounds__(256) rs_apply(u32* __restrict__ hist, const u32* __restrict__ chunk_base, int tiles,
                                                int rows_per_chunk) {
  int r0 = blockIdx.x * rows_per_chunk, r1 = min(tiles, r0 + rows_per_chunk);
  u32 run = chunk_base[(size_t)blockIdx.x * 256 + threadIdx.x];
  for (int r = r0; r < r1; r++) {
    u32 v = hist[(size_t)r * 256 + threadIdx.x];
    hist[(size_t)r * 256 + threadIdx.x] = run;
    run += v;
  }
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const u64* __restrict__ kin, u64* __restrict__ kout,
                                                         const u32* __restrict__ vin, u32* __restrict__ vout,
                                                         const u32* __restrict__ offsets, size_t n, int shift,
                                                         u32 mask) {
  __shared__ unsigned short warp_hist[RS_WARPS][256];  // <= 8 x 32 per warp and digit, <= 4096 after the scan
  __shared__ u32 digit_base[256];
  __shared__ u32 sub_start[256];
  __shared__ u32 scan_tmp[RS_WARPS + 1];
  __shared__ u64 stage[RS_SUB];
  __shared__ unsigned char sorted_digit[RS_SUB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  if (tid < 256) digit_base[tid] = offsets[(size_t)blockIdx.x * 256 + tid];

  for (int sub = 0; sub < RS_SUBTILES; sub++) {
    size_t base = (size_t)blockIdx.x * RS_TILE + (size_t)sub * RS_SUB;
    if (base >= n) break;  // uniform across the block
    int nvalid = (int)min((size_t)RS_SUB, n - base);
    for (int t = tid; t < RS_WARPS * 256; t += RS_THREADS) (&warp_hist[0][0])[t] = 0;
    __syncthreads();

    u64 key[RS_ITEMS];
    u32 pos[RS_ITEMS];
    // warp-striped: element order inside the sub-tile is (warp, item, lane) == ascending index
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      int e = warp * (RS_ITEMS * 32) + i * 32 + lane;
      key[i] = (e < nvalid) ? kin[base + e] : ~0ull;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      int e = warp * (RS_ITEMS * 32) + i * 32 + lane;
      u32 d = (e < nvalid) ? (u32)((key[i] >> shift) & mask) : 255u;  // padding ranks last
      u32 peers = __match_any_sync(0xffffffffu, d);
      u32 prev = warp_hist[warp][d];
      __syncwarp();
      if (lane == __ffs(peers) - 1) warp_hist[warp][d] = (unsigned short)(prev + __popc(peers));
      __syncwarp();
      pos[i] = prev + __popc(peers & lt_mask);
    }
    __syncthreads();
    // per digit (thread tid): exclusive scan over the warps, then over the digits
    u32 run = 0;
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) {
        u32 c = warp_hist[w][tid];
        warp_hist[w][tid] = (unsigned short)run;
        run += c;
      }
    }
    u32 total;
    u32 dstart = block_excl_scan(run, scan_tmp, total);  // threads >= 256 contribute 0
    if (tid < 256) sub_start[tid] = dstart;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      int e = warp * (RS_ITEMS * 32) + i * 32 + lane;
      u32 d = (e < nvalid) ? (u32)((key[i] >> shift) & mask) : 255u;
      pos[i] += sub_start[d] + warp_hist[warp][d];
      stage[pos[i]] = key[i];
      sorted_digit[pos[i]] = (unsigned char)d;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < RS_ITEMS; j++) {
      int p = j * RS_THREADS + tid;
      if (p < nvalid) {
        u32 d = sorted_digit[p];
        kout[digit_base[d] + (p - sub_start[d])] = stage[p];
      }
    }
    if (HAS_VALS) {
      __syncthreads();
      u32* vstage = reinterpret_cast<u32*>(stage);
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++) {
        int e = warp * (RS_ITEMS * 32) + i * 32 + lane;
        vstage[pos[i]] = (e < nvalid) ? vin[base + e] : 0u;
      }
      __syncthreads();
#pragma unroll 4
      for (int j = 0; j < RS_ITEMS; j++) {
        int p = j * RS_THREADS + tid;
        if (p < nvalid) {
          u32 d = sorted_digit[p];
          vout[digit_base[d] + (p - sub_start[d])] = vstage[p];
        }
      }
    }
    __syncthreads();
    if (tid < 256) digit_base[tid] += run;  // padding only inflates digit 255 of the very last sub-tile
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// onesweep: one histogram read of the keys per SORT (all digit positions at once), then ONE
// kernel per digit pass.  The per-tile digit offsets come from a chained scan with decoupled
// look-back: tile t publishes its digit counts (flag 1), adds up the published values of the
// tiles before it until it meets an inclusive prefix (flag 2), then publishes its own inclusive
// prefix.  Flag and value share one 32-bit word, so a single store publishes both.  Tile ids are
// handed out by an atomic counter in start order, so a tile only ever waits for tiles that are
// already running.  HBM traffic per pass: 12 B read + 12 B written per tuple.
// ------------------------------------------------------------------------------------------
constexpr int RS_MAXPASS = 8;
constexpr u32 ST_FLAG_LOCAL = 1u << 30, ST_FLAG_INCL = 2u << 30, ST_VALUE = (1u << 30) - 1;
constexpr u32 SPIN_LIMIT = 1u << 22;
constexpr int RS_LB = 8;  // look-back loads in flight (16 and 32 measured slower: registers, L2 traffic)

// peers of this lane's digit within the warp.  MATCH.ANY's latency grows with the number of
// distinct values among the lanes (ncu: on random 8-bit digits the instructions consuming its
// result collect 23 % of the kernel's stall samples, on clustered digits 1 %); eight ballots, one
// per digit bit, cost the same whatever the values are.  `ballot` is uniform over the grid: every
// pass measures how many distinct digits of the NEXT pass a warp will see (its write-out order is
// the next pass' input order) and the next pass picks the cheaper ranking from that.
template <int RB>
__device__ __forceinline__ u32 digit_peers(u32 d, bool ballot) {
  if (!ballot) return __match_any_sync(0xffffffffu, d);
  u32 peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < RB; b++) {
    const bool bit = (d >> b) & 1u;
    const u32 bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}
constexpr u32 RS_BALLOT_ABOVE = 12;  // distinct digits per warp above which the ballots win

// one sample of the digit diversity a warp sees: adds (#distinct digits, 1) to ent[0], ent[1]
__device__ __forceinline__ void sample_diversity(u32 d, int lane, u32* ent) {
  const u32 peers = __match_any_sync(0xffffffffu, d);
  const u32 leaders = __ballot_sync(0xffffffffu, lane == __ffs(peers) - 1);
  if (lane == 0) {
    atomicAdd(&ent[0], (u32)__popc(leaders));
    atomicAdd(&ent[1], 1u);
  }
}

// Pass plan of one sort: key bits [begin, end) are cut into ceil(bits / 9) digits of 8 or 9 bits
// (42 bits of a k=21 tuple key: 8+8+8+9+9, five passes instead of six; 62 bits for k=31: seven
// instead of eight).  A 9-bit pass uses 512 bins = one per thread of the scatter kernel.
constexpr int RS_MAXRADIX = 512;
struct PassPlan {
  int npass;
  int shift[RS_MAXPASS];
  int bits[RS_MAXPASS];
};
static PassPlan make_plan(int begin_bit, int end_bit, int max_digit_bits) {
  PassPlan pl;
  const int B = end_bit - begin_bit;
  pl.npass = (B + max_digit_bits - 1) / max_digit_bits;
  const int base = B / pl.npass, rem = B % pl.npass;
  int at = begin_bit;
  for (int p = 0; p < RS_MAXPASS; p++) {
    pl.shift[p] = at;
    pl.bits[p] = p < pl.npass ? base + (p >= pl.npass - rem ? 1 : 0) : 0;  // the wider digits go last: late passes see clustered digits, whose ranking is cheap
    at += pl.bits[p];
  }
  return pl;
}

__global__ void __launch_bounds__(256) rs_hist_all_kernel(const u64* __restrict__ kin, size_t n, PassPlan pl,
                                                          u32* __restrict__ ghist, u32* __restrict__ ent0) {
  extern __shared__ u32 hsm[];  // [npass][RS_MAXRADIX]
  for (int t = threadIdx.x; t < pl.npass * RS_MAXRADIX; t += blockDim.x) hsm[t] = 0;
  __syncthreads();
  if (threadIdx.x < 32 && (size_t)blockIdx.x * blockDim.x + 32 <= n)  // first warp: digit diversity of pass 0
    sample_diversity((u32)(kin[(size_t)blockIdx.x * blockDim.x + threadIdx.x] >> pl.shift[0]) & ((1u << pl.bits[0]) - 1),
                     (int)threadIdx.x, ent0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 k = kin[i];
#pragma unroll
    for (int p = 0; p < RS_MAXPASS; p++)
      if (p < pl.npass) atomicAdd(&hsm[p * RS_MAXRADIX + ((u32)(k >> pl.shift[p]) & ((1u << pl.bits[p]) - 1))], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < pl.npass * RS_MAXRADIX; t += blockDim.x) {
    u32 v = hsm[t];
    if (v) atomicAdd(&ghist[t], v);
  }
}

// one block of RS_MAXRADIX threads: exclusive scan of every pass' digit totals (in place)
__global__ void __launch_bounds__(RS_MAXRADIX) rs_scan_hist_kernel(u32* __restrict__ ghist, int npass) {
  __shared__ u32 tmp[RS_MAXRADIX / 32 + 1];
  for (int p = 0; p < npass; p++) {
    u32 v = ghist[p * RS_MAXRADIX + threadIdx.x];
    u32 total;
    u32 ex = block_excl_scan(v, tmp, total);
    ghist[p * RS_MAXRADIX + threadIdx.x] = ex;
  }
}

// status words are read and written with gpu-scope relaxed accesses (flag and value share the
// word, so no ordering with other data is needed); `volatile` compiles to system-scope accesses
__device__ __forceinline__ u32 ld_status(const volatile u32* p) {
  u32 v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(volatile u32* p, u32 v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// -DCB_OS_TIMING: clock64 phase breakdown of the onesweep kernel (thread 0 of every CTA), read
// with cb_debug_os_prof (tools/os_timing.py).  Not part of the product build.
#ifdef CB_OS_TIMING
__device__ unsigned long long g_os_prof[16];
#define OS_T(i) { if (tid == 0) { long long now_ = clock64(); atomicAdd(&g_os_prof[i], (unsigned long long)(now_ - t_prev_)); t_prev_ = now_; } }
extern "C" void cb_debug_os_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_os_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_os_prof, z, sizeof(z)); }
}
#else
#define OS_T(i)
#endif

template <bool HAS_VALS, int S, int RB>
constexpr size_t os_smem_bytes() {
  return (size_t)S * RS_SUB * (HAS_VALS ? 12 : 8) + RS_WARPS * (1 << RB) * 2 + 3 * S * (1 << RB) * 4;
}

// One logical tile = S sub-tiles of RS_SUB keys processed back to back by one CTA and staged in
// shared memory side by side; the digit counts of the sub-tiles are published together and ONE
// look-back serves them all.  (clock64 phase timing of the S = 1 shape on 69 M pairs: 37 % of a
// tile's 22.9 k cycles were the look-back -- 4.5 dependent round trips of ~1900 cycles each, the
// status loads queue behind the SM's bulk traffic -- over a window of 31 predecessor tiles.  With
// S = 2 the window halves, because tiles start half as often, and its cost is shared by twice the
// keys.)  Payloads are staged in their own buffer and leave in the same loop as the keys.
template <bool HAS_VALS, int S, int RB>
__global__ void __launch_bounds__(RS_THREADS, 2) rs_onesweep_kernel(const u64* __restrict__ kin, u64* __restrict__ kout,
                                                                    const u32* __restrict__ vin, u32* __restrict__ vout,
                                                                    u32* status, u32* tile_ctr,
                                                                    const u32* __restrict__ gbase, u32* err, u32 n,
                                                                    int shift, u32 mask, const u32* __restrict__ ent_in,
                                                                    u32* __restrict__ ent_out, int next_shift,
                                                                    u32 next_mask, int force_rank) {
  constexpr int RADIX = 1 << RB;  // 256 or 512 (= RS_THREADS) digit bins
  static_assert(RADIX <= RS_THREADS, "one thread per digit bin");
  extern __shared__ __align__(16) unsigned char os_smem[];
  u64* stage = reinterpret_cast<u64*>(os_smem);                                       // [S][RS_SUB]
  u32* vstage = reinterpret_cast<u32*>(os_smem + (size_t)S * RS_SUB * 8);             // [S][RS_SUB] (pairs only)
  unsigned short(*warp_hist)[RADIX] =
      reinterpret_cast<unsigned short(*)[RADIX]>(os_smem + (size_t)S * RS_SUB * (HAS_VALS ? 12 : 8));
  u32* cnt = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(warp_hist) + RS_WARPS * RADIX * 2);  // [S][RADIX]
  u32* sub_start = cnt + S * RADIX;                                                      // [S][256]
  u32* digit_base = sub_start + S * RADIX;                                               // [S][256]
  __shared__ u32 scan_tmp[RS_WARPS + 1];
  __shared__ u32 s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1;
#ifdef CB_OS_TIMING
  long long t_prev_ = clock64();
#endif
  if (tid == 0) s_tile = atomicAdd(tile_ctr, 1u);
  __syncthreads();
  const u32 tile = s_tile;
  const u32 base0 = tile * (u32)(S * RS_SUB);
  if (base0 >= n) return;  // uniform
  // grid-uniform choice of the ranking primitive from the diversity the previous pass measured
  const bool ballot = force_rank ? force_rank == 2 : ent_in[0] > RS_BALLOT_ABOVE * ent_in[1];
  const int e0 = warp * (RS_ITEMS * 32) + lane;  // warp-striped: element order is (warp, item, lane) == ascending index
  u32 cnt_pub = 0;

#pragma unroll 1
  for (int sub = 0; sub < S; sub++) {
    const u32 base = base0 + (u32)sub * RS_SUB;
    const int nvalid = base < n ? (int)min((u32)RS_SUB, n - base) : 0;
    for (int t = tid; t < RS_WARPS * RADIX / 2; t += RS_THREADS) reinterpret_cast<u32*>(&warp_hist[0][0])[t] = 0;
    u64 key[RS_ITEMS];
    u32 pos[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      const int e = e0 + i * 32;
      key[i] = (e < nvalid) ? kin[base + e] : ~0ull;
    }
    __syncthreads();  // histograms are zero (and the previous sub-tile is done with them)
    OS_T(0)  // tile id / zeroing / load issue
#ifdef CB_OS_TIMING
    if (tid == 0 && key[0] == 0x123456789abcdefull && key[7] == 1) atomicAdd(&g_os_prof[15], 1ull);  // wait for the loads
    OS_T(1)
#endif
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      const int e = e0 + i * 32;
      const u32 d = (e < nvalid) ? (u32)((key[i] >> shift) & mask) : (u32)(RADIX - 1);  // padding ranks last
      const u32 peers = digit_peers<RB>(d, ballot);
      const u32 prev = warp_hist[warp][d];
      __syncwarp();
      if (lane == __ffs(peers) - 1) warp_hist[warp][d] = (unsigned short)(prev + __popc(peers));
      __syncwarp();
      pos[i] = prev + __popc(peers & lt_mask);
    }
    OS_T(2)  // ranking
    __syncthreads();
    OS_T(3)  // barrier after ranking
    u32 run = 0;
    if (tid < RADIX) {
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) {
        const u32 c = warp_hist[w][tid];
        warp_hist[w][tid] = (unsigned short)run;
        run += c;
      }
      cnt[sub * RADIX + tid] = run;
      // padding was ranked as digit 255 (in the last logical tile only) and must not be published
      cnt_pub += run - (tid == RADIX - 1 ? (u32)(RS_SUB - nvalid) : 0u);
      // publish the logical tile's digit counts as soon as the last sub-tile is counted; the
      // look-back itself waits until that sub-tile is staged
      if (sub == S - 1) st_status(&status[(size_t)tile * RADIX + tid], cnt_pub | (tile == 0 ? ST_FLAG_INCL : ST_FLAG_LOCAL));
    }
    OS_T(4)  // per-digit scan over the warps + publish
    u32 total;
    const u32 dstart = block_excl_scan(run, scan_tmp, total);  // threads >= RADIX contribute 0
    if (tid < RADIX) sub_start[sub * RADIX + tid] = dstart;
    __syncthreads();
    OS_T(5)  // block scan
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      const int e = e0 + i * 32;
      const u32 d = (e < nvalid) ? (u32)((key[i] >> shift) & mask) : (u32)(RADIX - 1);
      pos[i] += sub_start[sub * RADIX + d] + warp_hist[warp][d];
      stage[sub * RS_SUB + pos[i]] = key[i];
    }
    if (HAS_VALS) {  // the key registers are dead now
      u32 val[RS_ITEMS];
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++) {
        const int e = e0 + i * 32;
        val[i] = (e < nvalid) ? vin[base + e] : 0u;
      }
#pragma unroll
      for (int i = 0; i < RS_ITEMS; i++) vstage[sub * RS_SUB + pos[i]] = val[i];
    }
    if (sub + 1 < S) __syncthreads();  // the next sub-tile zeroes the histograms this one's staging still reads
    OS_T(6)  // staging keys + payloads
  }
  if (tid < RADIX) {
    u32 excl = 0;
    if (tile != 0) {
      // decoupled look-back, RS_LB predecessors per round trip: the status words of tiles
      // p, p-1, ..., p-RS_LB+1 are loaded together (independent loads in flight) and consumed in
      // order up to the first one that is not published yet
      long long p = (long long)tile - 1;
      u32 spins = 0;
      bool done = false;
      while (!done) {
        u32 v[RS_LB];
#pragma unroll
        for (int i = 0; i < RS_LB; i++) {
          const long long q = p - i;
          v[i] = q >= 0 ? ld_status(&status[(size_t)q * RADIX + tid]) : (2u << 30);  // before tile 0: inclusive 0
        }
        int consumed = 0;
#pragma unroll
        for (int i = 0; i < RS_LB; i++) {
          if (!done && consumed == i && (v[i] >> 30) != 0) {
            excl += v[i] & ST_VALUE;
            consumed = i + 1;
            done = (v[i] >> 30) == 2;
          }
        }
        p -= consumed;
#ifdef CB_OS_TIMING
        if (tid == 0) { atomicAdd(&g_os_prof[11], 1ull); if (consumed == 0) atomicAdd(&g_os_prof[12], 1ull); atomicAdd(&g_os_prof[13], (unsigned long long)consumed); }
#endif
        if (!done && consumed == 0 && ++spins > SPIN_LIMIT) { *err = 1; break; }  // never hang the GPU on a bug
      }
      st_status(&status[(size_t)tile * RADIX + tid], (excl + cnt_pub) | ST_FLAG_INCL);
    }
    // global position of the element at sorted position p of sub-tile s: digit_base[s][d] + p
    u32 before = gbase[tid] + excl;
#pragma unroll
    for (int sub = 0; sub < S; sub++) {
      digit_base[sub * RADIX + tid] = before - sub_start[sub * RADIX + tid];
      before += cnt[sub * RADIX + tid];
    }
  }
  OS_T(7)  // look-back (digit 0)
  __syncthreads();
  OS_T(8)  // barrier after look-back
#pragma unroll 1
  for (int sub = 0; sub < S; sub++) {
    const u32 base = base0 + (u32)sub * RS_SUB;
    const int nvalid = base < n ? (int)min((u32)RS_SUB, n - base) : 0;
    if (ent_out && sub == 0 && warp == 0 && nvalid == RS_SUB)  // what the next pass will see in one warp
      sample_diversity((u32)((stage[tid] >> next_shift) & next_mask), lane, ent_out);
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
      const int p = j * RS_THREADS + tid;
      if (p < nvalid) {
        const u64 k = stage[sub * RS_SUB + p];
        const u32 g = digit_base[sub * RADIX + (u32)((k >> shift) & mask)] + (u32)p;
        kout[g] = k;
        if (HAS_VALS) vout[g] = vstage[sub * RS_SUB + p];
      }
    }
  }
  OS_T(9)  // write-out
#ifdef CB_OS_TIMING
  if (tid == 0) atomicAdd(&g_os_prof[14], 1ull);
#endif
}

template <bool HAS_VALS, int S, int RB>
static void launch_onesweep(u32 tiles, cudaStream_t s, const u64* kin, u64* kout, const u32* vin, u32* vout, u32* status,
                            u32* ctl, const u32* gbase, u32 n, int bit, u32 mask, const u32* ent_in, u32* ent_out,
                            int next_shift, u32 next_mask, int force_rank) {
  constexpr size_t dyn = os_smem_bytes<HAS_VALS, S, RB>();
  // the attribute is per device and contexts may live on several devices and threads: set it on every
  // launch (a driver-side table write, no synchronisation)
  CB_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<HAS_VALS, S, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)dyn));
  rs_onesweep_kernel<HAS_VALS, S, RB><<<tiles, RS_THREADS, dyn, s>>>(kin, kout, vin, vout, status, ctl, gbase, ctl + 1, n,
                                                                    bit, mask, ent_in, ent_out, next_shift, next_mask,
                                                                    force_rank);
}

static int radix_sort_onesweep(u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, size_t n, int begin_bit,
                               int end_bit, cudaStream_t s, LaunchCounter& lc, const char* tag) {
  static const int subs = getenv("CB_OS_SUB") ? atoi(getenv("CB_OS_SUB")) : 1;      // sub-tiles per CTA; 2 measured slower (3.64 vs 3.27 ms): 225 KB of shared memory per SM leave no L1 for loads in flight
  static const int force_rank = getenv("CB_RANK") ? atoi(getenv("CB_RANK")) : 0;   // experiment: 1 match, 2 ballots
  static const int digit_bits = getenv("CB_DIGIT_BITS") ? atoi(getenv("CB_DIGIT_BITS")) : 9;  // experiment: 8
  const int S = subs == 2 ? 2 : 1;
  const PassPlan pl = make_plan(begin_bit, end_bit, digit_bits == 8 ? 8 : 9);
  const int npass = pl.npass;
  const u32 tiles = (u32)((n + (size_t)S * RS_SUB - 1) / ((size_t)S * RS_SUB));
  DevBuf<u32> ghist, status, ctl;
  ghist.alloc((size_t)npass * RS_MAXRADIX, s);
  status.alloc((size_t)tiles * RS_MAXRADIX, s);
  ctl.alloc(2 + 2 * (RS_MAXPASS + 1), s);  // [0] tile counter, [1] error flag, then (distinct, samples) per pass
  u32* ent = ctl.p + 2;
  CB_CUDA(cudaMemsetAsync(ghist.p, 0, (size_t)npass * RS_MAXRADIX * 4, s));
  CB_CUDA(cudaMemsetAsync(ctl.p, 0, (2 + 2 * (RS_MAXPASS + 1)) * 4, s));
  const std::string nm_up_s = std::string("rs_upsweep:") + tag, nm_sc_s = std::string("rs_scatter:") + tag;
  lc.begin(nm_up_s.c_str());
  {
    int blocks = div_up(n, 256 * 16);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    rs_hist_all_kernel<<<blocks, 256, (size_t)npass * RS_MAXRADIX * 4, s>>>(keys_a, n, pl, ghist.p, ent);
    rs_scan_hist_kernel<<<1, RS_MAXRADIX, 0, s>>>(ghist.p, npass);
  }
  lc.end(nm_up_s.c_str());
  lc.n += 2;
  u64* kin = keys_a;
  u64* kout = keys_b;
  u32* vin = vals_a;
  u32* vout = vals_b;
  int where = 0;
  for (int p = 0; p < npass; p++) {
    const int bit = pl.shift[p];
    const u32 mask = (1u << pl.bits[p]) - 1;
    const bool last = p + 1 == npass;
    const int nbit = last ? 0 : pl.shift[p + 1];
    const u32 nmask = last ? 0u : (1u << pl.bits[p + 1]) - 1;
    u32* ent_out = last ? nullptr : ent + 2 * (p + 1);
    const int radix = pl.bits[p] > 8 ? 512 : 256;
    CB_CUDA(cudaMemsetAsync(status.p, 0, (size_t)tiles * radix * 4, s));
    CB_CUDA(cudaMemsetAsync(ctl.p, 0, 4, s));
    const u32* gb = ghist.p + p * RS_MAXRADIX;
    lc.begin(nm_sc_s.c_str());
#define CB_OS_GO(HV, SS, RB) \
  launch_onesweep<HV, SS, RB>(tiles, s, kin, kout, vin, vout, status.p, ctl.p, gb, (u32)n, bit, mask, ent + 2 * p, ent_out, nbit, nmask, force_rank)
#define CB_OS_GO_RB(HV, SS) { if (radix == 512) CB_OS_GO(HV, SS, 9); else CB_OS_GO(HV, SS, 8); }
    if (S == 2) { if (vin) CB_OS_GO_RB(true, 2) else CB_OS_GO_RB(false, 2) }
    else { if (vin) CB_OS_GO_RB(true, 1) else CB_OS_GO_RB(false, 1) }
#undef CB_OS_GO_RB
#undef CB_OS_GO
    lc.end(nm_sc_s.c_str());
    lc.n++;
    CB_CUDA(cudaGetLastError());
    u64* tk = kin; kin = kout; kout = tk;
    u32* tv = vin; vin = vout; vout = tv;
    where ^= 1;
  }
  u32 err = 0;
  lc.rb.get(&err, ctl.p + 1, 4, s);
  lc.rb.sync(s);
  if (err) throw Error(CB_E_CUDA, "radix sort: look-back spin limit exceeded");
  return where;
}

static bool use_onesweep() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CB_SORT");
    v = (e && std::string(e) == "classic") ? 0 : 1;
  }
  return v == 1;
}

int radix_sort(u64* keys_a, u64* keys_b, u32* vals_a, u32* vals_b, size_t n, int begin_bit, int end_bit,
               cudaStream_t s, LaunchCounter& lc, const char* tag) {
  if (n == 0 || end_bit <= begin_bit) return 0;
  if (n >= (1ull << 32)) throw Error(CB_E_CAPACITY, "radix_sort: more than 2^32-1 elements");
  if (n < (1ull << 30) && end_bit - begin_bit <= 64 && use_onesweep())
    return radix_sort_onesweep(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, s, lc, tag);
  int tiles = (int)((n + RS_TILE - 1) / RS_TILE);
  int chunks = tiles < 4 * kNumSMs ? tiles : 4 * kNumSMs;
  int rows_per_chunk = (tiles + chunks - 1) / chunks;
  chunks = (tiles + rows_per_chunk - 1) / rows_per_chunk;
  DevBuf<u32> hist, chunk;
  hist.alloc((size_t)tiles * 256, s);
  chunk.alloc((size_t)chunks * 256, s);
  const std::string nm_up_s = std::string("rs_upsweep:") + tag, nm_sc_s = std::string("rs_scatter:") + tag;
  const char* nm_up = nm_up_s.c_str();
  const char* nm_sc = nm_sc_s.c_str();
  u64* kin = keys_a;
  u64* kout = keys_b;
  u32* vin = vals_a;
  u32* vout = vals_b;
  int where = 0;
  for (int bit = begin_bit; bit < end_bit; bit += 8) {
    int nbits = end_bit - bit < 8 ? end_bit - bit : 8;
    u32 mask = (1u << nbits) - 1;
    lc.begin(nm_up);
    if (nbits < 6) rs_upsweep<true><<<tiles, RS_UP_THREADS, 0, s>>>(kin, hist.p, n, bit, mask);
    else rs_upsweep<false><<<tiles, RS_UP_THREADS, 0, s>>>(kin, hist.p, n, bit, mask);
    lc.end(nm_up);
    rs_chunk_sums<<<chunks, 256, 0, s>>>(hist.p, chunk.p, tiles, rows_per_chunk);
    rs_chunk_scan<<<1, 256, 0, s>>>(chunk.p, chunks);
    rs_apply<<<chunks, 256, 0, s>>>(hist.p, chunk.p, tiles, rows_per_chunk);
    lc.begin(nm_sc);
    if (vin)
      rs_scatter<true><<<tiles, RS_THREADS, 0, s>>>(kin, kout, vin, vout, hist.p, n, bit, mask);
    else
      rs_scatter<false><<<tiles, RS_THREADS, 0, s>>>(kin, kout, nullptr, nullptr, hist.p, n, bit, mask);
    lc.end(nm_sc);
    lc.n += 5;
    CB_CUDA(cudaGetLastError());
    u64* tk = kin; kin = kout; kout = tk;
    u32* tv = vin; vin = vout; vout = tv;
    where ^= 1;
  }
  return where;
}


// ------------------------------------------------------------------------------------------
// hash partition (see primitives.cuh): level-1 / level-2 digit histograms, the segment and tile
// tables, and the atomic-reserve scatter
// ------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 512;
constexpr int PS_ITEMS = 8;
constexpr int PS_TILE = PS_THREADS * PS_ITEMS;  // 4096 tuples per tile

BucketPlan make_bucket_plan_bits(int nbits) {
  // digit widths the scatter kernel is instantiated for; two passes, the smallest total that reaches the
  // wanted number of buckets (more buckets than wanted only makes them smaller)
  static const int widths[5] = {3, 6, 8, 9, 10};
  BucketPlan pl;
  int best = 99;
  for (int a : widths)
    for (int b : widths)
      if (a >= b && a + b >= nbits && a + b < best) { best = a + b; pl.b1 = a; pl.b2 = b; }
  pl.NB = 1u << (pl.b1 + pl.b2);
  pl.shift1 = 64 - pl.b1;
  pl.shift2 = 64 - pl.b1 - pl.b2;
  pl.lshift = 64 - pl.b1 - pl.b2 - kLocalBits;
  return pl;
}

BucketPlan make_bucket_plan(size_t n_tuples) {
  static const size_t avg = (getenv("CB_BUCKET_AVG") && atoi(getenv("CB_BUCKET_AVG")) > 0) ? (size_t)atoi(getenv("CB_BUCKET_AVG")) : 640;  // experiment knob
  int nbits = 2;
  while (nbits < 20 && (avg << nbits) < n_tuples) nbits++;
  BucketPlan pl = make_bucket_plan_bits(nbits);
  pl.fits = nbits < 20 || ((size_t)640 << nbits) >= n_tuples;  // beyond 2^20 buckets of <= 640: a third pass would be needed
  return pl;
}

// where the tiles of a pass come from: a flat array of n slots (tile_seg == nullptr), or the tiles
// of the level-1 segments (a tile never straddles two segments)
struct TileSrc {
  const u32* tile_seg;   // segment of tile t, ~0 beyond the last tile
  const u32* tile_off;   // index of tile t within its segment
  const u32* seg_start;  // [S + 1]
  u32 n;
};
__device__ __forceinline__ bool tile_range(const TileSrc& ts, u32 tile, u32& seg, u32& base, u32& count) {
  if (!ts.tile_seg) {
    seg = 0;
    if ((size_t)tile * PS_TILE >= (size_t)ts.n) return false;
    base = tile * (u32)PS_TILE;
    count = min((u32)PS_TILE, ts.n - base);
    return true;
  }
  seg = ts.tile_seg[tile];
  if (seg == 0xffffffffu) return false;
  const u32 s0 = ts.seg_start[seg], s1 = ts.seg_start[seg + 1];
  base = s0 + ts.tile_off[tile] * (u32)PS_TILE;
  count = min((u32)PS_TILE, s1 - base);
  return true;
}

// digit counts of one pass: per tile in shared memory, then one global add per non-empty bin
__global__ void __launch_bounds__(256) ph_hist_kernel(const u64* __restrict__ kin, TileSrc ts, u64 kmask, int shift,
                                                      int bits, u32* __restrict__ ghist) {
  extern __shared__ u32 ph_h[];
  const u32 bins = 1u << bits, mask = bins - 1;
  u32 seg, base, count;
  if (!tile_range(ts, blockIdx.x, seg, base, count)) return;  // uniform
  for (u32 t = threadIdx.x; t < bins; t += blockDim.x) ph_h[t] = 0;
  __syncthreads();
  for (u32 e = threadIdx.x; e < count; e += blockDim.x) {
    const u64 k = kin[base + e];
    if (k != ~0ull) atomicAdd(&ph_h[(u32)(kmer_hash(k & kmask) >> shift) & mask], 1u);
  }
  __syncthreads();
  for (u32 t = threadIdx.x; t < bins; t += blockDim.x) {
    const u32 v = ph_h[t];
    if (v) atomicAdd(&ghist[((size_t)seg << bits) + t], v);
  }
}

// one block: level-1 digit counts -> segment offsets, the scatter cursors of pass 1, and the tile table
// of the passes that read the segments (S <= 2048 segments, two per thread)
__global__ void __launch_bounds__(1024) ph_segments_kernel(const u32* __restrict__ cnt, int S, u32* __restrict__ seg_start,
                                                           u32* __restrict__ cursor, u32* __restrict__ tile_seg,
                                                           u32* __restrict__ tile_off, u32 max_tiles,
                                                           u32* __restrict__ total_out) {
  __shared__ u32 tmp[1024 / 32 + 1];
  const int s0 = 2 * (int)threadIdx.x, s1 = s0 + 1;
  const u32 c0 = s0 < S ? cnt[s0] : 0u, c1 = s1 < S ? cnt[s1] : 0u;
  const u32 t0 = (c0 + PS_TILE - 1) / PS_TILE, t1 = (c1 + PS_TILE - 1) / PS_TILE;
  u32 total, ttotal;
  const u32 ex = block_excl_scan(c0 + c1, tmp, total);
  const u32 et = block_excl_scan(t0 + t1, tmp, ttotal);
  if (s0 < S) { seg_start[s0] = ex; cursor[s0] = ex; }
  if (s1 < S) { seg_start[s1] = ex + c0; cursor[s1] = ex + c0; }
  if (threadIdx.x == 0) { seg_start[S] = total; *total_out = total; }
  for (u32 i = 0; i < t0; i++)
    if (et + i < max_tiles) { tile_seg[et + i] = (u32)s0; tile_off[et + i] = i; }
  for (u32 i = 0; i < t1; i++)
    if (et + t0 + i < max_tiles) { tile_seg[et + t0 + i] = (u32)s1; tile_off[et + t0 + i] = i; }
}

// One tile: rank the tuples by digit inside every warp (ballots over the digit bits, a warp-private
// histogram -- a shared-memory atomic per tuple costs 64 LSU cycles per warp instruction on this part and
// made the pass 70 % slower than the LSD pass it replaces), scan the warp histograms per digit, reserve the
// tile's range of every bin with ONE global atomic per bin, stage the tuples by bin in shared memory and
// write every bin's run as one contiguous piece.  Hashed digits are uniformly random, so the ballots win
// over MATCH.ANY (see digit_peers).  Slots whose key is ~0 (windows that emit nothing) are dropped.
template <int RB>
__global__ void __launch_bounds__(PS_THREADS, 2) ph_scatter_kernel(const u64* __restrict__ kin, const u32* __restrict__ vin,
                                                                   u64* __restrict__ kout, u32* __restrict__ vout,
                                                                   TileSrc ts, u64 kmask, int shift,
                                                                   u32* __restrict__ cursor) {
  constexpr int RADIX = 1 << RB;
  constexpr int WARPS = PS_THREADS / 32;
  constexpr int BPT = RADIX > PS_THREADS ? RADIX / PS_THREADS : 1;  // adjacent digit bins per thread (RB = 10: two)
  static_assert(RADIX <= 2 * PS_THREADS, "at most two digit bins per thread");
  extern __shared__ __align__(16) unsigned char ps_smem[];
  u64* stage = reinterpret_cast<u64*>(ps_smem);                 // [PS_TILE]
  u32* vstage = reinterpret_cast<u32*>(ps_smem + PS_TILE * 8);  // [PS_TILE]
  unsigned short(*warp_hist)[RADIX] = reinterpret_cast<unsigned short(*)[RADIX]>(vstage + PS_TILE);  // [WARPS][RADIX]
  u32* excl = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(warp_hist) + WARPS * RADIX * 2);  // [RADIX]
  u32* gbase = excl + RADIX;                                                                            // [RADIX]
  unsigned short* sdig = reinterpret_cast<unsigned short*>(gbase + RADIX);                              // [PS_TILE] digit of the staged tuple
  __shared__ u32 scan_tmp[WARPS + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  u32 seg, base, count;
  if (!tile_range(ts, blockIdx.x, seg, base, count)) return;  // uniform
  for (int t = tid; t < WARPS * RADIX / 2; t += PS_THREADS) reinterpret_cast<u32*>(&warp_hist[0][0])[t] = 0;
  u64 key[PS_ITEMS];
  u32 val[PS_ITEMS], pos[PS_ITEMS];
  const u32 e0 = (u32)warp * (PS_ITEMS * 32) + lane;  // warp-striped
#pragma unroll
  for (int i = 0; i < PS_ITEMS; i++) {
    const u32 e = e0 + i * 32;
    key[i] = e < count ? kin[base + e] : ~0ull;
  }
#pragma unroll
  for (int i = 0; i < PS_ITEMS; i++) {  // requested with the keys: one DRAM round trip per tile instead of two
    const u32 e = e0 + i * 32;
    val[i] = e < count ? vin[base + e] : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < PS_ITEMS; i++) {
    const bool valid = key[i] != ~0ull;
    const u32 d = (u32)(kmer_hash(key[i] & kmask) >> shift) & (RADIX - 1);
    u32 peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < RB; b++) {
      const bool bit = (d >> b) & 1u;
      const u32 bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
    const u32 prev = warp_hist[warp][d];
    __syncwarp();
    if (valid && lane == __ffs(peers) - 1) warp_hist[warp][d] = (unsigned short)(prev + __popc(peers));
    __syncwarp();
    pos[i] = valid ? ((prev + __popc(peers & lt_mask)) | (d << 12)) : 0xffffffffu;  // the hash is taken once per tuple
  }
  __syncthreads();
  u32 run = 0, cnt_b[BPT];
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const int b = tid * BPT + q;
    cnt_b[q] = 0;
    if (b < RADIX) {
      u32 r = 0;
#pragma unroll
      for (int w = 0; w < WARPS; w++) {
        const u32 c = warp_hist[w][b];
        warp_hist[w][b] = (unsigned short)r;
        r += c;
      }
      cnt_b[q] = r;
      run += r;
    }
  }
  u32 total;
  u32 ex = block_excl_scan(run, scan_tmp, total);  // threads beyond the bins contribute 0
#pragma unroll
  for (int q = 0; q < BPT; q++) {
    const int b = tid * BPT + q;
    if (b < RADIX) {
      excl[b] = ex;
      if (cnt_b[q]) gbase[b] = atomicAdd(&cursor[((size_t)seg << RB) + b], cnt_b[q]) - ex;
      ex += cnt_b[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < PS_ITEMS; i++) {
    if (pos[i] != 0xffffffffu) {
      const u32 d = pos[i] >> 12;
      pos[i] = (pos[i] & 4095u) + excl[d] + warp_hist[warp][d];
      stage[pos[i]] = key[i];
      sdig[pos[i]] = (unsigned short)d;
    }
  }
#pragma unroll
  for (int i = 0; i < PS_ITEMS; i++)
    if (pos[i] != 0xffffffffu) vstage[pos[i]] = val[i];
  __syncthreads();
  for (u32 p = tid; p < total; p += PS_THREADS) {
    const u32 g = gbase[sdig[p]] + p;
    kout[g] = stage[p];
    vout[g] = vstage[p];
  }
}

template <int RB>
static void launch_scatter_rb(u32 tiles, cudaStream_t s, const u64* kin, const u32* vin, u64* kout, u32* vout,
                              const TileSrc& ts, u64 kmask, int shift, u32* cursor) {
  constexpr size_t dyn = (size_t)PS_TILE * 14 + (PS_THREADS / 32) * ((size_t)2 << RB) + 2 * ((size_t)4 << RB);
  CB_CUDA(cudaFuncSetAttribute(ph_scatter_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  ph_scatter_kernel<RB><<<tiles, PS_THREADS, dyn, s>>>(kin, vin, kout, vout, ts, kmask, shift, cursor);
}
static void launch_scatter(u32 tiles, cudaStream_t s, const u64* kin, const u32* vin, u64* kout, u32* vout, const TileSrc& ts,
                           u64 kmask, int shift, int bits, u32* cursor) {
  switch (bits) {
    case 3: launch_scatter_rb<3>(tiles, s, kin, vin, kout, vout, ts, kmask, shift, cursor); break;
    case 6: launch_scatter_rb<6>(tiles, s, kin, vin, kout, vout, ts, kmask, shift, cursor); break;
    case 8: launch_scatter_rb<8>(tiles, s, kin, vin, kout, vout, ts, kmask, shift, cursor); break;
    case 9: launch_scatter_rb<9>(tiles, s, kin, vin, kout, vout, ts, kmask, shift, cursor); break;
    case 10: launch_scatter_rb<10>(tiles, s, kin, vin, kout, vout, ts, kmask, shift, cursor); break;
    default: throw Error(CB_E_INVALID, "bucket_partition: unsupported digit width");
  }
  CB_CUDA(cudaGetLastError());
}

void bucket_partition(u64* ka, u32* va, u64* kb, u32* vb, size_t n_slots, u64 kmask, const BucketPlan& pl,
                      const u32* d_cnt1, u32* bstart, u32* d_total, cudaStream_t s, LaunchCounter& lc) {
  if (n_slots >= (1ull << 32)) throw Error(CB_E_CAPACITY, "bucket_partition: more than 2^32-1 tuple slots");
  const int S = 1 << pl.b1;
  const u32 tiles1 = (u32)((n_slots + PS_TILE - 1) / PS_TILE);
  const u32 max_tiles2 = tiles1 + (u32)S + 1;
  DevBuf<u32> cnt1, seg_start, cursor1, tile_seg, tile_off, hist2, cursor2;
  seg_start.alloc((size_t)S + 1, s);
  cursor1.alloc(S, s);
  tile_seg.alloc(max_tiles2, s);
  tile_off.alloc(max_tiles2, s);
  hist2.alloc((size_t)pl.NB + 1, s);
  cursor2.alloc(pl.NB, s);
  TileSrc flat{nullptr, nullptr, nullptr, (u32)n_slots};
  TileSrc segs{tile_seg.p, tile_off.p, seg_start.p, 0u};
  if (!d_cnt1) {
    cnt1.alloc(S, s);
    CB_CUDA(cudaMemsetAsync(cnt1.p, 0, (size_t)S * 4, s));
    lc.begin("ph_hist1");
    if (tiles1) ph_hist_kernel<<<tiles1, 256, (size_t)4 << pl.b1, s>>>(ka, flat, kmask, pl.shift1, pl.b1, cnt1.p);
    lc.end("ph_hist1");
    lc.n++;
    d_cnt1 = cnt1.p;
  }
  CB_CUDA(cudaMemsetAsync(tile_seg.p, 0xff, (size_t)max_tiles2 * 4, s));
  CB_CUDA(cudaMemsetAsync(hist2.p, 0, ((size_t)pl.NB + 1) * 4, s));
  ph_segments_kernel<<<1, 1024, 0, s>>>(d_cnt1, S, seg_start.p, cursor1.p, tile_seg.p, tile_off.p, max_tiles2, d_total);
  lc.n++;
  CB_CUDA(cudaGetLastError());
  lc.begin("ph_scatter1");
  if (tiles1) launch_scatter(tiles1, s, ka, va, kb, vb, flat, kmask, pl.shift1, pl.b1, cursor1.p);
  lc.end("ph_scatter1");
  lc.begin("ph_hist2");
  ph_hist_kernel<<<max_tiles2, 256, (size_t)4 << pl.b2, s>>>(kb, segs, kmask, pl.shift2, pl.b2, hist2.p);
  lc.end("ph_hist2");
  lc.n += 2;
  CB_CUDA(cudaGetLastError());
  exclusive_scan_u32(hist2.p, bstart, (size_t)pl.NB + 1, s, lc);
  CB_CUDA(cudaMemcpyAsync(cursor2.p, bstart, (size_t)pl.NB * 4, cudaMemcpyDeviceToDevice, s));
  lc.begin("ph_scatter2");
  launch_scatter(max_tiles2, s, kb, vb, ka, va, segs, kmask, pl.shift2, pl.b2, cursor2.p);
  lc.end("ph_scatter2");
  lc.n++;
}

}  // namespace cb
