// stage_preprocess.cu -- SFA loader / 2-bit read packer and duplicate-read removal.
//
// Replaces (reference):
//   GenNonContainedReadsMapper.map      GenNonContainedReads.java:58-130   (line parse, filters)
//   GenNonContainedReadsReducer.reduce  GenNonContainedReads.java:142-251  (duplicate detection)
//   RedundantRemovalReducer.reduce      RedundantRemoval.java:66-103       (drop "*n" nodes)
//
// Net effect of those three: among reads that are identical up to reverse
// complement the read with the smallest id under String.compareTo survives in its own
// orientation with cov = 1 + #duplicates (a read equal to its own reverse complement counts
// every duplicate twice).
#include "ctx.cuh"

namespace cb {

constexpr int PT_THREADS = 256;
constexpr int PT_BYTES = 16;                       // bytes per thread
constexpr int PT_BLOCK = PT_THREADS * PT_BYTES;    // bytes per block

// Hadoop LineReader: a line ends at '\n', at "\r\n", or at a lone '\r'.
__device__ __forceinline__ bool is_term(const char* __restrict__ text, size_t i, size_t n) {
  char c = text[i];
  if (c == '\n') return true;
  if (c == '\r') return (i + 1 >= n) || text[i + 1] != '\n';
  return false;
}

// 0x80 in every byte of w that equals the byte replicated in `pat` (exact, no carry artefacts)
__device__ __forceinline__ u32 bytes_equal(u32 w, u32 pat) {
  const u32 t = w ^ pat;
  return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);
}

// bit j set: byte j of the thread's 16-byte chunk ends a line.  Four bytes per step (SWAR): a
// '\n' always ends a line; '\r' (rare) does unless a '\n' follows, which is resolved per byte.
__device__ __forceinline__ u32 thread_term_mask(const char* __restrict__ text, size_t base, size_t n) {
  u32 m = 0;
  if (base + PT_BYTES < n) {  // whole chunk and the look-ahead byte are in range
    uint4 v = *reinterpret_cast<const uint4*>(text + base);
    u32 w[4] = {v.x, v.y, v.z, v.w};
    u32 any_cr = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const u32 nl = bytes_equal(w[q], 0x0a0a0a0au);
      any_cr |= bytes_equal(w[q], 0x0d0d0d0du);
      // 0x80 of bytes 0..3 -> bits 0..3
      m |= (((nl >> 7) * 0x00204081u) >> 21 & 0xfu) << (4 * q);
    }
    if (any_cr) {  // a lone '\r' also ends a line (Hadoop LineReader)
      char nxt = text[base + PT_BYTES];
#pragma unroll
      for (int j = 0; j < PT_BYTES; j++) {
        char c = (char)((w[j >> 2] >> ((j & 3) * 8)) & 0xff);
        char c1 = (j + 1 < PT_BYTES) ? (char)((w[(j + 1) >> 2] >> (((j + 1) & 3) * 8)) & 0xff) : nxt;
        if (c == '\r' && c1 != '\n') m |= 1u << j;
      }
    }
  } else {
    for (int j = 0; j < PT_BYTES; j++)
      if (base + j < n && is_term(text, base + j, n)) m |= 1u << j;
  }
  return m;
}

// Both line-terminator kernels handle PT_SUB consecutive 4 KB sub-blocks per CTA with all their 16-byte loads
// issued before the first is consumed (one 16-byte load per thread left the kernels at 1.8 TB/s: too few
// bytes in flight); the per-sub-block counts and offsets keep their layout.
constexpr int PT_SUB = 4;

__global__ void __launch_bounds__(PT_THREADS) count_terms_kernel(const char* __restrict__ text, size_t n,
                                                                 u32* __restrict__ block_counts, u32 n_sub) {
  __shared__ u32 tmp[PT_SUB][PT_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 cnt[PT_SUB];
#pragma unroll
  for (int q = 0; q < PT_SUB; q++) {
    const size_t base = ((size_t)blockIdx.x * PT_SUB + q) * PT_BLOCK + (size_t)threadIdx.x * PT_BYTES;
    cnt[q] = base < n ? __popc(thread_term_mask(text, base, n)) : 0;
  }
#pragma unroll
  for (int q = 0; q < PT_SUB; q++) {
    u32 v = cnt[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) tmp[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < PT_SUB) {
    const u32 sub = blockIdx.x * PT_SUB + threadIdx.x;
    u32 total = 0;
    for (int w = 0; w < PT_THREADS / 32; w++) total += tmp[threadIdx.x][w];
    if (sub < n_sub) block_counts[sub] = total;
  }
}

__global__ void __launch_bounds__(PT_THREADS) fill_line_starts_kernel(const char* __restrict__ text, size_t n,
                                                                      const u32* __restrict__ block_offs,
                                                                      u32* __restrict__ line_start, u32 n_sub) {
  __shared__ u32 tmp[PT_SUB][PT_THREADS / 32 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 m[PT_SUB], inc[PT_SUB];
#pragma unroll
  for (int q = 0; q < PT_SUB; q++) {
    const size_t base = ((size_t)blockIdx.x * PT_SUB + q) * PT_BLOCK + (size_t)threadIdx.x * PT_BYTES;
    m[q] = base < n ? thread_term_mask(text, base, n) : 0;
  }
#pragma unroll
  for (int q = 0; q < PT_SUB; q++) {  // inclusive scan of the counts inside the warp
    u32 v = __popc(m[q]);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    inc[q] = v;
    if (lane == 31) tmp[q][warp] = v;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < PT_SUB; q++) {
    const u32 sub = blockIdx.x * PT_SUB + q;
    if (sub >= n_sub) break;
    const size_t base = (size_t)sub * PT_BLOCK + (size_t)threadIdx.x * PT_BYTES;
    u32 woff = 0;
    for (int w = 0; w < warp; w++) woff += tmp[q][w];
    u32 rank = block_offs[sub] + woff + inc[q] - __popc(m[q]);
    u32 mm = m[q];
    while (mm) {
      const int j = __ffs(mm) - 1;
      mm &= mm - 1;
      line_start[++rank] = (u32)(base + j + 1);  // line (rank) starts after terminator #(rank-1)
    }
  }
}

__global__ void set_u32_kernel(u32* p, u32 v) { *p = v; }

// One thread per SFA line, one pass over the line's bytes in aligned 16-byte chunks (vector
// loads; neighbouring threads share chunks through L1): Java split("\t") semantics, uppercase,
// ACGT filter, length filter, 2-bit packing (row written with 16-byte vector stores), 16-byte id
// key, and the 64-bit hash of the canonical read min(s, rc(s)) that keys the duplicate detection.
//   code(c) = ((c >> 1) ^ (c >> 2)) & 3 maps A,C,G,T (either case) to 0,1,2,3
//   valid(c): (c & 0xDF) - 'A' in {0, 2, 6, 19}
template <int NW>
__global__ void __launch_bounds__(128) parse_lines_kernel(const char* __restrict__ text, size_t text_n,
                                                          const u32* __restrict__ line_start, u32 n_lines, int K,
                                                          int L, uint8_t* __restrict__ status,
                                                          u64* __restrict__ reads, IdKey* __restrict__ idkey,
                                                          u64* __restrict__ hkeys, u32* __restrict__ hvals,
                                                          ull* __restrict__ dcount) {
  // grid-stride over the lines; counters are accumulated per thread and flushed once per warp at
  // the end (a same-address atomic per read would serialise the kernel at the L2 atomic unit)
  u32 c_inv = 0, c_skip = 0, c_short = 0, c_good = 0, c_mis = 0, c_long = 0;
  ull c_bp = 0;
  for (u32 line = blockIdx.x * blockDim.x + threadIdx.x; line < n_lines; line += gridDim.x * blockDim.x) {
  const u32 s = line_start[line];
  u32 e = line_start[line + 1];
  int field = 0, last_nonempty = -1;
  u32 idlen = 0;
  int nb = 0;  // bytes of field 1
  bool bad = false, done = false;
  IdKey ik;
  ik.hi = 0;
  ik.lo = 0;
  u64 w[NW];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
  u64 acc = 0;
  for (u32 chunk = s & ~15u; chunk < e && !done; chunk += 16) {
    u32 v[4];
    if ((size_t)chunk + 16 <= text_n) {
      uint4 q = *reinterpret_cast<const uint4*>(text + chunk);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      v[0] = v[1] = v[2] = v[3] = 0;
      for (int j = 0; j < 16; j++)
        if ((size_t)chunk + j < text_n) v[j >> 2] |= (u32)(unsigned char)text[chunk + j] << ((j & 3) * 8);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const u32 p0 = chunk + 4 * q;
      const u32 wq = v[q];
      // fast path: four sequence bytes at once (SWAR) -- all inside field 1 and the line, all of
      // them ACGT in either case (which also rules out a tab or a terminator), and not straddling
      // a 32-base word of the packed row
      const u32 cc = ((wq >> 1) ^ (wq >> 2)) & 0x03030303u;  // A,C,G,T -> 0,1,2,3 per byte
      const u32 hi2 = (cc >> 1) & 0x01010101u, both = cc & hi2;
      const u32 expect = 0x41414141u + 2u * cc + 2u * hi2 + 11u * both;  // 'A','C','G','T' of the codes
      if (field == 1 && !done && p0 >= s && p0 + 3 < e && expect == (wq & 0xdfdfdfdfu) && (nb & 31) <= 28) {
        acc = (acc << 8) | (u64)(((cc * 0x40100401u) >> 24) & 0xffu);  // first byte -> highest pair
        nb += 4;
        last_nonempty = 1;
        if ((nb & 31) == 0) {
          int wi = (nb >> 5) - 1;
#pragma unroll
          for (int t = 0; t < NW; t++)
            if (t == wi) w[t] = acc;
        }
        continue;
      }
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const u32 p = p0 + jj;
        const u32 c = (wq >> (jj * 8)) & 0xffu;
        if (p < s || p >= e || done) continue;
        if (c == '\n' || c == '\r') { done = true; continue; }  // the terminator is the end of the line
        if (c == '\t') { field++; continue; }
        last_nonempty = field;
        if (field == 0) {
          if (idlen < 8) ik.hi |= (u64)c << (56 - 8 * idlen);
          else if (idlen < 16) ik.lo |= (u64)c << (56 - 8 * (idlen - 8));
          idlen++;
        } else if (field == 1) {
          u32 u = (c & 0xdfu) - 'A';
          bool ok = u < 20u && ((0x80045u >> u) & 1u);
          bad |= !ok;
          acc = (acc << 2) | (u64)(((c >> 1) ^ (c >> 2)) & 3u);
          nb++;
          if ((nb & 31) == 0) {
            int wi = (nb >> 5) - 1;
#pragma unroll
            for (int t = 0; t < NW; t++)
              if (t == wi) w[t] = acc;
          }
        }
      }
    }
  }
  if (nb & 31) {  // left-align the last partial word
    int wi = nb >> 5;
    u64 last = acc << (64 - 2 * (nb & 31));
#pragma unroll
    for (int q = 0; q < NW; q++)
      if (q == wi) w[q] = last;
  }
  const int nf = last_nonempty + 1;
  uint8_t st = 0;
  u64 hk = ~0ull;
  if (nf != 2 && nf != 3) { st = 1; c_inv++; }  // GenNonContainedReads.java:64-69
  else if (bad) { st = 2; c_skip++; }          // :102-107
  else if (nb <= K) { st = 3; c_short++; }     // :110-115
  else {
    c_good++;
    c_bp += (ull)nb;
    if (nb != L) c_mis++;
    if (idlen > 16) c_long++;
  }
  if (st == 0) {
    idkey[line] = ik;
    u64 rc[NW];
    revcomp_read<NW>(w, rc, L);
    bool fwd = cmp_words<NW>(w, rc) <= 0;
    u64 h = 0x9e3779b97f4a7c15ull;
#pragma unroll
    for (int i = 0; i < NW; i++) h = mix64(h ^ (fwd ? w[i] : rc[i])) + 0x9e3779b97f4a7c15ull;
    hk = h & ~1ull;  // bit 0 clear: never collides with the all-ones key of non-reads
  }
  status[line] = st;
  // packed row: NW is even, so the row is NW/2 aligned 16-byte stores
  uint4* row = reinterpret_cast<uint4*>(reads + (size_t)line * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    uint4 v;
    v.x = (u32)w[2 * i]; v.y = (u32)(w[2 * i] >> 32);
    v.z = (u32)w[2 * i + 1]; v.w = (u32)(w[2 * i + 1] >> 32);
    row[i] = v;
  }
  hkeys[line] = hk;
  hvals[line] = line;
  }  // line loop
  const u32 full = 0xffffffffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c_inv += __shfl_xor_sync(full, c_inv, o);
    c_skip += __shfl_xor_sync(full, c_skip, o);
    c_short += __shfl_xor_sync(full, c_short, o);
    c_good += __shfl_xor_sync(full, c_good, o);
    c_mis += __shfl_xor_sync(full, c_mis, o);
    c_long += __shfl_xor_sync(full, c_long, o);
    c_bp += __shfl_xor_sync(full, c_bp, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c_inv) atomicAdd(&dcount[DC_LINES_INVALID], (ull)c_inv);
    if (c_skip) atomicAdd(&dcount[DC_READS_SKIPPED], (ull)c_skip);
    if (c_short) atomicAdd(&dcount[DC_READS_SHORT], (ull)c_short);
    if (c_good) atomicAdd(&dcount[DC_READS_GOOD], (ull)c_good);
    if (c_bp) atomicAdd(&dcount[DC_READS_GOODBP], c_bp);
    if (c_mis) atomicAdd(&dcount[DC_LEN_MISMATCH], (ull)c_mis);
    if (c_long) atomicAdd(&dcount[DC_LONG_ID], (ull)c_long);
  }
}

template <int NW>
__device__ __forceinline__ void load_canon(const u64* __restrict__ reads, u32 idx, int L, u64 (&canon)[NW],
                                           bool& selfrc) {
  u64 r[NW], rc[NW];
  load_read_row<NW>(reads, idx, r);
  revcomp_read<NW>(r, rc, L);
  int c = cmp_words<NW>(r, rc);
  selfrc = (c == 0);
#pragma unroll
  for (int i = 0; i < NW; i++) canon[i] = c <= 0 ? r[i] : rc[i];
}

template <int NW>
__device__ __forceinline__ void canon_of(const uint4 (&q)[NW / 2], int L, u64 (&canon)[NW], bool& selfrc) {
  u64 r[NW], rc[NW];
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    r[2 * i] = (u64)q[i].x | ((u64)q[i].y << 32);
    r[2 * i + 1] = (u64)q[i].z | ((u64)q[i].w << 32);
  }
  revcomp_read<NW>(r, rc, L);
  int c = cmp_words<NW>(r, rc);
  selfrc = (c == 0);
#pragma unroll
  for (int i = 0; i < NW; i++) canon[i] = c <= 0 ? r[i] : rc[i];
}

// Sorted by canonical-read hash: every read scans the run of equal (masked) hashes it lies in,
// compares the canonical reads exactly wherever the full 64-bit hashes agree, and decides on its
// own whether it is the survivor of its duplicate class (smallest id key) and what the class
// size is.  Runs are as long as the duplicate classes (a few reads), so the scans are short and
// every thread is busy.  The two nearest entries on either side are fetched up front together
// with the read's own row (the serial scan was a chain of dependent gathers: entry -> line ->
// row -> id key, once per neighbour); only longer runs continue one entry at a time.
template <int NW>
__global__ void __launch_bounds__(128) dedupe_runs_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                          u32 n_good, u64 keymask, const u64* __restrict__ reads,
                                                          const IdKey* __restrict__ idkey, int L,
                                                          u32* __restrict__ surv_flag, u32* __restrict__ cov,
                                                          ull* __restrict__ dcount) {
  u32 ns = 0, mx = 0;
  const int lane = threadIdx.x & 31;
  // block-uniform trip count: the neighbours' rows and id keys come over warp shuffles (every entry gathers
  // ITS OWN row and id key once; the entries next to it in the sorted order are the lanes next to it, so the
  // four neighbours need no gathers of their own: 250 -> ~100 bytes of DRAM traffic per read)
  for (u32 i0 = blockIdx.x * blockDim.x; i0 < n_good; i0 += gridDim.x * blockDim.x) {
    const u32 i = i0 + threadIdx.x;
    const bool valid = i < n_good;
    const u64 ki = valid ? keys[i] : 0ull, km = ki & keymask;
    const u32 li = valid ? vals[i] : 0u;
    // offsets -2, -1, +1, +2 (index 0..3)
    u64 kn[4];
    u32 ln[4];
    bool hit[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long long j = (long long)i + (q < 2 ? q - 2 : q - 1);
      const bool in = valid && j >= 0 && j < (long long)n_good;
      kn[q] = in ? keys[j] : ~ki;
      ln[q] = in ? vals[j] : 0u;
    }
    // an entry takes part iff the masked hashes agree all the way from i to it
    const bool run_l1 = valid && (kn[1] & keymask) == km, run_l2 = run_l1 && (kn[0] & keymask) == km;
    const bool run_r1 = valid && (kn[2] & keymask) == km, run_r2 = run_r1 && (kn[3] & keymask) == km;
    hit[0] = run_l2 && kn[0] == ki;
    hit[1] = run_l1 && kn[1] == ki;
    hit[2] = run_r1 && kn[2] == ki;
    hit[3] = run_r2 && kn[3] == ki;
    uint4 qi[NW / 2], qn[4][NW / 2];
    IdKey idn[4];
    IdKey mine;
    mine.hi = mine.lo = 0;
#pragma unroll
    for (int w = 0; w < NW / 2; w++) qi[w] = make_uint4(0, 0, 0, 0);
    if (valid) {
      const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)li * NW);
#pragma unroll
      for (int w = 0; w < NW / 2; w++) qi[w] = __ldg(row + w);
      mine = idkey[li];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int src = lane + (q < 2 ? q - 2 : q - 1);
      const bool in_warp = src >= 0 && src < 32;
      const int from = in_warp ? src : lane;
#pragma unroll
      for (int w = 0; w < NW / 2; w++) {
        qn[q][w].x = __shfl_sync(0xffffffffu, qi[w].x, from);
        qn[q][w].y = __shfl_sync(0xffffffffu, qi[w].y, from);
        qn[q][w].z = __shfl_sync(0xffffffffu, qi[w].z, from);
        qn[q][w].w = __shfl_sync(0xffffffffu, qi[w].w, from);
      }
      idn[q].hi = __shfl_sync(0xffffffffu, mine.hi, from);
      idn[q].lo = __shfl_sync(0xffffffffu, mine.lo, from);
      if (hit[q] && !in_warp) {  // the neighbour sits in another warp: gather it
        const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)ln[q] * NW);
#pragma unroll
        for (int w = 0; w < NW / 2; w++) qn[q][w] = __ldg(row + w);
        idn[q] = idkey[ln[q]];
      }
    }
    if (!valid) continue;
    u64 ci[NW], cj[NW];
    bool self, self2;
    canon_of<NW>(qi, L, ci, self);
    u32 cnt = 1;
    bool survivor = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (hit[q]) {
        canon_of<NW>(qn[q], L, cj, self2);
        if (cmp_words<NW>(ci, cj) == 0) {
          cnt++;
          if (idkey_less(idn[q], mine) || (!idkey_less(mine, idn[q]) && ln[q] < li)) survivor = false;
        }
      }
    }
    // longer runs: continue beyond the prefetched entries
    for (int dirn = -1; dirn <= 1; dirn += 2) {
      if (!(dirn < 0 ? run_l2 : run_r2)) continue;
      for (long long j = (long long)i + 3 * dirn; j >= 0 && j < (long long)n_good; j += dirn) {
        const u64 kj = keys[j];
        if ((kj & keymask) != km) break;
        if (kj != ki) continue;
        const u32 lj = vals[j];
        load_canon<NW>(reads, lj, L, cj, self2);
        if (cmp_words<NW>(ci, cj) != 0) continue;
        cnt++;
        const IdKey other = idkey[lj];
        if (idkey_less(other, mine) || (!idkey_less(mine, other) && lj < li)) survivor = false;
      }
    }
    if (survivor) {
      const u32 cv = 1 + (cnt - 1) * (self ? 2u : 1u);  // GenNonContainedReads.java:176-201
      surv_flag[li] = 1;
      cov[li] = cv;
      ns++;
      mx = max(mx, cv);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ns += __shfl_xor_sync(0xffffffffu, ns, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0 && ns) {
    atomicAdd(&dcount[DC_SURVIVORS], (ull)ns);
    atomicMax(&dcount[DC_MAX_COV], (ull)mx);
  }
}

template <int NW>
__global__ void __launch_bounds__(256) build_nodes_kernel(const u32* __restrict__ surv_flag,
                                                          const u32* __restrict__ surv_idx, u32 n_lines,
                                                          const u64* __restrict__ reads,
                                                          const IdKey* __restrict__ idkey, const u32* __restrict__ cov,
                                                          u32 line_base, u32* __restrict__ node_line,
                                                          u64* __restrict__ node_seq,
                                                          IdKey* __restrict__ node_idkey, u32* __restrict__ node_cov) {
  u32 line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines || !surv_flag[line]) return;
  u32 idx = surv_idx[line];
  node_line[idx] = line_base + line;
  node_idkey[idx] = idkey[line];
  node_cov[idx] = cov[line];
  const uint4* src = reinterpret_cast<const uint4*>(reads + (size_t)line * NW);
  uint4* dst = reinterpret_cast<uint4*>(node_seq + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) dst[i] = src[i];
}

// ---- multi-GPU duplicate detection: reads travel to the rank that owns their canonical hash ----
// record = [hash][global line][id key hi][id key lo][read words ...] as u64
template <int NW>
__global__ void __launch_bounds__(256) dd_pack_kernel(const u64* __restrict__ hkeys, const u64* __restrict__ reads,
                                                      const IdKey* __restrict__ idkey, u32 n_lines, u32 line_base,
                                                      int world, u64* __restrict__ recs, u32* __restrict__ dest) {
  u32 line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  const u64 h = hkeys[line];
  u64* r = recs + (size_t)line * (4 + NW);
  r[0] = h;
  r[1] = (u64)(line_base + line);
  if (h == ~0ull) {  // not a good read
    dest[line] = (u32)world;
    return;
  }
  IdKey ik = idkey[line];
  r[2] = ik.hi;
  r[3] = ik.lo;
#pragma unroll
  for (int w = 0; w < NW; w++) r[4 + w] = reads[(size_t)line * NW + w];
  dest[line] = (u32)((h >> 32) & (u64)(world - 1));
}

template <int NW>
__global__ void __launch_bounds__(256) dd_unpack_kernel(const u64* __restrict__ recs, size_t n,
                                                        u64* __restrict__ hkeys, u32* __restrict__ hvals,
                                                        u32* __restrict__ gline, IdKey* __restrict__ idkey,
                                                        u64* __restrict__ reads) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64* r = recs + i * (4 + NW);
  hkeys[i] = r[0];
  hvals[i] = (u32)i;
  gline[i] = (u32)r[1];
  IdKey ik;
  ik.hi = r[2];
  ik.lo = r[3];
  idkey[i] = ik;
#pragma unroll
  for (int w = 0; w < NW; w++) reads[i * NW + w] = r[4 + w];
}

// survivors go back to the rank that holds their SFA line: record = (global line, coverage)
__global__ void __launch_bounds__(256) dd_return_kernel(const u32* __restrict__ surv, const u32* __restrict__ cov,
                                                        const u32* __restrict__ gline, size_t n,
                                                        const u64* __restrict__ line_bases, int world,
                                                        u32* __restrict__ recs, u32* __restrict__ dest) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32 g = gline[i];
  recs[2 * i] = g;
  recs[2 * i + 1] = cov[i];
  u32 d = (u32)world;
  if (surv[i]) {
    d = 0;
    while ((int)d + 1 < world && (u64)g >= line_bases[d + 1]) d++;
  }
  dest[i] = d;
}

__global__ void __launch_bounds__(256) dd_scatter_kernel(const u32* __restrict__ recs, size_t n, u32 line_base,
                                                         u32* __restrict__ surv_flag, u32* __restrict__ cov) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 line = recs[2 * i] - line_base;
  surv_flag[line] = 1;
  cov[line] = recs[2 * i + 1];
}

void stage_parse_and_pack(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const int world = c->cfg.world;
  c->carried_code = 0;
  c->n_host_coll = 0;
  // (a shard of 4 GiB or more is treated as empty here and fails every rank at the first gather below)
  if (c->text_n >= (1ull << 32)) defer_error(c, CB_E_CAPACITY, "SFA input of 4 GiB or more per context");
  const size_t n = failed_locally(c) ? 0 : c->text_n;
  if (c->gather_deferred) {  // the previous run's node-table gather was never needed: it is not started at all
    c->gather_deferred = false;
    c->gather_pending_seq = c->gather_pending_rest = false;
  } else {
    wait_node_table(c, true);  // a gather of the previous run still in flight: its buffers are reused below
  }
  if (n == 0 && !distributed(c)) throw Error(CB_E_NOREADS, "No good reads");
  c->K = c->cfg.k;
  c->L = c->cfg.readlen;
  if (c->K < 1 || c->K > 31 || (c->K % 2) == 0)
    throw Error(CB_E_UNSUPPORTED, "k must be odd and <= 31 (even k lets palindromic k-mers drop nodes; SURVEY 8c)");
  if (c->L <= c->K) throw Error(CB_E_INVALID, "readlen must exceed k");
  if (c->L > 255) throw Error(CB_E_UNSUPPORTED, "readlen above 255");
  if (world < 1 || (world & (world - 1)) || world > 256) throw Error(CB_E_UNSUPPORTED, "world must be a power of two <= 256");
  if (distributed(c) && !c->has_ex && !c->nccl_comm)
    throw Error(CB_E_INVALID, "cfg.world > 1 needs cb_init_nccl or cb_set_exchange before cb_preprocess");
  c->lg_world = 0;
  while ((1 << c->lg_world) < world) c->lg_world++;
  c->pbit = c->K;  // owner of a k-mer group: bits [K, K + lg_world) of the canonical key (its middle bases)
  c->NW = ((c->L + 31) / 32 + 1) & ~1;
  c->W = c->L - c->K + 1;
  c->dcount.alloc(DC_COUNT, s);
  CB_CUDA(cudaMemsetAsync(c->dcount.p, 0, DC_COUNT * sizeof(u64), s));

  timer_start(c, "parse");
  c->n_lines = 0;
  DevBuf<u32> bcnt, boff;
  unsigned nb = (unsigned)div_up(n, PT_BLOCK);
  if (n) {
    bcnt.alloc(nb, s);
    boff.alloc(nb, s);
    count_terms_kernel<<<(nb + PT_SUB - 1) / PT_SUB, PT_THREADS, 0, s>>>(c->text_ptr, n, bcnt.p, nb);
    c->lc.n++;
    exclusive_scan_u32(bcnt.p, boff.p, nb, s, c->lc);
    u32 last_cnt = 0, last_off = 0;
    char last_byte = 0;
    c->lc.rb.get(&last_cnt, bcnt.p + nb - 1, 4, s);
    c->lc.rb.get(&last_off, boff.p + nb - 1, 4, s);
    c->lc.rb.get(&last_byte, c->text_ptr + n - 1, 1, s);
    c->lc.rb.sync(s);
    u32 terms = last_cnt + last_off;
    bool last_is_term = (last_byte == '\n' || last_byte == '\r');
    c->n_lines = terms + (last_is_term ? 0 : 1);
  }
  c->line_start.alloc((size_t)c->n_lines + 1, s);
  set_u32_kernel<<<1, 1, 0, s>>>(c->line_start.p, 0u);
  if (n) fill_line_starts_kernel<<<(nb + PT_SUB - 1) / PT_SUB, PT_THREADS, 0, s>>>(c->text_ptr, n, boff.p, c->line_start.p, nb);
  set_u32_kernel<<<1, 1, 0, s>>>(c->line_start.p + c->n_lines, (u32)n);
  c->lc.n += 3;

  const u32 nl = c->n_lines;
  const size_t nl1 = nl ? nl : 1;
  c->status.alloc(nl1, s);
  c->reads.alloc(nl1 * c->NW, s);
  c->idkey.alloc(nl1, s);
  DevBuf<u64> hk[2];
  DevBuf<u32> hv[2];
  for (int i = 0; i < 2; i++) { hk[i].alloc(nl1, s); hv[i].alloc(nl1, s); }
  if (nl) {
    c->lc.begin("parse_lines");
    CB_DISPATCH_NW(c->NW, (parse_lines_kernel<NW_><<<min(div_up(nl, 128), kNumSMs * 16), 128, 0, s>>>(
                              c->text_ptr, n, c->line_start.p, nl, c->K, c->L, c->status.p, c->reads.p, c->idkey.p,
                              hk[0].p, hv[0].p, (ull*)c->dcount.p)));
    c->lc.end("parse_lines");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "parse");
  download_counters(c);
  {
    // one gather: the counters, and the global line numbering (shards are consecutive line ranges in rank order)
    const u64 my_lines = c->n_lines;
    const Gathered G = reduce_counters(c, &my_lines, 1);
    c->line_bases.assign(world + 1, 0);
    for (int r = 0; r < world; r++) c->line_bases[r + 1] = c->line_bases[r] + G.at(r, DC_COUNT);
    if (c->line_bases[world] >= (1ull << 32)) throw Error(CB_E_CAPACITY, "more than 2^32 SFA lines");  // same on every rank
    c->line_base = (u32)c->line_bases[c->cfg.rank];
  }
  // decisions are taken on the global counters so that every rank takes the same branch
  if (c->gcount[DC_READS_GOOD] == 0) throw Error(CB_E_NOREADS, "No good reads");
  if (c->gcount[DC_LEN_MISMATCH])
    throw Error(CB_E_UNSUPPORTED, "reads of mixed length: " + std::to_string(c->gcount[DC_LEN_MISMATCH]) +
                                      " good reads differ from -readlen " + std::to_string(c->L));
  if (c->gcount[DC_LONG_ID]) throw Error(CB_E_UNSUPPORTED, "read ids longer than 16 bytes");

  // ---- duplicate detection ----
  timer_start(c, "dedupe");
  int bits = 0;
  while ((1ull << bits) < c->gcount[DC_READS_GOOD]) bits++;
  // hash bits that are sorted on: enough that equal masked hashes of different reads stay rare (about
  // 2^(bits-3) of the reads meet one, each costing one extra 64-bit compare in dedupe_runs), a whole number
  // of 9-bit digit passes -- three for the 12.8 M reads of config 2 (four measured 0.15 ms slower)
  int hb = ((bits + 3 + 8) / 9) * 9;
  if (hb > 64) hb = 64;
  const u64 keymask = hb == 64 ? ~0ull : ((1ull << hb) - 1);
  DevBuf<u32> surv_flag, surv_idx, cov;
  surv_flag.alloc(nl1, s);
  surv_idx.alloc(nl1, s);
  cov.alloc(nl1, s);
  CB_CUDA(cudaMemsetAsync(surv_flag.p, 0, nl1 * 4, s));
  u32 n_local_nodes = 0;
  if (!distributed(c)) {
    const u32 n_good = (u32)c->hcount[DC_READS_GOOD];
    int where = radix_sort(hk[0].p, hk[1].p, hv[0].p, hv[1].p, nl, 0, hb, s, c->lc, "dedupe");
    c->lc.begin("dedupe_runs");
    CB_DISPATCH_NW(c->NW, (dedupe_runs_kernel<NW_><<<min(div_up(n_good, 128), kNumSMs * 16), 128, 0, s>>>(
                              hk[where].p, hv[where].p, n_good, keymask, c->reads.p, c->idkey.p, c->L, surv_flag.p,
                              cov.p, (ull*)c->dcount.p)));
    c->lc.end("dedupe_runs");
    c->lc.n++;
    download_counters(c);
    n_local_nodes = (u32)c->hcount[DC_SURVIVORS];
  } else {
    // reads -> owner of their canonical hash -> (survivor, coverage) -> back to the line's rank
    const int NW = c->NW;
    const size_t rec_words = 4 + NW;
    timer_start(c, "dd_route");
    DevBuf<u64> recs;
    DevBuf<u32> dest;
    recs.alloc(nl1 * rec_words, s);
    dest.alloc(nl1, s);
    if (nl) {
      CB_DISPATCH_NW(NW, (dd_pack_kernel<NW_><<<div_up(nl, 256), 256, 0, s>>>(hk[0].p, c->reads.p, c->idkey.p, nl,
                                                                              c->line_base, world, recs.p, dest.p)));
      c->lc.n++;
    }
    DevBuf<char> got;
    size_t n_got = 0;
    route_records(c, recs.p, rec_words * 8, dest.p, nl, got, n_got, nullptr, 0, true);  // (a big exchange)
    recs.release();
    timer_stop(c, "dd_route");
    const size_t g1 = n_got ? n_got : 1;
    DevBuf<u64> rk[2], rreads;
    DevBuf<u32> rv[2], rline, rsurv, rcov;
    DevBuf<IdKey> rid;
    for (int i = 0; i < 2; i++) { rk[i].alloc(g1, s); rv[i].alloc(g1, s); }
    rline.alloc(g1, s); rsurv.alloc(g1, s); rcov.alloc(g1, s); rid.alloc(g1, s); rreads.alloc(g1 * NW, s);
    CB_CUDA(cudaMemsetAsync(rsurv.p, 0, g1 * 4, s));
    CB_CUDA(cudaMemsetAsync(rcov.p, 0, g1 * 4, s));
    int where = 0;
    guarded_defer(c, [&] {  // a rank-local sort failure travels with the next exchange's counts
      if (!n_got) return;
      CB_DISPATCH_NW(NW, (dd_unpack_kernel<NW_><<<div_up(n_got, 256), 256, 0, s>>>((const u64*)got.p, n_got, rk[0].p,
                                                                                  rv[0].p, rline.p, rid.p, rreads.p)));
      where = radix_sort(rk[0].p, rk[1].p, rv[0].p, rv[1].p, n_got, 0, hb, s, c->lc, "dedupe");
    });
    if (n_got && !failed_locally(c)) {
      c->lc.begin("dedupe_runs");
      CB_DISPATCH_NW(NW, (dedupe_runs_kernel<NW_><<<min(div_up(n_got, 128), kNumSMs * 16), 128, 0, s>>>(
                             rk[where].p, rv[where].p, (u32)n_got, keymask, rreads.p, rid.p, c->L, rsurv.p, rcov.p,
                             (ull*)c->dcount.p)));
      c->lc.end("dedupe_runs");
      c->lc.n += 2;
    }
    timer_start(c, "dd_return");
    DevBuf<u64> dbases;
    dbases.alloc(world + 1, s);
    c->lc.rb.put(dbases.p, c->line_bases.data(), (world + 1) * sizeof(u64), s);
    DevBuf<u32> ret, rdest;
    ret.alloc(g1 * 2, s);
    rdest.alloc(g1, s);
    if (n_got && !failed_locally(c)) {
      dd_return_kernel<<<div_up(n_got, 256), 256, 0, s>>>(rsurv.p, rcov.p, rline.p, n_got, dbases.p, world, ret.p,
                                                         rdest.p);
      c->lc.n++;
    }
    DevBuf<char> back;
    size_t n_back = 0;
    route_records(c, ret.p, 8, rdest.p, n_got, back, n_back);
    if (n_back) {
      dd_scatter_kernel<<<div_up(n_back, 256), 256, 0, s>>>((const u32*)back.p, n_back, c->line_base, surv_flag.p,
                                                           cov.p);
      c->lc.n++;
    }
    CB_CUDA(cudaGetLastError());
    timer_stop(c, "dd_return");
    download_counters(c);
    n_local_nodes = (u32)n_back;
  }
  const u64 my_nodes = n_local_nodes;
  const Gathered GN = reduce_counters(c, &my_nodes, 1);  // the counters and every rank's node count in one gather

  // ---- nodes of this shard, in line order ----
  exclusive_scan_u32(surv_flag.p, surv_idx.p, nl1, s, c->lc);
  const size_t nloc1 = n_local_nodes ? n_local_nodes : 1;
  DevBuf<u32> l_line, l_cov;
  DevBuf<u64> l_seq;
  DevBuf<IdKey> l_id;
  l_line.alloc(nloc1, s); l_cov.alloc(nloc1, s); l_seq.alloc(nloc1 * c->NW, s); l_id.alloc(nloc1, s);
  if (nl) {
    CB_DISPATCH_NW(c->NW, (build_nodes_kernel<NW_><<<div_up(nl, 256), 256, 0, s>>>(
                              surv_flag.p, surv_idx.p, nl, c->reads.p, c->idkey.p, cov.p, c->line_base, l_line.p,
                              l_seq.p, l_id.p, l_cov.p)));
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  if (!distributed(c)) {
    c->n_nodes = n_local_nodes;
    c->node_lo = 0;
    c->node_hi = n_local_nodes;
    c->node_bases.assign(2, 0);
    c->node_bases[1] = n_local_nodes;
    c->node_line = std::move(l_line);
    c->node_cov = std::move(l_cov);
    c->node_seq = std::move(l_seq);
    c->node_idkey = std::move(l_id);
  } else {
    // the node table (packed reads, id keys, coverage) is small and replicated on every rank, so
    // every later gather of a neighbour's sequence is a local HBM read
    std::vector<u64> counts(world, 0);
    for (int r = 0; r < world; r++) counts[r] = GN.at(r, DC_COUNT);
    u64 total = 0, lo = 0;
    c->node_bases.assign(world + 1, 0);
    for (int r = 0; r < world; r++) {
      if (r == c->cfg.rank) lo = total;
      total += counts[r];
      c->node_bases[r + 1] = total;
    }
    if (total >= (1ull << 27)) throw Error(CB_E_CAPACITY, "more than 2^27 nodes");
    c->n_nodes = (u32)total;
    c->node_lo = (u32)lo;
    c->node_hi = (u32)(lo + n_local_nodes);
    const size_t t1 = total ? total : 1;
    timer_start(c, "node_gather");
    c->node_line.alloc(t1, s); c->node_cov.alloc(t1, s); c->node_seq.alloc(t1 * c->NW, s); c->node_idkey.alloc(t1, s);
    // Native transport: the replicated table is gathered on a side stream while the main stream goes on with
    // keygen (which reads this shard's own nodes from own_seq / own_cov), the tuple exchange and the hash
    // partition; the join waits for the packed reads and coverage, the string stage for id keys and lines.
    const bool overlap = c->nccl_comm && c->side_stream;
    if (overlap) {
      // enqueued later (launch_node_gather: behind the tuple exchange of the overlap stage group, or by whoever
      // needs the table first); the send buffers stay with the context (keygen reads them too)
      c->own_seq = std::move(l_seq);
      c->own_cov = std::move(l_cov);
      c->own_line_tmp = std::move(l_line);
      c->own_id_tmp = std::move(l_id);
      c->gather_counts = counts;
      c->gather_deferred = true;
      c->gather_pending_seq = c->gather_pending_rest = true;
    } else {
      gather_all(c, l_seq.p, (size_t)c->NW * 8, n_local_nodes, counts, c->node_seq.p);
      gather_all(c, l_cov.p, 4, n_local_nodes, counts, c->node_cov.p);
      gather_all(c, l_line.p, 4, n_local_nodes, counts, c->node_line.p);
      gather_all(c, l_id.p, sizeof(IdKey), n_local_nodes, counts, c->node_idkey.p);
    }
    timer_stop(c, "node_gather");
  }
  timer_stop(c, "dedupe");

  // counters with the reference's names (BrushAssembler.java:266-289); global over the ranks
  auto& C = c->counters;
  C["reads_good"] = (int64_t)c->gcount[DC_READS_GOOD];
  C["reads_goodbp"] = (int64_t)c->gcount[DC_READS_GOODBP];
  C["reads_short"] = (int64_t)c->gcount[DC_READS_SHORT];
  C["reads_skipped"] = (int64_t)c->gcount[DC_READS_SKIPPED];
  C["input_lines_invalid"] = (int64_t)c->gcount[DC_LINES_INVALID];
  C["nodecount"] = (int64_t)c->gcount[DC_READS_GOOD];
  C["contained"] = (int64_t)(c->gcount[DC_READS_GOOD] - c->n_nodes);
  C["redundant"] = C["contained"];
  C["nodes"] = (int64_t)c->n_nodes;
  c->counters["host_collectives"] = c->n_host_coll;
  // free what later stages do not need (node arrays carry everything forward)
  c->reads.release();
  c->idkey.release();
  c->status.release();
}

}  // namespace cb
