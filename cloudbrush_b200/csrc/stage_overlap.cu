// stage_overlap.cu -- k-mer key generation, shuffle replacement, per-k-mer join fused with the
// exact overlap check, and the symmetric edge set of checkpoint A (01-overlap).
//
// Replaces (reference):
//   MatchPrefixMapper.map          MatchPrefix.java:107-193   keygen_kernel
//   Hadoop shuffle/sort/group      MatchPrefix.java:468       radix_sort (primitives.cu)
//   MatchPrefixReducer.reduce      MatchPrefix.java:241-431   join_verify_kernel (candidate part)
//   VerifyOverlapMapper/Reducer    VerifyOverlap.java:103-133, :203-344   join_verify_kernel (check part)
//   GenReverseEdgeMapper/Reducer   GenReverseEdge.java:54-81, :149-244    mirror records + sort + unique
//   BuildHighKmerList              BuildHighKmerList.java:50-117          "hkmer" counter (by-product)
//
// Every candidate (a,x,ov) -> (b,y) of the reference is produced inside ONE k-mer group (the
// group of b^y[0:K]), so the maximal-overlap filter of VerifyOverlap (:275-283) can be applied
// inside the group and candidates never have to be materialised: the join emits verified edges
// and their GenReverseEdge mirrors directly (SURVEY.md 8a').
#include "ctx.cuh"

namespace cb {

template <int NW>
__device__ __forceinline__ void load_read_b(const u64* __restrict__ reads, u32 idx, u64 (&r)[NW]) {
  const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    uint4 v = __ldg(row + i);
    r[2 * i] = (u64)v.x | ((u64)v.y << 32);
    r[2 * i + 1] = (u64)v.z | ((u64)v.w << 32);
  }
}

// ---------------------------------------------------------------------------------------------
// K3: warp-cooperative canonical k-mer key generator.  One warp per node, one lane per window:
// the packed read is loaded once (broadcast), every lane funnel-shifts its window out of the
// words, reverse-complements it in registers, and the warp stores 32 consecutive 8-byte keys
// and 4-byte payloads (fully coalesced).  Slot of (node, window i) is node*W + i; windows that
// emit nothing (MatchPrefix.java:163,170: canonical all-A interior windows) get the all-ones key
// and sort to the end.
//   payload meta = node << (pb+1) | i << 1 | dir      dir 0: window < rc(window) ('f'), 1: 'r'
//   key = canonical k-mer | (meta >> 32) << 2K        (bits above 2K are not sorted on)
// ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(256) keygen_kernel(const u64* __restrict__ node_seq,
                                                     const u32* __restrict__ node_cov, u32 node_lo, u32 n_nodes,
                                                     int L, int K, int W, int pb, u64* __restrict__ keys,
                                                     u32* __restrict__ vals, ull* __restrict__ dcount) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  u32 invalid = 0;
  // nodes [node_lo, node_lo + n_nodes) are this rank's (all of them on a single GPU)
  for (u32 loc = warp_global; loc < n_nodes; loc += nwarps) {
    const u32 node = node_lo + loc;
    u64 r[NW];
    load_read_b<NW>(node_seq, node, r);
    for (int w0 = 0; w0 < W; w0 += 32) {
      int i = w0 + lane;
      if (i < W) {
        u64 km = extract_kmer<NW>(r, i, K);
        u64 rk = revcomp_kmer(km, K);
        bool fwd = km < rk;
        u64 canon = fwd ? km : rk;
        bool interior = (i >= 1) && (i <= L - K - 1);
        bool emit = (km != rk) && !(interior && canon == 0);
        if (canon == 0 && i < L - K) atomicAdd(&dcount[DC_POLYA_WEIGHT], (ull)node_cov[node]);
        u64 meta = ((u64)node << (pb + 1)) | ((u64)i << 1) | (fwd ? 0ull : 1ull);
        size_t slot = (size_t)loc * W + i;
        keys[slot] = emit ? (canon | ((meta >> 32) << (2 * K))) : ~0ull;
        vals[slot] = emit ? (u32)meta : ~0u;
        if (!emit) invalid++;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
  if (lane == 0 && invalid) atomicAdd(&dcount[DC_TUPLES_INVALID], (ull)invalid);
}

// ---------------------------------------------------------------------------------------------
// K6+K7+K8+K9: segmented join fused with the packed-word overlap check and with the
// GenReverseEdge mirror.
//
// Node entry (b, y) -- the window-0 tuple of b for y = f, the window-(L-K) tuple for y = r --
// lives in exactly one k-mer group, and every verified edge (a,x,ov) -> (b,y) of that group has
// its mirror (b, flip y) -> (a, flip x) in the SAME adjacency list (b, flip y).  The join
// therefore writes the mirror records straight into that list's slot range (sized by
// slot_count_kernel: one slot per tuple of the group).  The edge itself belongs to list (a, x);
// that list is produced by a's own node entry (a, flip x) iff the mirror is proposed and kept
// there, which is decided locally (see `certain` below).  Only edges for which that is not
// certain go through the small side buffer X; duplicates are dropped when the lists are
// assembled.
// ---------------------------------------------------------------------------------------------
struct JoinParams {
  const u64* keys;
  const u32* vals;
  u32 T;               // valid tuples (sorted prefix)
  int L, K, pb;
  u64 kmask;           // 2K one-bits
  const u64* node_seq;
  const u32* node_cov;
  u32 max_cov;
  long long low_kmer, up_kmer;
  EdgeLayout lay;
  u32* slotcnt;        // slot_count_kernel output [n_lists]
  const u32* off1;     // slot offsets of the temporary list storage
  u64* mtmp;           // temporary list storage
  u32* cntM;           // records written per list
  u64* xbuf;           // side buffer
  u64 xcap;
  ull* dcount;
};

constexpr int JN_WARPS = 8;
constexpr int JN_EB = 256;  // staged side-buffer records per warp

__device__ __forceinline__ u64 tuple_meta(const JoinParams& P, u32 j) {
  return ((P.keys[j] >> (2 * P.K)) << 32) | (u64)P.vals[j];
}

// (x, ov) of tuple (pos, d) as an elist entry of list sel (0: canonical key c, 1: rc(c))
//   MatchPrefix.java:297-317 (NODEMSG: ov = len under c, K under rc(c) -- including the
//   swapped sizes for tags r1/f1) and :322-348 (SUFFIXMSG).
__device__ __forceinline__ void entry_of(int L, int K, int pos, int d, int sel, int& x, int& ov) {
  int ovc;
  if (pos == 0 || pos == L - K) ovc = L;
  else ovc = d == 0 ? L - pos : K + pos;
  x = sel ? (d ^ 1) : d;
  ov = sel ? (L + K - ovc) : ovc;
}

// walks the k-mer groups whose head lies in this warp's 32-tuple chunks
#define CB_FOR_EACH_SEGMENT(P, ...)                                                          \
  {                                                                                          \
    const u32 nchunks_ = ((P).T + 31) / 32;                                                  \
    for (u32 chunk_ = warp_global; chunk_ < nchunks_; chunk_ += nwarps) {                    \
      u32 i_ = chunk_ * 32 + lane;                                                           \
      bool valid_ = i_ < (P).T;                                                              \
      u64 ki_ = valid_ ? ((P).keys[i_] & (P).kmask) : 0;                                     \
      u64 kprev_ = (valid_ && i_ > 0) ? ((P).keys[i_ - 1] & (P).kmask) : ~ki_;               \
      u32 heads_ = __ballot_sync(0xffffffffu, valid_ && ki_ != kprev_);                      \
      while (heads_) {                                                                       \
        int h_ = __ffs(heads_) - 1;                                                          \
        heads_ &= heads_ - 1;                                                                \
        const u32 s = chunk_ * 32 + h_;                                                      \
        const u64 kseg = __shfl_sync(0xffffffffu, ki_, h_);                                  \
        u32 e = s + 1;                                                                       \
        for (;;) {                                                                           \
          u32 j_ = e + lane;                                                                 \
          bool diff_ = (j_ >= (P).T) || (((P).keys[j_ < (P).T ? j_ : 0] & (P).kmask) != kseg); \
          u32 b_ = __ballot_sync(0xffffffffu, diff_);                                        \
          if (b_) { e += __ffs(b_) - 1; break; }                                             \
          e += 32;                                                                           \
        }                                                                                    \
        __VA_ARGS__                                                                               \
      }                                                                                      \
    }                                                                                        \
  }

// one slot per tuple of the group for every node entry of a group the reducer processes
__global__ void __launch_bounds__(JN_WARPS * 32) slot_count_kernel(JoinParams P) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const u32 posmask = (1u << P.pb) - 1;
  CB_FOR_EACH_SEGMENT(P, {
    (void)kseg;
    const u32 g = e - s;
    if ((long long)g > P.low_kmer) {
      for (u32 j = s + lane; j < e; j += 32) {
        u64 m = tuple_meta(P, j);
        int pos = (int)((m >> 1) & posmask);
        if (pos == 0 || pos == P.L - P.K) {
          u32 b = (u32)(m >> (P.pb + 1));
          u32 y = pos == 0 ? 0u : 1u;
          P.slotcnt[2 * b + (y ^ 1u)] = g;
        }
      }
    }
  })
}

template <int NW>
__global__ void __launch_bounds__(JN_WARPS * 32) join_verify_kernel(JoinParams P) {
  __shared__ u64 s_B[JN_WARPS][32][NW];
  __shared__ u32 s_Bmeta[JN_WARPS][32];
  __shared__ u32 s_Bbase[JN_WARPS][32];
  __shared__ u32 s_Bcnt[JN_WARPS][32];
  __shared__ u64 s_edges[JN_WARPS][JN_EB];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L, K = P.K;
  const u32 posmask = (1u << P.pb) - 1;
  u32 ebuf = 0;        // staged side-buffer records (warp uniform)
  u32 n_verified = 0;  // lane-local count of maximal verified candidates

  auto flush = [&]() {
    if (ebuf == 0) return;
    ull base = 0;
    if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_RAW], (ull)ebuf);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (u32 t = lane; t < ebuf; t += 32)
      if (base + t < P.xcap) P.xbuf[base + t] = s_edges[warp][t];
    __syncwarp();
    ebuf = 0;
  };

  CB_FOR_EACH_SEGMENT(P, {
    const u32 g = e - s;
    // BuildHighKmerList.java:96-117: weighted count over windows [0, L-K) above UP_KMER
    if (kseg != 0 && (long long)g * (long long)P.max_cov > P.up_kmer) {
      ull sum = 0;
      for (u32 j = s + lane; j < e; j += 32) {
        u64 m = tuple_meta(P, j);
        int pos = (int)((m >> 1) & posmask);
        if (pos < L - K) sum += P.node_cov[(u32)(m >> (P.pb + 1))];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0 && (long long)sum > P.up_kmer) atomicAdd(&P.dcount[DC_HKMER], 1ull);
    }
    if ((long long)g <= P.low_kmer) continue;  // MatchPrefix.java:366
    if ((long long)g > P.up_kmer && lane == 0) atomicAdd(&P.dcount[DC_UPKMER_LISTS], 1ull);

    u32 scan_pos = s;
    for (;;) {
      // ---- collect the next batch of <= 32 node entries (tuples of window 0 / L-K) ----
      u32 nB = 0;
      while (scan_pos < e && nB < 32) {
        u32 j = scan_pos + lane;
        u64 m = (j < e) ? tuple_meta(P, j) : 0;
        int pos = (int)((m >> 1) & posmask);
        bool isn = (j < e) && (pos == 0 || pos == L - K);
        u32 bal = __ballot_sync(0xffffffffu, isn);
        u32 cnt = __popc(bal);
        u32 advance = 32;
        if (nB + cnt > 32) {
          u32 need = 32 - nB, tb = bal;
          int last = 0;
          for (u32 t = 0; t < need; t++) { last = __ffs(tb) - 1; tb &= tb - 1; }
          bal &= (last == 31) ? 0xffffffffu : ((2u << last) - 1);
          cnt = need;
          advance = (u32)last + 1;
        }
        if (isn && ((bal >> lane) & 1u)) {
          u32 slot = nB + __popc(bal & lt_mask);
          u32 b = (u32)(m >> (P.pb + 1));
          int d = (int)(m & 1);
          int y = pos == 0 ? 0 : 1;          // prefix entry: b^f starts with X; suffix entry: b^r
          int sel = pos == 0 ? d : (d ^ 1);  // MatchPrefix.java:275-296
          u64 r[NW], rcw[NW];
          load_read_b<NW>(P.node_seq, b, r);
          if (y) revcomp_read<NW>(r, rcw, L);
#pragma unroll
          for (int w = 0; w < NW; w++) s_B[warp][slot][w] = y ? rcw[w] : r[w];
          s_Bmeta[warp][slot] = b | ((u32)y << 28) | ((u32)sel << 29);
          s_Bbase[warp][slot] = P.off1[2 * b + (u32)(y ^ 1)];
          s_Bcnt[warp][slot] = 0;
        }
        nB += cnt;
        scan_pos += advance;
      }
      if (nB == 0) break;
      __syncwarp();
      // ---- every tuple of the group against the batch ----
      for (u32 j0 = s; j0 < e; j0 += 32) {
        u32 j = j0 + lane;
        bool act = j < e;
        u64 m = act ? tuple_meta(P, j) : 0;
        u32 a = (u32)(m >> (P.pb + 1));
        int pos = (int)((m >> 1) & posmask);
        int d = (int)(m & 1);
        u64 R[NW], Rc[NW];
        if (act) load_read_b<NW>(P.node_seq, a, R);
        else {
#pragma unroll
          for (int w = 0; w < NW; w++) R[w] = 0;
        }
        revcomp_read<NW>(R, Rc, L);
        // last k-mers of a^f and a^r: the group of the mirror candidate is keyed by their rc
        const u64 ksuf_f = extract_kmer<NW>(R, L - K, K), ksuf_r = extract_kmer<NW>(Rc, L - K, K);
        // does a neighbouring tuple of the group belong to the same read?  (tuples of one read
        // are contiguous: the sort is stable and keygen emits in (node, window) order)
        bool same_l = act && j > s && (u32)(tuple_meta(P, j - 1) >> (P.pb + 1)) == a;
        bool same_r = act && j + 1 < e && (u32)(tuple_meta(P, j + 1) >> (P.pb + 1)) == a;
        for (u32 n = 0; n < nB; n++) {
          u32 bm = s_Bmeta[warp][n];
          u32 b = bm & 0x0fffffffu;
          int y = (bm >> 28) & 1, sel = (bm >> 29) & 1;
          int x, ov;
          entry_of(L, K, pos, d, sel, x, ov);
          u64 A[NW], B[NW];
#pragma unroll
          for (int w = 0; w < NW; w++) { A[w] = x ? Rc[w] : R[w]; B[w] = s_B[warp][n][w]; }
          shl_bases<NW>(A, L - ov);
          // VerifyOverlap.java:295-309: a^x[L-ov:] == b^y[0:ov]; self candidates skipped at
          // MatchPrefix.java:405-407
          bool ok = act && a != b && prefix_equal<NW>(A, B, ov);
          if (ok && (same_l || same_r)) {
            // maximal-overlap filter, VerifyOverlap.java:275-283: a larger verified overlap of
            // the same (a, x) with the same (b, y) wins; an identical entry of an earlier tuple
            // (MatchPrefix.java:409 hasEdge) is emitted only once
            for (int dirn = -1; dirn <= 1 && ok; dirn += 2) {
              long long jj = (long long)j + dirn;
              while (jj >= (long long)s && jj < (long long)e) {
                u64 m2 = tuple_meta(P, (u32)jj);
                if ((u32)(m2 >> (P.pb + 1)) != a) break;
                int x2, ov2;
                entry_of(L, K, (int)((m2 >> 1) & posmask), (int)(m2 & 1), sel, x2, ov2);
                if (x2 == x && ov2 == ov && dirn < 0) { ok = false; break; }
                if (x2 == x && ov2 > ov) {
                  u64 A2[NW];
#pragma unroll
                  for (int w = 0; w < NW; w++) A2[w] = x ? Rc[w] : R[w];
                  shl_bases<NW>(A2, L - ov2);
                  if (prefix_equal<NW>(A2, B, ov2)) { ok = false; break; }
                }
                jj += dirn;
              }
            }
          }
          u32 bal = __ballot_sync(0xffffffffu, ok);
          if (bal) {
            u32 cnt = __popc(bal);
            u32 rank = __popc(bal & lt_mask);
            u32 mbase = s_Bbase[warp][n] + s_Bcnt[warp][n];
            // Is the mirror candidate (b, flip y, ov) -> (a, flip x) certain to be proposed and
            // kept in the group of X' = a^(flip x)[0:K]?  Then a's own node entry writes this
            // edge into list (a, x) and it need not travel.  (SURVEY.md 8a': proposed iff
            // ov > K and X' not poly-A/T, or ov == K and X' > rc(X'); kept iff no larger TRUE
            // overlap, which would show as another tuple of a in this group.)
            bool certain = false;
            if (ok) {
              const u64 ksuf = x ? ksuf_r : ksuf_f;
              const u64 xp = revcomp_kmer(ksuf, K);
              const bool proposed = (ov > K) ? ((xp < ksuf ? xp : ksuf) != 0) : (xp > ksuf);
              certain = proposed && kseg != 0 && !same_l && !same_r && P.low_kmer == 1;
            }
            u32 xbal = __ballot_sync(0xffffffffu, ok && !certain);
            u32 xcnt = __popc(xbal);
            if (xcnt && ebuf + xcnt > JN_EB) flush();
            if (ok) {
              P.mtmp[mbase + rank] = P.lay.encode(b, (u32)(y ^ 1), (u32)ov, (u32)(x ^ 1), a);  // GenReverseEdge
              if (!certain) s_edges[warp][ebuf + __popc(xbal & lt_mask)] = P.lay.encode(a, (u32)x, (u32)ov, (u32)y, b);
              n_verified++;
            }
            ebuf += xcnt;
            __syncwarp();
            if (lane == 0) s_Bcnt[warp][n] += cnt;
            __syncwarp();
          }
        }
      }
      __syncwarp();
      if ((u32)lane < nB) {
        u32 bm = s_Bmeta[warp][lane];
        u32 b = bm & 0x0fffffffu, y = (bm >> 28) & 1;
        P.cntM[2 * b + (y ^ 1u)] = s_Bcnt[warp][lane];
      }
      __syncwarp();
      if (scan_pos >= e) break;
    }
  })
  flush();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_verified += __shfl_xor_sync(0xffffffffu, n_verified, o);
  if (lane == 0 && n_verified) {
    atomicAdd(&P.dcount[DC_CAND_VERIFIED], (ull)n_verified);
    atomicAdd(&P.dcount[DC_EDGES_M], (ull)n_verified);
  }
}

// ---------------------------------------------------------------------------------------------
// Side-buffer records: edge (a,x,ov,y,b) goes into list (a,x) unless a's own node entry already
// wrote it.  One thread per record: scan the list's own records, then claim a free slot.  A
// record that does not fit (only possible around poly-A/T k-mers, whose interior windows have no
// tuple and therefore no slot) is counted in extra[list]; the host then re-runs the join with
// enlarged slot ranges.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) side_insert_kernel(const u64* __restrict__ xs, size_t n, EdgeLayout lay,
                                                          const u32* __restrict__ off, const u32* __restrict__ cntM,
                                                          u32* __restrict__ cnt, u64* __restrict__ keys,
                                                          u32* __restrict__ extra, ull* __restrict__ dcount) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool inserted = false, overflow = false;
  if (i < n) {
    const u64 key = xs[i];
    const u32 lid = lay.list(key);
    const u32 base = off[lid], m = cntM[lid], cap = off[lid + 1] - base;
    bool dup = false;
    for (u32 t = 0; t < m; t++)
      if (keys[base + t] == key) { dup = true; break; }
    if (!dup) {
      u32 pos = atomicAdd(&cnt[lid], 1u);
      if (pos < cap) {
        keys[base + pos] = key;
        inserted = true;
      } else {
        atomicAdd(&extra[lid], 1u);
        overflow = true;
      }
    }
  }
  u32 ni = __popc(__ballot_sync(0xffffffffu, inserted)), no = __popc(__ballot_sync(0xffffffffu, overflow));
  if ((threadIdx.x & 31) == 0) {
    if (ni) atomicAdd(&dcount[DC_EDGES_A], (ull)ni);
    if (no) atomicAdd(&dcount[DC_MAX_LIST], (ull)no);  // reused as the overflow count of this attempt
  }
}

// owner rank of adjacency list (a, x): the rank of the k-mer group of node entry (a, flip x),
// i.e. of canon(last k-mer of a) for x = f and canon(first k-mer of a) for x = r
template <int NW>
__device__ __forceinline__ u32 list_owner(const u64* __restrict__ node_seq, u32 lid, int L, int K, int pbit,
                                          u32 wmask) {
  u64 r[NW];
  load_read_b<NW>(node_seq, lid >> 1, r);
  u64 km = extract_kmer<NW>(r, (lid & 1u) ? 0 : L - K, K);
  u64 rk = revcomp_kmer(km, K);
  u64 canon = km < rk ? km : rk;
  return (u32)(canon >> pbit) & wmask;
}

template <int NW>
__global__ void __launch_bounds__(256) key_owner_kernel(const u64* __restrict__ keys, size_t n, EdgeLayout lay,
                                                        const u64* __restrict__ node_seq, int L, int K, int pbit,
                                                        u32 wmask, u32* __restrict__ dest) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dest[i] = list_owner<NW>(node_seq, lay.list(keys[i]), L, K, pbit, wmask);
}

__global__ void __launch_bounds__(256) tuple_dest_hist_kernel(const u64* __restrict__ keys, size_t n, int pbit,
                                                              u32 wmask, ull* __restrict__ counts) {
  __shared__ u32 h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(&h[(u32)(keys[i] >> pbit) & wmask], 1u);
  __syncthreads();
  if (threadIdx.x <= wmask && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (ull)h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) count_invalid_kernel(const u64* __restrict__ keys, size_t n,
                                                            ull* __restrict__ out) {
  u32 c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (keys[i] == ~0ull) c++;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (ull)c);
}

__global__ void __launch_bounds__(256) add_u32_kernel(u32* __restrict__ a, const u32* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

static int bits_for(u64 nvalues) {  // bits needed to store values 0 .. nvalues-1
  int b = 1;
  while ((1ull << b) < nvalues) b++;
  return b;
}

#define CB_DISPATCH_NW(nw, CALL)                                             \
  switch (nw) {                                                              \
    case 2: { constexpr int NW_ = 2; CALL; } break;                          \
    case 4: { constexpr int NW_ = 4; CALL; } break;                          \
    case 6: { constexpr int NW_ = 6; CALL; } break;                          \
    case 8: { constexpr int NW_ = 8; CALL; } break;                          \
    default: throw Error(CB_E_UNSUPPORTED, "read length above 256 bases");   \
  }

static u32 read_u32(cb_ctx* c, const u32* p) {
  u32 v = 0;
  CB_CUDA(cudaMemcpyAsync(&v, p, 4, cudaMemcpyDeviceToHost, c->stream));
  CB_CUDA(cudaStreamSynchronize(c->stream));
  return v;
}

void stage_build_overlap(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const u32 nn = c->n_nodes;
  const int L = c->L, K = c->K, W = c->W;
  const int pb = bits_for((u64)W);
  const int nbits = bits_for((u64)nn);
  if (nbits > 27) throw Error(CB_E_CAPACITY, "more than 2^27 nodes per context");
  if (nbits + pb + 1 > 32 + (64 - 2 * K))
    throw Error(CB_E_CAPACITY, "node/window payload does not fit the 12-byte tuple for this k");
  const u32 n_own = c->node_hi - c->node_lo;
  const int world = c->cfg.world;
  const u32 wmask = (u32)world - 1;
  size_t Tmax = (size_t)n_own * W;  // tuple slots generated here
  if (Tmax >= (1ull << 32)) throw Error(CB_E_CAPACITY, "more than 2^32 k-mer tuples per context");
  c->lay.nb = nbits;
  c->lay.ob = bits_for((u64)(L - K + 1));
  c->lay.L = L;
  const u32 n_lists = 2 * nn;

  // ---- keygen ----
  timer_start(c, "keygen");
  DevBuf<u64> keys[2];
  DevBuf<u32> vals[2];
  for (int i = 0; i < 2; i++) { keys[i].alloc(Tmax ? Tmax : 1, s); vals[i].alloc(Tmax ? Tmax : 1, s); }
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_TUPLES_INVALID, 0, (DC_COUNT - DC_TUPLES_INVALID) * sizeof(u64), s));
  if (n_own) {
    int blocks = div_up((size_t)n_own * 32, 256);
    int cap = kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    c->lc.begin("keygen");
    CB_DISPATCH_NW(c->NW, (keygen_kernel<NW_><<<blocks, 256, 0, s>>>(c->node_seq.p, c->node_cov.p, c->node_lo, n_own, L,
                                                                    K, W, pb, keys[0].p, vals[0].p,
                                                                    (ull*)c->dcount.p)));
    c->lc.end("keygen");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "keygen");

  // ---- shuffle replacement ----
  if (distributed(c)) {
    // partition by owner rank (one stable radix pass over the owner bits of the key), then the
    // all-to-all that replaces the MapReduce shuffle.  Received order is (source rank, node,
    // window), i.e. ascending (node, window): the join relies on it.
    timer_start(c, "exchange_tuples");
    int w0 = radix_sort(keys[0].p, keys[1].p, vals[0].p, vals[1].p, Tmax, c->pbit, c->pbit + c->lg_world, s, c->lc,
                        "partition");
    DevBuf<u64> dcnt;
    dcnt.alloc(world, s);
    CB_CUDA(cudaMemsetAsync(dcnt.p, 0, world * sizeof(u64), s));
    if (Tmax) {
      int blocks = div_up(Tmax, 256 * 16);
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      tuple_dest_hist_kernel<<<blocks, 256, 0, s>>>(keys[w0].p, Tmax, c->pbit, wmask, (ull*)dcnt.p);
      c->lc.n++;
    }
    std::vector<u64> cnt(world), rcnt(world), sb(world), rb(world);
    CB_CUDA(cudaMemcpyAsync(cnt.data(), dcnt.p, world * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    ex_counts(c, cnt.data(), rcnt.data(), 1);
    size_t Trecv = 0;
    for (int r = 0; r < world; r++) Trecv += (size_t)rcnt[r];
    if (Trecv >= (1ull << 32)) throw Error(CB_E_CAPACITY, "more than 2^32 k-mer tuples per context");
    DevBuf<u64> rkeys;
    DevBuf<u32> rvals;
    rkeys.alloc(Trecv ? Trecv : 1, s);
    rvals.alloc(Trecv ? Trecv : 1, s);
    for (int r = 0; r < world; r++) { sb[r] = cnt[r] * 8; rb[r] = rcnt[r] * 8; }
    ex_alltoallv(c, keys[w0].p, sb.data(), rkeys.p, rb.data());
    for (int r = 0; r < world; r++) { sb[r] = cnt[r] * 4; rb[r] = rcnt[r] * 4; }
    ex_alltoallv(c, vals[w0].p, sb.data(), rvals.p, rb.data());
    Tmax = Trecv;
    keys[0] = std::move(rkeys);
    vals[0] = std::move(rvals);
    keys[1].alloc(Tmax ? Tmax : 1, s);
    vals[1].alloc(Tmax ? Tmax : 1, s);
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_TUPLES_INVALID, 0, sizeof(u64), s));
    if (Tmax) {
      int blocks = div_up(Tmax, 256 * 16);
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      count_invalid_kernel<<<blocks, 256, 0, s>>>(keys[0].p, Tmax, (ull*)c->dcount.p + DC_TUPLES_INVALID);
      c->lc.n++;
    }
    timer_stop(c, "exchange_tuples");
  }
  timer_start(c, "sort_tuples");
  int where = radix_sort(keys[0].p, keys[1].p, vals[0].p, vals[1].p, Tmax, 0, 2 * K, s, c->lc, "tuples");
  timer_stop(c, "sort_tuples");
  download_counters(c);
  const u32 T = (u32)(Tmax - c->hcount[DC_TUPLES_INVALID]);
  {
    u64 tg = T;
    ex_allreduce(c, &tg, 1, 0);
    c->counters["tuples"] = (int64_t)tg;
  }
  keys[where ^ 1].release();
  vals[where ^ 1].release();

  // ---- join + verify + mirror ----
  timer_start(c, "join_verify");
  JoinParams P;
  P.keys = keys[where].p;
  P.vals = vals[where].p;
  P.T = T;
  P.L = L; P.K = K; P.pb = pb;
  P.kmask = (2 * K == 64) ? ~0ull : ((1ull << (2 * K)) - 1);
  P.node_seq = c->node_seq.p;
  P.node_cov = c->node_cov.p;
  P.max_cov = (u32)c->hcount[DC_MAX_COV];
  P.low_kmer = c->cfg.low_kmer;
  P.up_kmer = c->cfg.up_kmer;
  P.lay = c->lay;
  P.dcount = (ull*)c->dcount.p;
  int jblocks = kNumSMs * 8;
  {
    int need = div_up(((size_t)T + 31) / 32, JN_WARPS);
    if (need < jblocks) jblocks = need > 0 ? need : 1;
  }
  // slots: one per tuple of the group for every node entry
  DevBuf<u32> slotcnt, off1, cntM;
  slotcnt.alloc((size_t)n_lists + 1, s);
  off1.alloc((size_t)n_lists + 1, s);
  cntM.alloc(n_lists, s);
  CB_CUDA(cudaMemsetAsync(slotcnt.p, 0, ((size_t)n_lists + 1) * 4, s));
  CB_CUDA(cudaMemsetAsync(cntM.p, 0, (size_t)n_lists * 4, s));
  P.slotcnt = slotcnt.p;
  c->lc.begin("slot_count");
  slot_count_kernel<<<jblocks, JN_WARPS * 32, 0, s>>>(P);
  c->lc.end("slot_count");
  c->lc.n++;
  exclusive_scan_u32(slotcnt.p, off1.p, (size_t)n_lists + 1, s, c->lc);
  size_t slots1 = read_u32(c, off1.p + n_lists);
  DevBuf<u64> mtmp, xbuf;
  DevBuf<u32> cnt, extra;
  cnt.alloc(n_lists ? n_lists : 1, s);
  extra.alloc(n_lists ? n_lists : 1, s);
  P.cntM = cntM.p;
  u64 xcap = (u64)T / 4 + 4096;
  u64 n_x = 0;
  for (int attempt = 0;; attempt++) {
    mtmp.alloc(slots1 ? slots1 : 1, s);
    CB_CUDA(cudaMemsetAsync(mtmp.p, 0xff, (slots1 ? slots1 : 1) * sizeof(u64), s));  // free slots read as ~0
    xbuf.alloc(xcap, s);
    P.off1 = off1.p;
    P.mtmp = mtmp.p;
    P.xbuf = xbuf.p;
    P.xcap = xcap;
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_HKMER, 0, (DC_EDGE_REMOVAL - DC_HKMER) * sizeof(u64), s));
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_MAX_LIST, 0, sizeof(u64), s));
    c->lc.begin("join_verify");
    CB_DISPATCH_NW(c->NW, (join_verify_kernel<NW_><<<jblocks, JN_WARPS * 32, 0, s>>>(P)));
    c->lc.end("join_verify");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    download_counters(c);
    n_x = c->hcount[DC_EDGES_RAW];
    u64 x_over = n_x > xcap ? 1 : 0;
    ex_allreduce(c, &x_over, 1, 1);  // every rank must take the same branch: collectives follow
    if (x_over) {  // side buffer too small: its exact size is known now
      if (attempt >= 3) throw Error(CB_E_CAPACITY, "side buffer overflow");
      if (n_x > xcap) xcap = n_x;
      continue;
    }
    // side-buffer records into their lists (on the rank that owns the list)
    CB_CUDA(cudaMemcpyAsync(cnt.p, cntM.p, (size_t)n_lists * 4, cudaMemcpyDeviceToDevice, s));
    CB_CUDA(cudaMemsetAsync(extra.p, 0, (size_t)n_lists * 4, s));
    const u64* xin = xbuf.p;
    size_t n_xin = n_x;
    DevBuf<char> xrecv;
    if (distributed(c)) {
      DevBuf<u32> xdest;
      xdest.alloc(n_x ? n_x : 1, s);
      if (n_x) {
        CB_DISPATCH_NW(c->NW, (key_owner_kernel<NW_><<<div_up(n_x, 256), 256, 0, s>>>(
                                  xbuf.p, n_x, c->lay, c->node_seq.p, L, K, c->pbit, wmask, xdest.p)));
        c->lc.n++;
      }
      route_records(c, xbuf.p, 8, xdest.p, n_x, xrecv, n_xin);
      xin = (const u64*)xrecv.p;
    }
    if (n_xin) {
      c->lc.begin("side_insert");
      side_insert_kernel<<<div_up(n_xin, 256), 256, 0, s>>>(xin, n_xin, c->lay, off1.p, cntM.p, cnt.p, mtmp.p, extra.p,
                                                           (ull*)c->dcount.p);
      c->lc.end("side_insert");
      c->lc.n++;
      CB_CUDA(cudaGetLastError());
    }
    download_counters(c);
    u64 slot_over = c->hcount[DC_MAX_LIST];
    ex_allreduce(c, &slot_over, 1, 0);
    if (slot_over == 0) break;
    if (attempt >= 3) throw Error(CB_E_CAPACITY, "adjacency slot overflow");
    // enlarge the slot ranges of the lists that overflowed and run the join again
    add_u32_kernel<<<div_up((size_t)n_lists, 256), 256, 0, s>>>(slotcnt.p, extra.p, n_lists);
    c->lc.n++;
    exclusive_scan_u32(slotcnt.p, off1.p, (size_t)n_lists + 1, s, c->lc);
    slots1 = read_u32(c, off1.p + n_lists);
  }
  timer_stop(c, "join_verify");
  keys[where].release();
  vals[where].release();

  // BuildHighKmerList's counter (poly-A/T k-mer handled through its global weight)
  reduce_counters(c);
  int64_t hkmer = (int64_t)c->gcount[DC_HKMER] + ((long long)c->gcount[DC_POLYA_WEIGHT] > c->cfg.up_kmer ? 1 : 0);
  c->counters["hkmer"] = hkmer;
  c->counters["candidates_verified"] = (int64_t)c->gcount[DC_CAND_VERIFIED];
  c->counters["side_buffer_edges"] = (int64_t)c->gcount[DC_EDGES_RAW];
  c->counters["upkmer_truncated_lists"] = (int64_t)c->gcount[DC_UPKMER_LISTS];
  c->counters["order_dependent"] = (int64_t)c->gcount[DC_UPKMER_LISTS];
  if (hkmer > 0)
    throw Error(CB_E_UNSUPPORTED,
                "non-empty high-frequency k-mer list: the reference's stopword loader (MatchPrefix.java:81) is "
                "ill-defined for it, parity undefined");

  // checkpoint A: the slot array is the adjacency structure
  c->adj_off = std::move(off1);
  c->slots = slots1;
  c->ckA.keys = std::move(mtmp);
  c->ckA.cnt = std::move(cnt);
  c->ckA.n = (size_t)(c->hcount[DC_EDGES_M] + c->hcount[DC_EDGES_A]);  // the lists this rank owns
  c->ckA.valid = true;
  c->counters["edges_overlap"] = (int64_t)(c->gcount[DC_EDGES_M] + c->gcount[DC_EDGES_A]);
}

}  // namespace cb
