// stage_overlap.cu -- k-mer key generation, shuffle replacement, per-k-mer join fused with the
// exact overlap check, and the symmetric edge set of checkpoint A (01-overlap).
//
// Replaces (reference):
//   MatchPrefixMapper.map          MatchPrefix.java:107-193   keygen_kernel
//   Hadoop shuffle/sort/group      MatchPrefix.java:468       radix_sort (primitives.cu)
//   MatchPrefixReducer.reduce      MatchPrefix.java:241-431   join_verify_kernel (candidate part)
//   VerifyOverlapMapper/Reducer    VerifyOverlap.java:103-133, :203-344   join_verify_kernel (check part)
//   GenReverseEdgeMapper/Reducer   GenReverseEdge.java:54-81, :149-244    mirror records written in place (join_verify_kernel), side_insert_kernel
//   BuildHighKmerList              BuildHighKmerList.java:50-117          "hkmer" counter (by-product)
//
// Every candidate (a,x,ov) -> (b,y) of the reference is produced inside ONE k-mer group (the
// group of b^y[0:K]), so the maximal-overlap filter of VerifyOverlap (:275-283) can be applied
// inside the group and candidates never have to be materialised: the join emits verified edges
// and their GenReverseEdge mirrors directly (SURVEY.md 8a').
#include <cstdlib>
#include <cstring>

#include "ctx.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------
// K3: warp-cooperative canonical k-mer key generator.  One thread per (node, window) slot in
// slot order, so a warp covers 32 consecutive windows -- of one read or of several short ones
// (36 bp, k=21: two reads per warp; a lane-per-window-of-one-read layout left half the lanes
// idle).  The lanes of one read load the same 16..64-byte packed row (one broadcast request),
// funnel-shift their window out of the words, reverse-complement it in registers, and the warp
// stores 32 consecutive 8-byte keys and 4-byte payloads (fully coalesced).  Slot of (node,
// window i) is node*W + i; windows that emit nothing (MatchPrefix.java:163,170: canonical all-A
// interior windows) get the all-ones key and sort to the end.
//   payload meta = node << (pb+1) | i << 1 | dir      dir 0: window < rc(window) ('f'), 1: 'r'
//   key = canonical k-mer | (meta >> 32) << 2K        (bits above 2K are not sorted on)
// ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(256) keygen_kernel(const u64* __restrict__ node_seq,
                                                     const u32* __restrict__ node_cov, u32 node_lo, u32 n_nodes,
                                                     int L, int K, int W, int pb, u64* __restrict__ keys,
                                                     u32* __restrict__ vals, ull* __restrict__ dcount,
                                                     u32* __restrict__ cnt1, int shift1, int bits1) {
  // cnt1 != nullptr: the level-1 digit counts of the hash partition (primitives.cuh) are taken while the
  // tuples are written, so the first scatter pass needs no counting read of its own
  extern __shared__ u32 kg_hist[];
  const u32 bins1 = cnt1 ? (1u << bits1) : 0u;
  for (u32 t = threadIdx.x; t < bins1; t += blockDim.x) kg_hist[t] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const u32 total = n_nodes * (u32)W;  // < 2^32 (checked by the host)
  const u32 stride = gridDim.x * blockDim.x;
  u32 invalid = 0;
  // nodes [node_lo, node_lo + n_nodes) are this rank's (all of them on a single GPU)
  for (u32 slot0 = blockIdx.x * blockDim.x; slot0 < total; slot0 += stride) {
    const u32 slot = slot0 + threadIdx.x;
    if (slot < total) {
      const u32 loc = slot / (u32)W;
      const int i = (int)(slot - loc * (u32)W);
      const u32 node = node_lo + loc;
      // the two words the window straddles, straight from the row (L1-resident: the lanes of a
      // read share it; indexing a register copy of the row dynamically would put it in local memory)
      const u64* row = node_seq + (size_t)node * NW;
      const int w = i >> 5, off = (i & 31) * 2;
      const u64 hi = __ldg(row + w), lo = (w + 1 < NW) ? __ldg(row + w + 1) : 0ull;
      const u64 top = off ? ((hi << off) | (lo >> (64 - off))) : hi;
      u64 km = top >> (64 - 2 * K);
      u64 rk = revcomp_kmer(km, K);
      bool fwd = km < rk;
      u64 canon = fwd ? km : rk;
      bool interior = (i >= 1) && (i <= L - K - 1);
      bool emit = (km != rk) && !(interior && canon == 0);
      if (canon == 0 && i < L - K) atomicAdd(&dcount[DC_POLYA_WEIGHT], (ull)node_cov[node]);
      u64 meta = ((u64)node << (pb + 1)) | ((u64)i << 1) | (fwd ? 0ull : 1ull);
      keys[slot] = emit ? (canon | ((meta >> 32) << (2 * K))) : ~0ull;
      vals[slot] = emit ? (u32)meta : ~0u;
      if (!emit) invalid++;
      else if (cnt1) atomicAdd(&kg_hist[(u32)(kmer_hash(canon) >> shift1) & (bins1 - 1)], 1u);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
  if (lane == 0 && invalid) atomicAdd(&dcount[DC_TUPLES_INVALID], (ull)invalid);
  __syncthreads();
  for (u32 t = threadIdx.x; t < bins1; t += blockDim.x) {
    const u32 v = kg_hist[t];
    if (v) atomicAdd(&cnt1[t], v);
  }
}

// ---------------------------------------------------------------------------------------------
// K6+K7+K8+K9: per-k-mer join fused with the packed-word overlap check and with the
// GenReverseEdge mirror -- flat: one thread per k-mer tuple.
//
// Node entry (b, y) -- the window-0 tuple of b for y = f, the window-(L-K) tuple for y = r --
// lives in exactly one k-mer group, and every verified edge (a,x,ov) -> (b,y) of that group has
// its mirror (b, flip y) -> (a, flip x) in the SAME adjacency list (b, flip y).  That list gets
// one slot per tuple of the group: the record of the pair (node entry, tuple j) has the fixed
// position  off[list] + (j - first tuple of the group), so the join needs no ranking and no
// atomics, and a pair that does not verify simply leaves its slot empty (~0).
// The edge itself belongs to list (a, x); that list is produced by a's own node entry (a, flip x)
// iff the mirror is proposed and kept there, which is decided locally (see `certain`).  Only the
// other edges (~3 %) go through the side buffer and are inserted into a free slot of their list.
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
  u32* slotcnt;        // group_walk output [n_lists]: slots of the list (= tuples of the group)
  u32* seg_head;       // [T] first tuple of the tuple's group
  u32* ne_list;        // [T+1] positions s.. of a group: (list id | entry's dir bit << 31) of its node entries, ~0 ends
  const u32* off1;     // slot offsets of the lists
  // bucket path (bucket_join_kernel): tuples bucketed by k-mer hash, lists sized and placed per bucket
  const u32* bstart;   // [NB + 1] bucket offsets into keys / vals
  int lshift;          // local digit of a tuple = (kmer_hash >> lshift) & 1023
  uint2* adj;          // [n_lists] (first slot, slots) of every list that has a node entry
  const u32* extra;    // [n_lists] additional slots per list (retry after a slot overflow), or nullptr
  ull* slot_cursor;    // [0] slots handed out so far, [1] lists published so far
  ull slot_cap;        // slots allocated
  uint4* list_order;   // [n_lists] (list, first slot, slots) of the lists that own slots, in slot order
  u64* slots;          // list storage
  u32* ovh;            // [slots][nwo] overhang of the record in the slot
  int nwo;
  u64* xbuf;           // side buffer
  u64 xcap;
  ull* dcount;
};

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

// first index in [lo, hi) whose masked key is >= (UPPER: >) kseg (keys are sorted on the masked bits)
template <bool UPPER>
__device__ __forceinline__ u32 key_bound(const JoinParams& P, u64 kseg, u32 lo, u32 hi) {
  while (lo < hi) {
    const u32 mid = lo + ((hi - lo) >> 1);
    const u64 km = P.keys[mid] & P.kmask;
    if (UPPER ? km <= kseg : km < kseg) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// One flat pass over the sorted tuples, one thread per tuple: every tuple learns the first tuple
// of its k-mer group (block-wide running maximum of the group heads; a group that began before the
// block is found by binary search), and the node entries of a group the reducer processes
// (MatchPrefix.java:366) claim the first free positions of the group's own index range in
// `ne_list` (pre-filled with the terminator ~0) and size the adjacency list they produce: one
// slot per tuple of the group.  The end of a group is the next head, found the same way from the
// right (a serial forward scan by the node entries measured 2.3 ms; a group that runs past the
// block is followed in global memory).  BuildHighKmerList's counter and the
// UP_KMER flag are per group and are evaluated by the group's head.
constexpr int GW_THREADS = 256;
__global__ void __launch_bounds__(GW_THREADS) group_walk_kernel(JoinParams P) {
  __shared__ u32 s_wmax[GW_THREADS / 32], s_wmin[GW_THREADS / 32], s_wcnt[GW_THREADS / 32], s_first, s_last;
  __shared__ unsigned short s_pre[GW_THREADS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 posmask = (1u << P.pb) - 1;
  const int L = P.L, K = P.K;
  for (u32 j0 = blockIdx.x * GW_THREADS; j0 < P.T; j0 += gridDim.x * GW_THREADS) {  // block-uniform
    const u32 j = j0 + threadIdx.x;
    const bool valid = j < P.T;
    const u64 kj = valid ? P.keys[j] : 0ull;
    const u32 vj = valid ? P.vals[j] : 0u;
    const u64 kseg = kj & P.kmask;
    const u64 kprev = (valid && j > 0) ? (P.keys[j - 1] & P.kmask) : ~kseg;
    const bool head = valid && kseg != kprev;
    const u64 m = ((kj >> (2 * K)) << 32) | (u64)vj;
    const int pos = (int)((m >> 1) & posmask);
    const bool cand = valid && (pos == 0 || pos == L - K);  // a NODEMSG tuple
    // running maximum of the head positions (+1; 0 = no head yet in this block)
    u32 h = head ? j + 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, h, o);
      if (lane >= o) h = max(h, t);
    }
    // ... and the running minimum, from the right, of the head positions: where the next group begins
    u32 nx = head ? j : 0xffffffffu;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_down_sync(0xffffffffu, nx, o);
      if (lane + o < 32) nx = min(nx, t);
    }
    const u32 cbal = __ballot_sync(0xffffffffu, cand);
    if (lane == 31) s_wmax[warp] = h;
    if (lane == 0) { s_wmin[warp] = nx; s_wcnt[warp] = (u32)__popc(cbal); }
    // the two groups that may extend beyond the block are followed by one warp each, 32 keys per
    // (coalesced) load: a per-thread binary search here cost 1.2 ms in dependent L2 round trips
    if (warp == 0) {  // where does the group of the block's first tuple begin?
      const u64 kf = __shfl_sync(0xffffffffu, kseg, 0);
      u32 sfirst = 0xffffffffu, hi = j0;  // keys[hi .. j0] all equal kf
      for (int rounds = 0; sfirst == 0xffffffffu; rounds++) {
        if (hi == 0) { sfirst = 0; break; }
        if (rounds == 4) { sfirst = key_bound<false>(P, kf, 0, hi); break; }
        const u32 base = hi >= 32 ? hi - 32 : 0;
        const u32 idx = base + lane;
        const bool differs = idx < hi && (P.keys[idx] & P.kmask) != kf;
        const u32 bal = __ballot_sync(0xffffffffu, differs);
        if (bal) sfirst = base + (31 - __clz(bal)) + 1;
        else hi = base;
      }
      if (lane == 0) s_first = sfirst;
    }
    if (warp == GW_THREADS / 32 - 1) {  // where does the group of the block's last tuple end?
      const u32 jl = min(j0 + (u32)GW_THREADS, P.T) - 1;
      const u64 kl = P.keys[jl] & P.kmask;
      u32 elast = 0xffffffffu, lo = jl + 1;  // keys[jl .. lo-1] all equal kl
      for (int rounds = 0; elast == 0xffffffffu; rounds++) {
        if (lo >= P.T) { elast = P.T; break; }
        if (rounds == 4) { elast = key_bound<true>(P, kl, lo, P.T); break; }
        const u32 idx = lo + lane;
        const bool differs = idx >= P.T || (P.keys[idx] & P.kmask) != kl;
        const u32 bal = __ballot_sync(0xffffffffu, differs);
        if (bal) elast = lo + (__ffs(bal) - 1);
        else lo += 32;
      }
      if (lane == 0) s_last = elast;
    }
    __syncthreads();
    const u32 sf = s_first, sl = s_last;  // read before the next barrier: the next round overwrites them
    u32 before = 0, after = 0xffffffffu, cpre = (u32)__popc(cbal & lt_mask);
    for (int w = 0; w < warp; w++) { before = max(before, s_wmax[w]); cpre += s_wcnt[w]; }
    for (int w = warp + 1; w < GW_THREADS / 32; w++) after = min(after, s_wmin[w]);
    h = max(h, before);
    u32 e = __shfl_down_sync(0xffffffffu, nx, 1);
    if (lane == 31) e = 0xffffffffu;
    e = min(e, after);
    s_pre[threadIdx.x] = (unsigned short)cpre;  // NODEMSG tuples of the block before this one
    __syncthreads();
    if (!valid) continue;
    const bool inside = h != 0 && e != 0xffffffffu;  // the whole group lies in this block
    const u32 s = h ? h - 1 : sf;                    // else it began before the block ...
    if (e == 0xffffffffu) e = sl;                    // ... or runs past it
    P.seg_head[j] = s;
    if (!cand && !head) continue;
    const u32 g = e - s;
    const bool live = (long long)g > P.low_kmer;
    if (head) {
      // BuildHighKmerList.java:96-117: weighted count over windows [0, L-K) above UP_KMER
      if (kseg != 0 && (long long)g * (long long)P.max_cov > P.up_kmer) {
        ull sum = 0;
        for (u32 t = s; t < e; t++) {
          const u64 m2 = tuple_meta(P, t);
          if ((int)((m2 >> 1) & posmask) < L - K) sum += P.node_cov[(u32)(m2 >> (P.pb + 1))];
        }
        if ((long long)sum > P.up_kmer) atomicAdd(&P.dcount[DC_HKMER], 1ull);
      }
    }
    if (cand && live) {
      // MatchPrefix.java:366-368: the elist of a k-mer that has a node entry is cut to its first UP_KMER
      // entries by overlap; which entries survive depends on Hadoop's value order (refused by the host)
      if ((long long)g > P.up_kmer) atomicAdd(&P.dcount[DC_UPKMER_LISTS], 1ull);
      const u32 lid = 2 * (u32)(m >> (P.pb + 1)) + (pos == 0 ? 1u : 0u);  // list (b, flip y)
      const u32 v = lid | ((u32)(m & 1) << 31);
      if (inside) {  // position = number of NODEMSG tuples of the group before this one: no atomics
        P.ne_list[s + (cpre - (u32)s_pre[s - j0])] = v;
      } else {       // the group straddles blocks: claim the first free position of its range
        for (u32 t = s; t < e; t++)
          if (atomicCAS(&P.ne_list[t], 0xffffffffu, v) == 0xffffffffu) break;
      }
      P.slotcnt[lid] = g;
    }
  }
}

constexpr int JN_XW = 64;

// exclusive scan of one value per thread across a block; tmp holds blockDim.x / 32 + 1 values
__device__ __forceinline__ u32 block_excl_scan_bj(u32 v, u32* tmp, u32& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const u32 w = lane < nw ? tmp[lane] : 0;
    u32 wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < nw) tmp[lane] = wi - w;
    if (lane == nw - 1) tmp[nw] = wi;
  }
  __syncthreads();
  const u32 r = tmp[warp] + inc - v;
  total = tmp[nw];
  __syncthreads();
  return r;
}  // side-buffer records staged per warp (flushed once 32 are waiting)

template <int NW>
__device__ __forceinline__ void unpack_row(const uint4 (&q)[NW / 2], u64 (&r)[NW]) {
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    r[2 * i] = (u64)q[i].x | ((u64)q[i].y << 32);
    r[2 * i + 1] = (u64)q[i].z | ((u64)q[i].w << 32);
  }
}
template <int NW>
__device__ __forceinline__ void load_row(const u64* __restrict__ reads, u32 idx, uint4 (&q)[NW / 2]) {
  const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) q[i] = __ldg(row + i);
}

// The kernel is bound by load latency, not bandwidth (ncu: one dependent chain seg_head -> tuple
// -> read a -> ne_list -> read b -> off1 -> store, six DRAM round trips deep), so everything a
// thread will need is requested as early as its address is known: level 1 the tuple and its
// group head, level 2 the read, both neighbour tuples and the first two node entries of the
// group (most groups have at most two: one window-0 and one last-window tuple), level 3 the rows
// and slot offsets of those entries.  Side-buffer records are staged per WARP in shared memory
// and flushed warp-synchronously, so the tuple loop has no block barrier.
template <int NW, int MINB>
__global__ void __launch_bounds__(256, MINB) join_verify_kernel(JoinParams P) {
  __shared__ u64 s_x[8][JN_XW];
  const int L = P.L, K = P.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 posmask = (1u << P.pb) - 1;
  u32 n_verified = 0;
  u32 wn = 0;  // records staged by this warp (warp-uniform)
  const u32 stride = gridDim.x * blockDim.x;
  const u32 rounds = (P.T + stride - 1) / stride;
  for (u32 rnd = 0; rnd < rounds; rnd++) {
    const u32 j = rnd * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < P.T;
    u64 xrec[4];  // side-buffer records of this tuple in this round (beyond 4: straight to global)
    int nx = 0;
    if (active) {
      // ---- level 1 ----
      const u32 s = P.seg_head[j];
      const u64 kj = P.keys[j];
      const u32 vj = P.vals[j];
      const u64 m = ((kj >> (2 * K)) << 32) | (u64)vj;
      const u32 a = (u32)(m >> (P.pb + 1));
      // ---- level 2: everything addressed by (j, s, a) ----
      uint4 qa[NW / 2];
      load_row<NW>(P.node_seq, a, qa);
      const bool has_l = j > s, has_r = j + 1 < P.T;
      const u64 k_l = has_l ? P.keys[j - 1] : 0ull;
      const u32 v_l = has_l ? P.vals[j - 1] : 0u;
      const u64 k_r = has_r ? P.keys[j + 1] : 0ull;
      const u32 v_r = has_r ? P.vals[j + 1] : 0u;
      const u32 sh_r = has_r ? P.seg_head[j + 1] : 0xffffffffu;
      u32 ne[2];
      ne[0] = P.ne_list[s];
      const u32 sh_1 = s + 1 < P.T ? P.seg_head[s + 1] : 0xffffffffu;
      ne[1] = s + 1 < P.T ? P.ne_list[s + 1] : 0xffffffffu;
      if (ne[0] == 0xffffffffu || sh_1 != s) ne[1] = 0xffffffffu;
      // ---- level 3: rows and slot offsets of the first two node entries ----
      uint4 qb0[NW / 2], qb1[NW / 2];
      u32 offv0 = 0, offv1 = 0;
      if (ne[0] != 0xffffffffu) {
        load_row<NW>(P.node_seq, (ne[0] & 0x7fffffffu) >> 1, qb0);
        offv0 = P.off1[ne[0] & 0x7fffffffu];
      }
      if (ne[1] != 0xffffffffu) {
        load_row<NW>(P.node_seq, (ne[1] & 0x7fffffffu) >> 1, qb1);
        offv1 = P.off1[ne[1] & 0x7fffffffu];
      }

      const int pos = (int)((m >> 1) & posmask);
      const int d = (int)(m & 1);
      const u64 kseg = kj & P.kmask;
      u64 R[NW], Rc[NW];
      unpack_row<NW>(qa, R);
      revcomp_read<NW>(R, Rc, L);
      // Everything that depends on the tuple only, for both lists it feeds (sel 0: list of the
      // canonical key c, sel 1: list of rc(c)): orientation x, overlap ov, the oriented read
      // shifted so that its last ov bases lead, and whether the mirror candidate is proposed in
      // the group of X' = a^(flip x)[0:K] (SURVEY.md 8a': ov > K and X' not poly-A/T, or ov == K
      // and X' > rc(X')).
      int x0, ov0, x1, ov1;
      entry_of(L, K, pos, d, 0, x0, ov0);
      entry_of(L, K, pos, d, 1, x1, ov1);
      u64 A0[NW], A1[NW];
#pragma unroll
      for (int w = 0; w < NW; w++) { A0[w] = x0 ? Rc[w] : R[w]; A1[w] = x1 ? Rc[w] : R[w]; }
      shl_bases<NW>(A0, L - ov0);
      shl_bases<NW>(A1, L - ov1);
      const u64 ksuf_f = extract_kmer<NW>(R, L - K, K), ksuf_r = extract_kmer<NW>(Rc, L - K, K);
      const u64 xp_f = revcomp_kmer(ksuf_f, K), xp_r = revcomp_kmer(ksuf_r, K);
      const bool pf_gt = (xp_f < ksuf_f ? xp_f : ksuf_f) != 0, pf_eq = xp_f > ksuf_f;
      const bool pr_gt = (xp_r < ksuf_r ? xp_r : ksuf_r) != 0, pr_eq = xp_r > ksuf_r;
      const bool prop0 = x0 ? (ov0 > K ? pr_gt : pr_eq) : (ov0 > K ? pf_gt : pf_eq);
      const bool prop1 = x1 ? (ov1 > K ? pr_gt : pr_eq) : (ov1 > K ? pf_gt : pf_eq);
      // does a neighbouring tuple of the group belong to the same read?  (tuples of one read are
      // contiguous: the sort is stable and keygen emits in (node, window) order)
      const u64 m_l = ((k_l >> (2 * K)) << 32) | (u64)v_l, m_r = ((k_r >> (2 * K)) << 32) | (u64)v_r;
      const bool same_l = has_l && (u32)(m_l >> (P.pb + 1)) == a;
      const bool same_r = has_r && sh_r == s && (u32)(m_r >> (P.pb + 1)) == a;
      const bool lone = !same_l && !same_r && kseg != 0 && P.low_kmer == 1;
      // one node entry of the group against this tuple (inlined for the two prefetched entries
      // and for the rare further ones)
      auto process = [&](const u32 v, const u32 off_l, u64 (&B)[NW]) {
        const u32 lid = v & 0x7fffffffu;
        const u32 b = lid >> 1;
        const int y = (int)((lid & 1u) ^ 1u);          // list (b, flip y) belongs to node entry (b, y)
        const int dn = (int)(v >> 31);
        const int sel = y == 0 ? dn : (dn ^ 1);        // MatchPrefix.java:275-296 (y == 0: window-0 entry)
        const int x = sel ? x1 : x0, ov = sel ? ov1 : ov0;
        if (y) {
          u64 r[NW];
#pragma unroll
          for (int w = 0; w < NW; w++) r[w] = B[w];
          revcomp_read<NW>(r, B, L);
        }
        u64 A[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) A[w] = sel ? A1[w] : A0[w];
        // VerifyOverlap.java:295-309: a^x[L-ov:] == b^y[0:ov]; self candidates skipped at
        // MatchPrefix.java:405-407
        bool ok = a != b && prefix_equal<NW>(A, B, ov);
        if (ok && (same_l || same_r)) {
          // maximal-overlap filter, VerifyOverlap.java:275-283: a larger verified overlap of the
          // same (a, x) with the same (b, y) wins; an identical entry of an earlier tuple
          // (MatchPrefix.java:409 hasEdge) is emitted only once
          for (int dirn = -1; dirn <= 1 && ok; dirn += 2) {
            long long jj = (long long)j + dirn;
            while (jj >= (long long)s && jj < (long long)P.T && P.seg_head[jj] == s) {
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
        if (ok) {
          const size_t slot = (size_t)off_l + (j - s);
          P.slots[slot] = P.lay.encode(b, (u32)(y ^ 1), (u32)ov, (u32)(x ^ 1), a);  // GenReverseEdge
          {  // the record's overhang as seen from b: a^(flip x)[ov:]
            u64 H[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) H[w] = x ? R[w] : Rc[w];
            shl_bases<NW>(H, ov);
            store_overhang<NW>(P.ovh + slot * (size_t)P.nwo, H, P.nwo);
          }
          n_verified++;
          // Is the mirror candidate (b, flip y, ov) -> (a, flip x) certain to be proposed and
          // kept at a's own node entry?  Then that entry writes this edge into list (a, x) and it
          // need not travel.  (kept iff no larger TRUE overlap, which would show as another tuple
          // of a in this group.)
          const bool certain = lone && (sel ? prop1 : prop0);
          if (!certain) {
            const u64 rec = P.lay.encode(a, (u32)x, (u32)ov, (u32)y, b);
            if (nx < 4) {
#pragma unroll
              for (int q = 0; q < 4; q++)
                if (q == nx) xrec[q] = rec;
              nx++;
            } else {  // more than four per tuple: straight to the side buffer
              ull gp = atomicAdd(&P.dcount[DC_EDGES_RAW], 1ull);
              if (gp < P.xcap) P.xbuf[gp] = rec;
            }
          }
        }
      };
      if (ne[0] != 0xffffffffu) {
        u64 B[NW];
        unpack_row<NW>(qb0, B);
        process(ne[0], offv0, B);
        if (ne[1] != 0xffffffffu) {
          unpack_row<NW>(qb1, B);
          process(ne[1], offv1, B);
          for (u32 t = 2;; t++) {
            const u32 idx = s + t;
            if (idx >= P.T || P.seg_head[idx] != s) break;
            const u32 v = P.ne_list[idx];
            if (v == 0xffffffffu) break;
            load_read_row<NW>(P.node_seq, (v & 0x7fffffffu) >> 1, B);
            process(v, P.off1[v & 0x7fffffffu], B);
          }
        }
      }
    }
    // ---- stage this round's side-buffer records per warp; flush once 32 are waiting ----
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const bool have = q < nx;
      const u32 bal = __ballot_sync(0xffffffffu, have);
      if (bal == 0) break;  // warp-uniform
      if (have) s_x[warp][wn + __popc(bal & lt_mask)] = xrec[q];
      wn += __popc(bal);
      __syncwarp();
      if (wn >= 32) {
        ull base = 0;
        if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_RAW], (ull)wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (u32 t = lane; t < wn; t += 32)
          if (base + t < P.xcap) P.xbuf[base + t] = s_x[warp][t];
        wn = 0;
        __syncwarp();
      }
    }
  }
  if (wn) {
    ull base = 0;
    if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_RAW], (ull)wn);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (u32 t = lane; t < wn; t += 32)
      if (base + t < P.xcap) P.xbuf[base + t] = s_x[warp][t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_verified += __shfl_xor_sync(0xffffffffu, n_verified, o);
  if (lane == 0 && n_verified) {
    atomicAdd(&P.dcount[DC_CAND_VERIFIED], (ull)n_verified);
    atomicAdd(&P.dcount[DC_EDGES_M], (ull)n_verified);
  }
}

// ---------------------------------------------------------------------------------------------
// Bucket path: grouping + join + verify + mirror in ONE kernel, one CTA per hash bucket.
//
// bucket_partition (primitives.cu) leaves the tuples of every k-mer in one bucket of a few hundred
// tuples, in no particular order.  The CTA loads its bucket once, groups it by k-mer in shared memory
// (counting sort on 10 more hash bits; two k-mers that share those bits are separated afterwards, by
// one thread per such bin), finds the node entries of every group, reserves the slot ranges of the
// lists they produce (one global atomic per CTA) and runs the join on the groups where they lie.
// Against the flat pipeline (LSD sort + group_walk_kernel + join_verify_kernel, kept below as the path
// for buckets that do not fit) the sorted tuple array, seg_head, ne_list, the list-size scan and the
// pre-clearing of the slot array never touch HBM: every slot of a list is written exactly once, with
// its record or with ~0.
//
// Nothing here depends on the order of the tuples inside a group: "another tuple of the same read in
// this group" (maximal-overlap filter, `certain`) is found by scanning the group in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int BJ_LBINS = 1 << kLocalBits;      // 1024 local bins
// per position: key 8, payload 4, group 4, slot offset 4, node-entry list 2 bytes
constexpr size_t bj_smem_bytes(int cap) {
  return (size_t)cap * (8 + 4 + 4 + 4 + 2) + (BJ_LBINS + 1) * 4 + BJ_LBINS * 4 + (BJ_LBINS / 32) * 4 + 64;
}

// M32: node | window | dir of every tuple fit the 32-bit payload (no payload bits above the k-mer in
// the key), which is the case up to 2^26 nodes at W <= 32 and 2^24 nodes at W <= 128.
//
// All per-position loops are warp-granular (warp w takes the 32-position chunks w, w + 8, ...), so a
// bucket of n tuples costs ceil(n / 32) warp iterations whatever n is.  (Measured dead end: fetching the
// packed reads of the bucket's tuples into shared memory with cp.async while the grouping runs -- the join
// then reads nothing from global memory -- was 5 % slower: the gathers hit the L2, which holds node_seq
// in its persisting window, and the 16 KB of rows cost occupancy.)
// CAP: tuples per bucket the kernel holds.  1024 for reads of up to 64 bases (22.5 KB + 8 KB of shared memory:
// seven CTAs per SM), 2048 for longer reads, whose kernel is limited to four CTAs per SM by its registers.
template <int NW, int TPB, int MINB, bool M32, int CAP>
__global__ void __launch_bounds__(TPB, MINB) bucket_join_kernel(JoinParams P) {
  constexpr int BJ_CAP = CAP;
  static_assert(BJ_CAP / 2 <= BJ_LBINS, "node-entry counters are packed two per word of scnt");
  static_assert(BJ_CAP <= (1 << 11), "list index of a node entry is packed into 11 bits");
  constexpr int BJ_THREADS = TPB, BJ_WARPS = TPB / 32, BJ_CHUNKS = BJ_CAP / 32 / BJ_WARPS;
  extern __shared__ __align__(16) unsigned char bj_smem[];
  u64* skey = reinterpret_cast<u64*>(bj_smem);           // [CAP] tuples grouped by k-mer
  u32* sval = reinterpret_cast<u32*>(skey + BJ_CAP);     // [CAP]
  u32* sgrp = sval + BJ_CAP;                             // [CAP] head | end << 16 of the group of position p
  u32* soff = sgrp + BJ_CAP;                             // [CAP] node entry i: slot offset in the CTA's range | list index << 21
  u32* sstart = soff + BJ_CAP;                           // [LBINS + 1] bin counts, then bin offsets
  u32* scnt = sstart + BJ_LBINS + 1;                     // [LBINS] node entries per group head, two 16-bit counters per word
  u32* smixed = scnt + BJ_LBINS;                         // [LBINS / 32] bins that hold more than one k-mer
  unsigned short* sne = reinterpret_cast<unsigned short*>(smixed + BJ_LBINS / 32);  // [CAP] positions head.. : node entries of the group
  __shared__ u32 scan_tmp[BJ_THREADS / 32 + 1];
  __shared__ u32 s_any_mixed, s_slots, s_lists;
  __shared__ ull s_base, s_lbase;

  const int L = P.L, K = P.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 posmask = (1u << P.pb) - 1;
  const int nsh = P.pb + 1;  // node = meta >> nsh
  const u32 lo = P.bstart[blockIdx.x];
  const u32 n = P.bstart[blockIdx.x + 1] - lo;
  if (n == 0) return;
  if (n > (u32)BJ_CAP) {  // the host falls back to the flat pipeline
    if (tid == 0) atomicAdd(&P.dcount[DC_BUCKET_OVER], 1ull);
    return;
  }
  const u32 n_chunks = (n + 31) >> 5;
  auto ldigit = [&](u64 k) { return (u32)(kmer_hash(k & P.kmask) >> P.lshift) & (BJ_LBINS - 1); };
  auto meta_at = [&](u32 p) -> u64 { return M32 ? (u64)sval[p] : (((skey[p] >> (2 * K)) << 32) | (u64)sval[p]); };
  auto node_at = [&](u32 p) -> u32 { return M32 ? (sval[p] >> nsh) : (u32)(meta_at(p) >> nsh); };
  auto necnt = [&](u32 head) { return (scnt[head >> 1] >> ((head & 1u) * 16)) & 0xffffu; };

  // ---- group the bucket by k-mer ----
  for (int t = tid; t < BJ_LBINS; t += BJ_THREADS) { sstart[t] = 0; scnt[t] = 0; }
  if (tid < BJ_LBINS / 32) smixed[tid] = 0;
  if (tid == 0) { s_any_mixed = 0; s_slots = 0; s_lists = 0; }
  __syncthreads();
  // one shared-memory atomic per tuple: its return value is the tuple's rank inside its bin
  // up to 8 chunks per warp the tuples stay in registers between the two passes; beyond that they are read
  // again (from L1 / L2) when they are placed
  constexpr bool KEEP = BJ_CHUNKS <= 8;
  u64 kreg[KEEP ? BJ_CHUNKS : 1];
  u32 vreg[KEEP ? BJ_CHUNKS : 1], dr[BJ_CHUNKS];
#pragma unroll
  for (int i = 0; i < BJ_CHUNKS; i++) {
    const u32 e = ((u32)(warp + i * BJ_WARPS) << 5) + lane;
    dr[i] = 0xffffffffu;
    if (e < n) {
      const u64 k = P.keys[lo + e];
      if (KEEP) {
        kreg[i] = k;
        vreg[i] = P.vals[lo + e];  // requested together with the key: one DRAM round trip per CTA, not two
      }
      const u32 d = ldigit(k);
      dr[i] = (d << 16) | atomicAdd(&sstart[d], 1u);
    }
  }
  __syncthreads();
  {
    u32 c[BJ_LBINS / BJ_THREADS], run = 0;
#pragma unroll
    for (int q = 0; q < BJ_LBINS / BJ_THREADS; q++) { c[q] = sstart[tid * (BJ_LBINS / BJ_THREADS) + q]; run += c[q]; }
    u32 total;
    u32 ex = block_excl_scan_bj(run, scan_tmp, total);
#pragma unroll
    for (int q = 0; q < BJ_LBINS / BJ_THREADS; q++) { sstart[tid * (BJ_LBINS / BJ_THREADS) + q] = ex; ex += c[q]; }
    if (tid == 0) sstart[BJ_LBINS] = n;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < BJ_CHUNKS; i++) {
    if (dr[i] != 0xffffffffu) {
      const u32 pos = sstart[dr[i] >> 16] + (dr[i] & 0xffffu);
      if (KEEP) {
        skey[pos] = kreg[i];
        sval[pos] = vreg[i];
      } else {
        const u32 e = ((u32)(warp + i * BJ_WARPS) << 5) + lane;
        skey[pos] = P.keys[lo + e];
        sval[pos] = P.vals[lo + e];
      }
    }
  }
  __syncthreads();
  // bins that hold more than one k-mer (two of the k-mers of a bucket share the 10 bits): rare enough
  // that one thread per such bin gathers equal k-mers together
  for (u32 p = tid; p < n; p += BJ_THREADS) {
    const u32 d = ldigit(skey[p]);
    if ((skey[p] ^ skey[sstart[d]]) & P.kmask) {
      atomicOr(&smixed[d >> 5], 1u << (d & 31));
      s_any_mixed = 1;
    }
  }
  __syncthreads();
  if (s_any_mixed) {  // block-uniform
    for (u32 d = tid; d < (u32)BJ_LBINS; d += BJ_THREADS) {
      if (!((smixed[d >> 5] >> (d & 31)) & 1u)) continue;
      u32 i = sstart[d];
      const u32 e = sstart[d + 1];
      while (i < e) {
        const u64 pivot = skey[i] & P.kmask;
        u32 j = i + 1;
        for (u32 t = i + 1; t < e; t++) {
          if ((skey[t] & P.kmask) == pivot) {
            if (t != j) {
              const u64 tk = skey[t]; skey[t] = skey[j]; skey[j] = tk;
              const u32 tv = sval[t]; sval[t] = sval[j]; sval[j] = tv;
            }
            j++;
          }
        }
        i = j;
      }
    }
    __syncthreads();
  }
  // ---- [head, end) of every position's group; node entries (MatchPrefix.java:366: groups above LOW_KMER) ----
  for (u32 ch = warp; ch < n_chunks; ch += BJ_WARPS) {  // warp-uniform trip count: the reservation below is warp-collective
    const u32 p = (ch << 5) + lane;
    const bool valid = p < n;
    bool entry = false;
    u32 lid = 0, len = 0, head = 0;
    if (valid) {
      const u64 kp = skey[p];
      const u32 d = ldigit(kp);
      u32 end = sstart[d + 1];
      head = sstart[d];
      if ((smixed[d >> 5] >> (d & 31)) & 1u) {
        const u64 ck = kp & P.kmask;
        u32 h = p, e = p + 1;
        while (h > head && (skey[h - 1] & P.kmask) == ck) h--;
        while (e < end && (skey[e] & P.kmask) == ck) e++;
        head = h;
        end = e;
      }
      sgrp[p] = head | (end << 16);
      const u64 m = meta_at(p);
      const int pos = (int)((m >> 1) & posmask);
      const u32 g = end - head;
      if (p == head) {
        // BuildHighKmerList.java:96-117: weighted count over windows [0, L-K) above UP_KMER
        if ((kp & P.kmask) != 0 && (long long)g * (long long)P.max_cov > P.up_kmer) {
          ull sum = 0;
          for (u32 t = head; t < end; t++) {
            const u64 m2 = meta_at(t);
            if ((int)((m2 >> 1) & posmask) < L - K) sum += P.node_cov[(u32)(m2 >> nsh)];
          }
          if ((long long)sum > P.up_kmer) atomicAdd(&P.dcount[DC_HKMER], 1ull);
        }
      }
      entry = (pos == 0 || pos == L - K) && (long long)g > P.low_kmer;  // a live NODEMSG tuple
      if (entry) {
        if ((long long)g > P.up_kmer) atomicAdd(&P.dcount[DC_UPKMER_LISTS], 1ull);  // MatchPrefix.java:366-368 (refused by the host)
        lid = 2 * (u32)(m >> nsh) + (pos == 0 ? 1u : 0u);  // list (b, flip y)
        len = g + (P.extra ? P.extra[lid] : 0u);
      }
    }
    // slots and list indices of the CTA's ranges: one shared atomic pair per warp (a same-address atomic per
    // node entry was 23 % of the kernel's stall samples), lanes take consecutive pieces
    const u32 ebal = __ballot_sync(0xffffffffu, entry);
    if (ebal) {
      u32 incl = len;  // inclusive prefix sum of len over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const u32 wtotal = __shfl_sync(0xffffffffu, incl, 31);
      u32 wbase = 0, lbase = 0;
      if (lane == 31) { wbase = atomicAdd(&s_slots, wtotal); lbase = atomicAdd(&s_lists, (u32)__popc(ebal)); }
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      lbase = __shfl_sync(0xffffffffu, lbase, 31);
      if (entry) {
        const u32 old = atomicAdd(&scnt[head >> 1], (head & 1u) ? 0x10000u : 1u);
        const u32 kk = (old >> ((head & 1u) * 16)) & 0xffffu;
        sne[head + kk] = (unsigned short)p;
        soff[head + kk] = (wbase + incl - len) | ((lbase + (u32)__popc(ebal & lt_mask)) << 21);
      }
    }
  }
  __syncthreads();
  const u32 n_slots = s_slots, n_lists = s_lists;
  if (n_lists == 0) return;       // no node entry in this bucket: nothing to join
  if (n_slots >= (1u << 21)) {    // packed offsets overflowed (thousands of node entries in one group): flat pipeline
    if (tid == 0) atomicAdd(&P.dcount[DC_BUCKET_OVER], 1ull);
    return;
  }
  if (tid == 0) {
    s_base = atomicAdd(&P.slot_cursor[0], (ull)n_slots);
    s_lbase = atomicAdd(&P.slot_cursor[1], (ull)n_lists);
  }
  __syncthreads();
  const ull base = s_base;
  if (base + n_slots > P.slot_cap) return;  // the host sees slot_cursor > slot_cap and retries with the exact size
  for (u32 i = tid; i < n; i += BJ_THREADS) {  // publish the lists; the tail beyond the group's own slots is empty
    const u32 head = sgrp[i] & 0xffffu, end = sgrp[i] >> 16;
    if (i - head < necnt(head)) {
      const u64 m = meta_at(sne[i]);
      const u32 lid = 2 * (u32)(m >> nsh) + ((int)((m >> 1) & posmask) == 0 ? 1u : 0u);
      const u32 g = end - head, len = g + (P.extra ? P.extra[lid] : 0u);
      const u32 off = (u32)(base + (soff[i] & 0x1fffffu));
      P.adj[lid] = make_uint2(off, len);
      P.list_order[s_lbase + (soff[i] >> 21)] = make_uint4(lid, off, len, 0u);
      for (u32 t = g; t < len; t++) P.slots[(size_t)off + t] = ~0ull;
    }
  }

  // ---- join + verify + mirror: every (node entry, tuple) pair of a group, one thread per tuple ----
  const int sh_node = P.lay.nb + P.lay.ob + 2, sh_x = P.lay.nb + P.lay.ob + 1, sh_ov = P.lay.nb + 1, sh_y = P.lay.nb;
  u32 n_verified = 0;
  for (u32 ch = warp; ch < n_chunks; ch += BJ_WARPS) {  // warp-uniform trip count: the MATCH.ANY below is warp-collective
    const u32 p = (ch << 5) + lane;
    const bool valid = p < n;
    const u32 head = valid ? (sgrp[p] & 0xffffu) : 0u, end = valid ? (sgrp[p] >> 16) : 0u;
    const u64 m = valid ? meta_at(p) : 0ull;
    const u32 a = (u32)(m >> nsh);
    // Another tuple of the same read in this group?  The lanes of a warp hold consecutive positions, so the
    // part of the group inside the warp's window is one MATCH.ANY; only the part outside is scanned.
    const u32 peers = __match_any_sync(0xffffffffu, valid ? (((u64)head << 32) | (u64)a) : (0xffffffff00000000ull | (u64)lane));
    const u32 nne = valid ? necnt(head) : 0u;
    if (nne == 0) continue;
    uint4 qa[NW / 2];
    load_row<NW>(P.node_seq, a, qa);
    bool multi = (peers & (peers - 1)) != 0;
    {
      const u32 w0 = ch << 5, w1 = w0 + 32;
      for (u32 t = head; t < end && t < w0; t++) multi = multi || node_at(t) == a;
      for (u32 t = w1 > head ? w1 : head; t < end; t++) multi = multi || node_at(t) == a;
    }
    const int pos = (int)((m >> 1) & posmask);
    const int d = (int)(m & 1);
    u64 R[NW], Rc[NW];
    unpack_row<NW>(qa, R);
    revcomp_read<NW>(R, Rc, L);
    // sel 0: candidate in the list of the canonical k-mer c, sel 1: of rc(c) (entry_of); the oriented read
    // shifted so that its last ov bases lead
    int x0, ov0, x1, ov1;
    entry_of(L, K, pos, d, 0, x0, ov0);
    entry_of(L, K, pos, d, 1, x1, ov1);
    u64 A0[NW], A1[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) { A0[w] = d ? Rc[w] : R[w]; A1[w] = d ? R[w] : Rc[w]; }  // x0 = d, x1 = d ^ 1
    shl_bases<NW>(A0, L - ov0);
    shl_bases<NW>(A1, L - ov1);
    // is the mirror candidate proposed in the group of X' = a^(flip x)[0:K]?  (SURVEY.md 8a': ov > K and X' not
    // poly-A/T, or ov == K and X' > rc(X')).  X' of x = f is rc(last k-mer of a), of x = r the first k-mer of a.
    const u64 kfirst = R[0] >> (64 - 2 * K), klast = extract_kmer<NW>(R, L - K, K);
    const u64 rc_first = revcomp_kmer(kfirst, K), rc_last = revcomp_kmer(klast, K);
    const bool pf_gt = (rc_last < klast ? rc_last : klast) != 0, pf_eq = rc_last > klast;
    const bool pr_gt = (kfirst < rc_first ? kfirst : rc_first) != 0, pr_eq = kfirst > rc_first;
    const bool prop0 = d ? (ov0 > K ? pr_gt : pr_eq) : (ov0 > K ? pf_gt : pf_eq);
    const bool prop1 = d ? (ov1 > K ? pf_gt : pf_eq) : (ov1 > K ? pr_gt : pr_eq);
    const bool lone = !multi && (skey[p] & P.kmask) != 0 && P.low_kmer == 1;
    // the tuple's half of the mirror record (b, flip y, ov, flip x, a) for both lists
    const u64 apart0 = ((u64)(u32)(L - ov0) << sh_ov) | ((u64)(u32)(x0 ^ 1) << sh_y) | (u64)a;
    const u64 apart1 = ((u64)(u32)(L - ov1) << sh_ov) | ((u64)(u32)(x1 ^ 1) << sh_y) | (u64)a;
    for (u32 q = 0; q < nne; q++) {  // one node entry (b, y) against this tuple
      const u32 pe = sne[head + q];
      const u64 me = meta_at(pe);
      const u32 off_l = (u32)(base + (soff[head + q] & 0x1fffffu));
      const u32 b = (u32)(me >> nsh);
      const int y = ((int)((me >> 1) & posmask) == 0) ? 0 : 1;  // window-0 entry: y = f
      const int sel = y ^ (int)(me & 1);                          // MatchPrefix.java:275-296
      const int x = sel ? x1 : x0, ov = sel ? ov1 : ov0;
      u64 B[NW];
      load_read_row<NW>(P.node_seq, b, B);  // L1: every tuple of the group reads the same row
      if (y) {
        u64 r[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) r[w] = B[w];
        revcomp_read<NW>(r, B, L);
      }
      u64 A[NW];
#pragma unroll
      for (int w = 0; w < NW; w++) A[w] = sel ? A1[w] : A0[w];
      // VerifyOverlap.java:295-309; self candidates skipped at MatchPrefix.java:405-407
      bool ok = a != b && prefix_equal<NW>(A, B, ov);
      if (ok && multi) {
        // maximal-overlap filter (VerifyOverlap.java:275-283) against the other tuples of read a in the
        // group; an identical entry (MatchPrefix.java:409 hasEdge) is emitted by the first of them only
        for (u32 t = head; t < end && ok; t++) {
          if (t == p || node_at(t) != a) continue;
          const u64 m2 = meta_at(t);
          int x2, ov2;
          entry_of(L, K, (int)((m2 >> 1) & posmask), (int)(m2 & 1), sel, x2, ov2);
          if (x2 == x && ov2 == ov && t < p) ok = false;
          else if (x2 == x && ov2 > ov) {
            u64 A2[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) A2[w] = x ? Rc[w] : R[w];
            shl_bases<NW>(A2, L - ov2);
            if (prefix_equal<NW>(A2, B, ov2)) ok = false;
          }
        }
      }
      const size_t slot = (size_t)off_l + (p - head);
      if (ok) {
        P.slots[slot] = ((u64)b << sh_node) | ((u64)(u32)(y ^ 1) << sh_x) | (sel ? apart1 : apart0);  // GenReverseEdge
        u64 H[NW];  // the record's overhang as seen from b: a^(flip x)[ov:]
#pragma unroll
        for (int w = 0; w < NW; w++) H[w] = x ? R[w] : Rc[w];
        shl_bases<NW>(H, ov);
        store_overhang<NW>(P.ovh + slot * (size_t)P.nwo, H, P.nwo);
        n_verified++;
        // Is the mirror candidate certain to be proposed and kept at a's own node entry?  Then that entry
        // writes this edge into list (a, x) itself (see join_verify_kernel); the others (~3 %) go through
        // the side buffer
        if (!(lone && (sel ? prop1 : prop0))) {
          // warp-aggregated append (one same-address global atomic per side record was 14 % of the stall samples)
          const u32 act = __activemask();
          const int leader = __ffs(act) - 1;
          ull gp = 0;
          if (lane == leader) gp = atomicAdd(&P.dcount[DC_EDGES_RAW], (ull)__popc(act));
          gp = __shfl_sync(act, gp, leader) + (ull)__popc(act & lt_mask);
          if (gp < P.xcap) P.xbuf[gp] = P.lay.encode(a, (u32)x, (u32)ov, (u32)y, b);
        }
      } else {
        P.slots[slot] = ~0ull;  // every slot is written exactly once: the slot array is never pre-cleared
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_verified += __shfl_xor_sync(0xffffffffu, n_verified, o);
  if (lane == 0 && n_verified) {
    atomicAdd(&P.dcount[DC_CAND_VERIFIED], (ull)n_verified);
    atomicAdd(&P.dcount[DC_EDGES_M], (ull)n_verified);
  }
}

// ---------------------------------------------------------------------------------------------
// Side-buffer records: edge (a,x,ov,y,b) goes into a free slot of list (a,x) unless a's own node
// entry already wrote it.  One thread per record.  A record that finds no free slot (only
// possible around poly-A/T k-mers, whose interior windows have no tuple and therefore no slot)
// is counted in extra[list]; the host then re-runs the join with enlarged slot ranges.
// ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(256) side_insert_kernel(const u64* __restrict__ xs, size_t n, EdgeLayout lay,
                                                          const uint2* __restrict__ adj, u64* __restrict__ slots,
                                                          u32* __restrict__ extra, ull* __restrict__ dcount,
                                                          const u64* __restrict__ node_seq, u32* __restrict__ ovh,
                                                          int nwo, int L) {
  u32 inserted = 0, overflow = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 key = xs[i];
    const u32 lid = lay.list(key);
    const uint2 lr = adj[lid];
    const u32 lo = lr.x, hi = lr.x + lr.y;
    // the record's overhang b^y[ov:] (requested before the list is scanned)
    u64 H[NW];
    load_read_row<NW>(node_seq, lay.nbr(key), H);
    // one pass over the list's slots, eight loads in flight: duplicate? first free slot?
    bool dup = false;
    u32 hole = 0xffffffffu;
    for (u32 t = lo; t < hi && !dup; t += 8) {
      u64 k[8];
#pragma unroll
      for (int i = 0; i < 8; i++) k[i] = (t + i < hi) ? slots[t + i] : 0ull;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (k[i] == key) dup = true;
        if (k[i] == ~0ull && hole == 0xffffffffu && t + i < hi) hole = t + i;
      }
    }
    if (dup) continue;
    bool placed = false;
    u32 at = 0;
    for (u32 t = hole; t < hi && !placed; t++) {  // another record may claim the slot first: move on
      if (slots[t] == ~0ull) {
        ull old = atomicCAS((ull*)&slots[t], ~0ull, (ull)key);
        placed = old == ~0ull;
        at = t;
      }
    }
    if (placed) {
      inserted++;
      if (lay.y(key)) {
        u64 r[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) r[w] = H[w];
        revcomp_read<NW>(r, H, L);
      }
      shl_bases<NW>(H, (int)lay.ov(key));
      store_overhang<NW>(ovh + (size_t)at * nwo, H, nwo);
    } else { atomicAdd(&extra[lid], 1u); overflow++; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    inserted += __shfl_xor_sync(0xffffffffu, inserted, o);
    overflow += __shfl_xor_sync(0xffffffffu, overflow, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (inserted) atomicAdd(&dcount[DC_EDGES_A], (ull)inserted);
    if (overflow) atomicAdd(&dcount[DC_MAX_LIST], (ull)overflow);  // reused as the overflow count of this attempt
  }
}

// owner rank of adjacency list (a, x): the rank of the k-mer group of node entry (a, flip x),
// i.e. of canon(last k-mer of a) for x = f and canon(first k-mer of a) for x = r
template <int NW>
__device__ __forceinline__ u32 list_owner(const u64* __restrict__ node_seq, u32 lid, int L, int K, int pbit,
                                          u32 wmask) {
  u64 r[NW];
  load_read_row<NW>(node_seq, lid >> 1, r);
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

static u32 read_u32(cb_ctx* c, const u32* p) {
  u32 v = 0;
  c->lc.rb.get(&v, p, 4, c->stream);
  c->lc.rb.sync(c->stream);
  return v;
}

// lists that receive side-buffer edges but have no node entry of their own (the entry's k-mer group is
// below LOW_KMER, or its k-mer emits no tuple): their slots are the extra ones alone
__global__ void __launch_bounds__(256) orphan_lists_kernel(const u32* __restrict__ extra, u32 n_lists, uint2* __restrict__ adj,
                                                           u64* __restrict__ slots, ull* __restrict__ cursor, ull slot_cap,
                                                           uint4* __restrict__ list_order) {
  const u32 lid = blockIdx.x * blockDim.x + threadIdx.x;
  if (lid >= n_lists) return;
  const u32 x = extra[lid];
  if (x == 0 || adj[lid].y != 0) return;
  const ull off = atomicAdd(&cursor[0], (ull)x);
  const ull li = atomicAdd(&cursor[1], 1ull);
  if (off + x > slot_cap) return;  // retried by the host with the exact size
  adj[lid] = make_uint2((u32)off, x);
  list_order[li] = make_uint4(lid, (u32)off, x, 0u);
  for (u32 t = 0; t < x; t++) slots[off + t] = ~0ull;
}

// 128 threads per CTA: the kernel is bound by the latency chain of a CTA's phases (loads -> grouping ->
// slot reservation -> join), so what counts is the number of CTAs in flight: 7 x 128 threads per SM ran
// the join in 4.6 ms where 4 x 256 threads took 6.1 ms (config 2)
template <int NW, int MINB, int CAP>
static void launch_bucket_join(const JoinParams& P, u32 n_buckets, bool m32, cudaStream_t s) {
  constexpr size_t smem = bj_smem_bytes(CAP);
  if (m32) {
    CB_CUDA(cudaFuncSetAttribute(bucket_join_kernel<NW, 128, MINB, true, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bucket_join_kernel<NW, 128, MINB, true, CAP><<<n_buckets, 128, smem, s>>>(P);
  } else {
    CB_CUDA(cudaFuncSetAttribute(bucket_join_kernel<NW, 128, MINB, false, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bucket_join_kernel<NW, 128, MINB, false, CAP><<<n_buckets, 128, smem, s>>>(P);
  }
}

__global__ void __launch_bounds__(256) make_adj_kernel(const u32* __restrict__ off, const u32* __restrict__ cnt, u32 n,
                                                       uint2* __restrict__ adj) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) adj[i] = make_uint2(off[i], cnt[i]);
}

// CB_SORT=classic: the flat pipeline (LSD sort + group_walk + join_verify) for every input
static bool use_bucket_path() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CB_SORT");
    v = (e && std::string(e) == "classic") ? 0 : 1;
  }
  return v == 1;
}

// node_seq is gathered by every tuple of the join (16..64-byte rows at random).  Experiment (CB_L2_WINDOW=1):
// pin it in the L2 with a persisting access window while the join runs.  It cuts the kernel's DRAM reads
// (ncu: 7.9 -> 1.5 GB on config 2) but not its time (4.57 ms either way: the kernel is bound by instruction
// issue and latency chains, not by bandwidth), the set-aside must be given back afterwards (left in place it
// starved the scatter passes: 0.62 -> 0.85 ms) and cudaDeviceSetLimit synchronises the whole device, which
// serialises the join behind the copy stream's prefetch of the next input (e2e 24.0 vs 15.9 ms).  Off by default.
static void set_l2_window(cb_ctx* c, const void* p, size_t bytes) {
  int dev = c->cfg.device, max_win = 0, max_persist = 0;
  cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  if (max_win <= 0 || max_persist <= 0) return;
  static const bool on = getenv("CB_L2_WINDOW") != nullptr;
  if (!on) return;
  // only when the whole table fits the set-aside part: a partial window (hitRatio < 1, the rest streaming)
  // made the gathers of a 138 MB table slower than plain LRU (2 GPUs)
  if (bytes > (size_t)max_persist || bytes > (size_t)max_win) return;
  const size_t carve = bytes;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
  cudaStreamAttrValue a;
  memset(&a, 0, sizeof(a));
  a.accessPolicyWindow.base_ptr = const_cast<void*>(p);
  a.accessPolicyWindow.num_bytes = bytes < (size_t)max_win ? bytes : (size_t)max_win;
  a.accessPolicyWindow.hitRatio = bytes <= carve ? 1.0f : (float)((double)carve / (double)bytes);
  a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
}
static void clear_l2_window(cb_ctx* c) {
  static const bool on = getenv("CB_L2_WINDOW") != nullptr;
  if (!on) return;
  cudaStreamAttrValue a;
  memset(&a, 0, sizeof(a));
  a.accessPolicyWindow.num_bytes = 0;
  a.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
  a.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
  cudaCtxResetPersistingL2Cache();
  // give the set-aside part of the L2 back: left in place it starves every other kernel of the path
  // (2 GPUs: dedupe_runs 1.03 -> 1.48 ms, consistency 0.89 -> 1.34 ms with 79 MB still set aside)
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0) != cudaSuccess) cudaGetLastError();
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
  c->carried_code = 0;
  // (on several GPUs a rank-local failure up to the tuple exchange travels with the exchange's counts)
  if (Tmax >= (1ull << 32)) {
    defer_error(c, CB_E_CAPACITY, "more than 2^32 k-mer tuples per context");
    Tmax = 0;
  }
  c->lay.nb = nbits;
  c->lay.ob = bits_for((u64)(L - K + 1));
  c->lay.L = L;
  const u32 n_lists = 2 * nn;
  const u64 kmask = (2 * K == 64) ? ~0ull : ((1ull << (2 * K)) - 1);
  bool bucket = use_bucket_path() && !(c->cfg.flags & CB_FLAG_FLAT_PIPELINE);
  // single GPU: the tuples go straight from keygen into the hash partition, whose level-1 digit counts
  // keygen takes while it writes
  BucketPlan plan = make_bucket_plan(Tmax);
  if (!distributed(c) && !plan.fits) bucket = false;  // more tuples than two scatter passes cut small enough
  DevBuf<u32> cnt1;
  const bool fused_count = bucket && !distributed(c);
  if (fused_count) {
    cnt1.alloc((size_t)1 << plan.b1, s);
    CB_CUDA(cudaMemsetAsync(cnt1.p, 0, ((size_t)1 << plan.b1) * 4, s));
  }

  // ---- keygen ----
  timer_start(c, "keygen");
  DevBuf<u64> keys[2];
  DevBuf<u32> vals[2];
  guarded_defer(c, [&] {
    for (int i = 0; i < 2; i++) { keys[i].alloc(Tmax ? Tmax : 1, s); vals[i].alloc(Tmax ? Tmax : 1, s); }
  });
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_TUPLES_INVALID, 0, (DC_COUNT - DC_TUPLES_INVALID) * sizeof(u64), s));
  if (n_own && !failed_locally(c)) {
    int blocks = div_up(Tmax, 256);
    int cap = kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    c->lc.begin("keygen");
    // (while the replicated node table is still being gathered, this shard's own nodes are read from own_seq / own_cov)
    const u64* kg_seq = c->gather_pending_seq ? c->own_seq.p - (size_t)c->node_lo * c->NW : c->node_seq.p;
    const u32* kg_cov = c->gather_pending_seq ? c->own_cov.p - (size_t)c->node_lo : c->node_cov.p;
    CB_DISPATCH_NW(c->NW, (keygen_kernel<NW_><<<blocks, 256, fused_count ? ((size_t)4 << plan.b1) : 0, s>>>(
                              kg_seq, kg_cov, c->node_lo, n_own, L, K, W, pb, keys[0].p, vals[0].p,
                              (ull*)c->dcount.p, fused_count ? cnt1.p : nullptr, plan.shift1, plan.b1)));
    c->lc.end("keygen");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "keygen");

  // ---- shuffle replacement ----
  if (distributed(c)) {
    // partition by owner rank (one stable radix pass over the owner bits of the key), then the
    // all-to-all that replaces the MapReduce shuffle
    timer_start(c, "exchange_tuples");
    int w0 = 0;
    std::vector<u64> cnt(world, 0), rcnt(world);
    guarded_defer(c, [&] {
      if (failed_locally(c)) return;
      w0 = radix_sort(keys[0].p, keys[1].p, vals[0].p, vals[1].p, Tmax, c->pbit, c->pbit + c->lg_world, s, c->lc,
                      "partition");
      DevBuf<u64> dcnt;
      dcnt.alloc(world, s);
      CB_CUDA(cudaMemsetAsync(dcnt.p, 0, world * sizeof(u64), s));
      sorted_digit_counts(keys[w0].p, Tmax, c->pbit, wmask, world, dcnt.p, s);  // (the tuples are sorted by owner now)
      c->lc.n++;
      c->lc.rb.get(cnt.data(), dcnt.p, world * sizeof(u64), s);
      c->lc.rb.sync(s);
    });
    // ONE gather: the whole count matrix (row = sender) and the failure flags.  Every rank sees what every
    // rank will receive, so the capacity check and the choice of pipeline need no further round trips.
    if (failed_locally(c)) std::fill(cnt.begin(), cnt.end(), 0);
    const Gathered G = ex_gather_checked(c, cnt.data(), world);
    size_t Trecv = 0;
    bool all_fit = true;
    for (int r = 0; r < world; r++) {
      u64 t = 0;
      for (int q = 0; q < world; q++) t += G.at(q, r);
      if (t >= (1ull << 32))
        throw Error(CB_E_CAPACITY, "more than 2^32 k-mer tuples per context (rank " + std::to_string(r) + ")");
      if (!make_bucket_plan((size_t)t).fits) all_fit = false;
      if (r == c->cfg.rank) Trecv = (size_t)t;
      rcnt[r] = G.at(r, c->cfg.rank);
    }
    DevBuf<u64> rkeys;
    DevBuf<u32> rvals;
    timer_start(c, "a2a_tuples");
    {
      // one kernel of stores into the peers' mailboxes over NVLink; NCCL's grouped send/recv otherwise
      const void* snd[2] = {keys[w0].p, vals[w0].p};
      const size_t el[2] = {8, 4};
      void* rcv[2] = {nullptr, nullptr};
      if (p2p_exchange(c, G, 2, snd, el, rcv)) {
        rkeys.borrow((u64*)rcv[0], Trecv ? Trecv : 1);
        rvals.borrow((u32*)rcv[1], Trecv ? Trecv : 1);
        c->counters["p2p_exchange"] = 1;
      } else {
        rkeys.alloc(Trecv ? Trecv : 1, s);
        rvals.alloc(Trecv ? Trecv : 1, s);
        ex_alltoallv2(c, keys[w0].p, rkeys.p, 8, vals[w0].p, rvals.p, 4, cnt.data(), rcnt.data());  // one transfer group
        c->counters["p2p_exchange"] = 0;
      }
    }
    timer_stop(c, "a2a_tuples");
    launch_node_gather(c);  // the replicated node table follows the tuples over the NVLink, behind the hash partition
    Tmax = Trecv;
    keys[0] = std::move(rkeys);
    vals[0] = std::move(rvals);
    keys[1].alloc(Tmax ? Tmax : 1, s);
    vals[1].alloc(Tmax ? Tmax : 1, s);
    plan = make_bucket_plan(Tmax);
    if (!all_fit) bucket = false;  // the same pipeline on every rank
    timer_stop(c, "exchange_tuples");
  }

  JoinParams P;
  memset(&P, 0, sizeof(P));
  P.L = L; P.K = K; P.pb = pb;
  P.kmask = kmask;
  P.node_seq = c->node_seq.p;
  P.node_cov = c->node_cov.p;
  P.max_cov = (u32)c->gcount[DC_MAX_COV];
  P.low_kmer = c->cfg.low_kmer;
  P.up_kmer = c->cfg.up_kmer;
  P.lay = c->lay;
  P.dcount = (ull*)c->dcount.p;
  c->NWo = (L - K + 15) / 16;
  P.nwo = c->NWo;

  DevBuf<u32> bstart, d_total, extra;
  DevBuf<ull> slot_cursor;
  DevBuf<uint2> adj;
  DevBuf<uint4> list_order;
  size_t n_list_order = 0;
  DevBuf<u64> slots, xbuf;
  DevBuf<u32> ovh;
  adj.alloc(n_lists ? n_lists : 1, s);
  extra.alloc(n_lists ? n_lists : 1, s);
  u32 T = 0;
  size_t slots1 = 0;
  u64 n_x = 0;
  int where = 0;        // which of keys[]/vals[] holds the tuples the join reads
  bool sorted = false;  // flat path: the tuples are sorted by k-mer

  timer_start(c, "sort_tuples");
  if (bucket) {
    bstart.alloc((size_t)plan.NB + 1, s);
    d_total.alloc(1, s);
    guarded_defer(c, [&] {  // a failure here surfaces at the gather after the join attempt
      bucket_partition(keys[0].p, vals[0].p, keys[1].p, vals[1].p, Tmax, kmask, plan, fused_count ? cnt1.p : nullptr,
                       bstart.p, d_total.p, s, c->lc);
      T = read_u32(c, d_total.p);
    });
    where = 0;
  }
  // the LSD sort of the flat pipeline over the first n_sort tuple slots (all-ones keys of windows that emit
  // nothing go last and are dropped)
  auto flat_sort = [&](size_t n_sort) {
    guarded(c, [&] { where = radix_sort(keys[0].p, keys[1].p, vals[0].p, vals[1].p, n_sort, 0, 2 * K, s, c->lc, "tuples"); });
    DevBuf<ull> inval;
    inval.alloc(1, s);
    CB_CUDA(cudaMemsetAsync(inval.p, 0, sizeof(ull), s));
    if (n_sort) {
      int blocks = div_up(n_sort, 256 * 16);
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      count_invalid_kernel<<<blocks, 256, 0, s>>>(keys[where].p, n_sort, inval.p);
      c->lc.n++;
    }
    ull h_inval = 0;
    CB_CUDA(cudaMemcpyAsync(&h_inval, inval.p, sizeof(ull), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    T = (u32)(n_sort - h_inval);
    sorted = true;
  };
  if (!bucket) flat_sort(Tmax);
  timer_stop(c, "sort_tuples");

  // ---- join + verify + mirror ----
  wait_node_table(c, false);  // packed reads + coverage of every rank's nodes
  timer_start(c, "join_verify");
  struct L2Window {  // cleared on every way out of this function
    cb_ctx* c;
    ~L2Window() { clear_l2_window(c); }
  } l2_window{c};
  set_l2_window(c, c->node_seq.p, c->node_seq.bytes());
  DevBuf<u32> slotcnt, off1, seg_head, ne_list;  // flat path only
  DevBuf<u32> extra_total;                       // bucket path: slots added to the lists over the attempts
  u64 xcap = (u64)T / 4 + 4096;
  bool have_extra = false;
  int refine = 0;  // bucket path: times the partition was refined after a bucket overflow
  auto flat_prepare = [&] {  // group walk: seg_head, ne_list, list sizes
    slotcnt.alloc((size_t)n_lists + 1, s);
    off1.alloc((size_t)n_lists + 1, s);
    seg_head.alloc(T ? T : 1, s);
    ne_list.alloc((size_t)T + 1, s);
    CB_CUDA(cudaMemsetAsync(slotcnt.p, 0, ((size_t)n_lists + 1) * 4, s));
    CB_CUDA(cudaMemsetAsync(ne_list.p, 0xff, ((size_t)T + 1) * 4, s));  // terminators; node entries claim from the left
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_HKMER, 0, sizeof(u64), s));
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_UPKMER_LISTS, 0, sizeof(u64), s));
    P.keys = keys[where].p;
    P.vals = vals[where].p;
    P.T = T;
    P.slotcnt = slotcnt.p;
    P.seg_head = seg_head.p;
    P.ne_list = ne_list.p;
    int wblocks = div_up((size_t)T, GW_THREADS);
    if (wblocks > kNumSMs * 8) wblocks = kNumSMs * 8;
    if (wblocks < 1) wblocks = 1;
    c->lc.begin("group_walk");
    group_walk_kernel<<<wblocks, GW_THREADS, 0, s>>>(P);
    c->lc.end("group_walk");
    c->lc.n++;
    exclusive_scan_u32(slotcnt.p, off1.p, (size_t)n_lists + 1, s, c->lc);
    slots1 = read_u32(c, off1.p + n_lists);
  };
  if (bucket) slots1 = (size_t)T * 2 + (size_t)T / 8 + 65536;  // two node entries per group is the rule; retried if short
  else flat_prepare();
  for (int attempt = 0;; attempt++) {
    if (slots1 >= (1ull << 32)) throw Error(CB_E_CAPACITY, "more than 2^32 adjacency slots per context");
    slots.alloc(slots1 ? slots1 : 1, s);
    xbuf.alloc(xcap, s);
    ovh.alloc((slots1 ? slots1 : 1) * (size_t)c->NWo, s);
    P.slots = slots.p;
    P.ovh = ovh.p;
    P.xbuf = xbuf.p;
    P.xcap = xcap;
    // (DC_HKMER and DC_UPKMER_LISTS of the flat path were counted by group_walk_kernel and stay)
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGES_RAW, 0, (DC_UPKMER_LISTS - DC_EDGES_RAW) * sizeof(u64), s));
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_MAX_LIST, 0, sizeof(u64), s));
    u64 x_over = 0;
    if (bucket) {
      slot_cursor.alloc(2, s);
      list_order.alloc(n_lists ? n_lists : 1, s);
      CB_CUDA(cudaMemsetAsync(slot_cursor.p, 0, 2 * sizeof(ull), s));
      CB_CUDA(cudaMemsetAsync(adj.p, 0, (size_t)(n_lists ? n_lists : 1) * sizeof(uint2), s));
      CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_HKMER, 0, sizeof(u64), s));
      CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_UPKMER_LISTS, 0, sizeof(u64), s));
      CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_BUCKET_OVER, 0, sizeof(u64), s));
      P.keys = keys[where].p;
      P.vals = vals[where].p;
      P.T = T;
      P.bstart = bstart.p;
      P.lshift = plan.lshift;
      P.adj = adj.p;
      P.extra = have_extra ? extra_total.p : nullptr;
      P.slot_cursor = slot_cursor.p;
      P.slot_cap = slots1;
      P.list_order = list_order.p;
      ull cur[2] = {0, 0};
      if (!failed_locally(c)) {
        c->lc.begin("bucket_join");
        static const bool force_m64 = getenv("CB_FORCE_M64") != nullptr;  // test hook: the wide-payload variant on small inputs
        const bool m32 = nbits + pb + 1 <= 32 && !force_m64;  // node | window | dir fit the 32-bit payload
        switch (c->NW) {
          case 2: launch_bucket_join<2, 7, 1024>(P, plan.NB, m32, s); break;
          case 4: launch_bucket_join<4, 4, 2048>(P, plan.NB, m32, s); break;
          case 6: launch_bucket_join<6, 4, 2048>(P, plan.NB, m32, s); break;
          case 8: launch_bucket_join<8, 4, 2048>(P, plan.NB, m32, s); break;
          default: throw Error(CB_E_UNSUPPORTED, "read length above 256 bases");
        }
        c->lc.end("bucket_join");
        c->lc.n++;
        if (have_extra && n_lists) {
          orphan_lists_kernel<<<div_up((size_t)n_lists, 256), 256, 0, s>>>(extra_total.p, n_lists, adj.p, slots.p, slot_cursor.p,
                                                                          (ull)slots1, list_order.p);
          c->lc.n++;
        }
        CB_CUDA(cudaGetLastError());
        c->lc.rb.get(cur, slot_cursor.p, 2 * sizeof(ull), s);
        download_counters(c);
      }
      const ull used = cur[0];
      n_list_order = (size_t)cur[1];
      n_x = c->hcount[DC_EDGES_RAW];
      // every rank takes the same branch (collectives follow): one gather for the three retry conditions
      // and whatever failed locally since the tuple exchange
      u64 flags[3] = {c->hcount[DC_BUCKET_OVER] ? 1ull : 0ull, used > slots1 ? 1ull : 0ull, n_x > xcap ? 1ull : 0ull};
      if (distributed(c)) {
        const Gathered G = ex_gather_checked(c, flags, 3);
        for (int j = 0; j < 3; j++) flags[j] = G.max(j);
      }
      x_over = flags[2];
      if (flags[0] && plan.b1 + plan.b2 < 20 && refine < 2) {
        // a bucket does not fit the kernel: cut the tuples (already compacted to the front of keys[0]) into
        // four times as many buckets first
        refine++;
        plan = make_bucket_plan_bits(plan.b1 + plan.b2 + 2 > 20 ? 20 : plan.b1 + plan.b2 + 2);
        bstart.alloc((size_t)plan.NB + 1, s);
        guarded_defer(c, [&] {
          bucket_partition(keys[0].p, vals[0].p, keys[1].p, vals[1].p, T, kmask, plan, nullptr, bstart.p, d_total.p, s, c->lc);
        });
        attempt = -1;
        continue;
      }
      if (flags[0]) {
        // still does not fit (one k-mer with thousands of tuples, or thousands of node entries in one group):
        // the flat pipeline takes over for this input
        bucket = false;
        flat_sort(T);  // the partition has compacted the T valid tuples to the front of keys[0] / vals[0]
        flat_prepare();
        attempt = -1;
        continue;
      }
      if (flags[1]) {  // more node entries per group than expected: the exact need is known now
        if (attempt >= 4) throw Error(CB_E_CAPACITY, "adjacency slot capacity");
        if (used > slots1) slots1 = (size_t)used + 1024;
        if (n_x > xcap) xcap = n_x;  // (known already: saves an attempt)
        continue;
      }
      slots1 = (size_t)used;
    } else {
      CB_CUDA(cudaMemsetAsync(slots.p, 0xff, (slots1 ? slots1 : 1) * sizeof(u64), s));  // empty slots read as ~0
      P.off1 = off1.p;
      make_adj_kernel<<<div_up((size_t)(n_lists ? n_lists : 1), 256), 256, 0, s>>>(off1.p, slotcnt.p, n_lists, adj.p);
      int jgrid = div_up((size_t)T, 256);
      if (jgrid > kNumSMs * 8) jgrid = kNumSMs * 8;
      if (jgrid < 1) jgrid = 1;
      c->lc.begin("join_verify");
      {
        static const int minb = getenv("CB_JOIN_MINB") ? atoi(getenv("CB_JOIN_MINB")) : 4;
        if (c->NW == 2 && minb == 4) join_verify_kernel<2, 4><<<jgrid, 256, 0, s>>>(P);
        else if (c->NW == 2 && minb == 3) join_verify_kernel<2, 3><<<jgrid, 256, 0, s>>>(P);
        else CB_DISPATCH_NW(c->NW, (join_verify_kernel<NW_, 2><<<jgrid, 256, 0, s>>>(P)));
      }
      c->lc.end("join_verify");
      c->lc.n += 2;
      CB_CUDA(cudaGetLastError());
      download_counters(c);
      n_x = c->hcount[DC_EDGES_RAW];
      x_over = n_x > xcap ? 1 : 0;
      ex_allreduce(c, &x_over, 1, 1);  // every rank must take the same branch: collectives follow
    }
    if (x_over) {  // side buffer too small: its exact size is known now
      if (attempt >= 4) throw Error(CB_E_CAPACITY, "side buffer overflow");
      if (n_x > xcap) xcap = n_x;
      continue;
    }
    // side-buffer records into their lists (on the rank that owns the list)
    CB_CUDA(cudaMemsetAsync(extra.p, 0, (size_t)(n_lists ? n_lists : 1) * 4, s));
    const u64* xin = xbuf.p;
    size_t n_xin = n_x;
    DevBuf<char> xrecv;
    if (distributed(c)) {
      timer_start(c, "side_route");
      DevBuf<u32> xdest;
      xdest.alloc(n_x ? n_x : 1, s);
      if (n_x) {
        CB_DISPATCH_NW(c->NW, (key_owner_kernel<NW_><<<div_up(n_x, 256), 256, 0, s>>>(
                                  xbuf.p, n_x, c->lay, c->node_seq.p, L, K, c->pbit, wmask, xdest.p)));
        c->lc.n++;
      }
      route_records(c, xbuf.p, 8, xdest.p, n_x, xrecv, n_xin);
      xin = (const u64*)xrecv.p;
      timer_stop(c, "side_route");
    }
    if (n_xin) {
      int blocks = div_up(n_xin, 256);
      if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
      c->lc.begin("side_insert");
      CB_DISPATCH_NW(c->NW, (side_insert_kernel<NW_><<<blocks, 256, 0, s>>>(xin, n_xin, c->lay, adj.p, slots.p, extra.p,
                                                                           (ull*)c->dcount.p, c->node_seq.p, ovh.p,
                                                                           c->NWo, L)));
      c->lc.end("side_insert");
      c->lc.n++;
      CB_CUDA(cudaGetLastError());
    }
    download_counters(c);
    u64 slot_over = c->hcount[DC_MAX_LIST];
    ex_allreduce(c, &slot_over, 1, 0);
    if (slot_over == 0) break;
    if (attempt >= 4) throw Error(CB_E_CAPACITY, "adjacency slot overflow");
    // enlarge the slot ranges of the lists that overflowed and run the join again
    if (bucket) {
      if (!have_extra) {
        extra_total.alloc(n_lists ? n_lists : 1, s);
        CB_CUDA(cudaMemsetAsync(extra_total.p, 0, (size_t)(n_lists ? n_lists : 1) * 4, s));
        have_extra = true;
      }
      add_u32_kernel<<<div_up((size_t)n_lists, 256), 256, 0, s>>>(extra_total.p, extra.p, n_lists);
      c->lc.n++;
      slots1 += (size_t)c->hcount[DC_MAX_LIST] + 1024;
    } else {
      add_u32_kernel<<<div_up((size_t)n_lists, 256), 256, 0, s>>>(slotcnt.p, extra.p, n_lists);
      c->lc.n++;
      exclusive_scan_u32(slotcnt.p, off1.p, (size_t)n_lists + 1, s, c->lc);
      slots1 = read_u32(c, off1.p + n_lists);
    }
  }
  timer_stop(c, "join_verify");
  keys[0].release(); keys[1].release();
  vals[0].release(); vals[1].release();
  (void)sorted;

  // BuildHighKmerList's counter (poly-A/T k-mer handled through its global weight)
  const u64 totals[2] = {T, slots1};
  const Gathered GT = reduce_counters(c, totals, 2);
  c->counters["tuples"] = (int64_t)GT.sum(DC_COUNT);
  int64_t hkmer = (int64_t)c->gcount[DC_HKMER] + ((long long)c->gcount[DC_POLYA_WEIGHT] > c->cfg.up_kmer ? 1 : 0);
  c->counters["hkmer"] = hkmer;
  c->counters["candidates_verified"] = (int64_t)c->gcount[DC_CAND_VERIFIED];
  c->counters["side_buffer_edges"] = (int64_t)c->gcount[DC_EDGES_RAW];
  c->counters["upkmer_truncated_lists"] = (int64_t)c->gcount[DC_UPKMER_LISTS];
  c->counters["order_dependent"] = (int64_t)c->gcount[DC_UPKMER_LISTS];
  c->counters["bucket_path"] = bucket ? 1 : 0;
  c->counters["host_collectives"] = c->n_host_coll;
  c->counters["bucket_refine"] = refine;
  c->counters["bucket_bits"] = plan.b1 + plan.b2;
  c->counters["slots"] = (int64_t)GT.sum(DC_COUNT + 1);
  if (c->gcount[DC_UPKMER_LISTS] > 0)
    throw Error(CB_E_UNSUPPORTED,
                "a k-mer group with a node entry holds more than -kmerup tuples: the reference truncates its "
                "candidate list in Hadoop's value order (MatchPrefix.java:366-368), parity undefined");
  if (hkmer > 0)
    throw Error(CB_E_UNSUPPORTED,
                "non-empty high-frequency k-mer list: the reference's stopword loader (MatchPrefix.java:81) is "
                "ill-defined for it, parity undefined");

  // checkpoint A: the slot array is the adjacency structure
  c->adj = std::move(adj);
  if (bucket) { c->list_order = std::move(list_order); c->n_list_order = n_list_order; }
  else { c->list_order.release(); c->n_list_order = 0; }
  c->slots = slots1;
  c->ckA.keys = std::move(slots);
  c->ovh = std::move(ovh);
  c->ckA.n = (size_t)(c->hcount[DC_EDGES_M] + c->hcount[DC_EDGES_A]);  // the lists this rank owns
  c->ckA.valid = true;
  c->ckA.compact = false;
  c->n_cut_log = 0;
  c->counters["edges_overlap"] = (int64_t)(c->gcount[DC_EDGES_M] + c->gcount[DC_EDGES_A]);
}

}  // namespace cb
