// stage_string.cu -- per-node adjacency pruning: CutChimericLinks, EdgeRemoval and
// TransitiveReduction on the adjacency lists the join left in HBM.
//
// Replaces (reference):
//   CutChimericLinksMapper/Reducer  CutChimericLinks.java:87-114, :258-409  consistency_kernel, cut_kernel
//   Node.Consensus                  Node.java:1293-1377                     cut_kernel (PWM part)
//   EdgeRemovalMapper/Reducer       EdgeRemoval.java:58-87, :121-202        remove_flagged_kernel, tr_final_kernel
//   TransitiveReductionReducer      TransitiveReduction.java:239-547        tr_kernel, tr_final_kernel
//   driver loop                     BrushAssembler.java:337-379             stage_string_graph
//
// The graph of checkpoint A is symmetric (S u mirror(S)), so the neighbour messages each
// reducer rebuilds its lists from (DARKMSG / OVALMSG, sent by the mirror edge) are exactly the
// node's own adjacency list: the records in keys[adj[lid].x .. adj[lid].x + adj[lid].y), lid = 2a + x, in
// no particular order, empty slots ~0.  The overhang of an entry is b^y[ov:], the bases of the
// neighbour that extend past the node; the join stored it next to the record (ovh[slot]).
//
// First every list is tested for "consistency": is every overhang a prefix of the longest one?
// (one thread per list, one pass: consistency_kernel).  For a consistent list (a) the PWM
// consensus of CutChimericLinks equals the longest overhang wherever it is defined, has no N and
// cuts nothing (needs MAJORITY < 1); (b) the greedy scan of TransitiveReduction keeps exactly the
// entries of maximal overlap, so the list is summarised by ONE byte (that overlap) and
// tr_final_kernel decides every record, and the fate of its mirror at the other end, from the
// summaries alone.  Lists that are not consistent (sequencing errors, repeats) go through work
// lists to the literal warp-per-list kernels.  CutChimericLinks removes records in place and logs
// them (ctx.cuh: cut_log), checkpoint B is written as a compact record list.
#include "ctx.cuh"

namespace cb {

struct StringParams {
  const u64* edges;    // slot array; list lid owns [adj[lid].x, adj[lid].x + adj[lid].y); ~0 = empty slot
  const uint2* adj;    // [n_lists] (first slot, slots) of list lid
  u32 n_lists;         // 2 * n_nodes
  size_t slots;
  EdgeLayout lay;
  const u64* node_seq;
  const u32* ovh;      // [slots][nwo] stored overhang of the record in the slot
  int nwo;
  const IdKey* node_idkey;
  const u32* node_cov;
  int L, K;
  float majority, pwm_n;
  uint8_t* removed;    // cut output, per slot
  uint8_t* kept;       // tr output, per slot
  uint8_t* list_flag;  // per list: 1 consistent, 0 not, dirty bits 2 / 4 (lost an entry in cut round 1 / 2)
  uint4* lsumm;        // per list, from consistency_kernel, ONE 16-byte record (lists lie in bucket order in HBM, so
                       // per-list results are scattered writes: one store instead of three): .x, .y = (nbr | y << 31) of
                       // the first two records of maximal overlap (~0: none), .z = maximal overlap | count of such
                       // records (saturating at 255) << 8
  uint4* lseq;         // bucket path, per list_order entry (sequential writes): .x, .y = slots of the first two records of
                       // maximal overlap, .z = maximal overlap | count of such records (sat. 255) << 8 | contained
                       // records (ov == L, sat. 65535) << 16, .w = records TransitiveReduction counts as transitive
  const uint8_t* lsum; // per list, all ranks: ovmax if the list is consistent and untouched by the cut rounds
                       // (TransitiveReduction keeps exactly its records of that overlap), else 0
  const u32* work;     // work list of list ids (literal kernels)
  u32 n_work;
  u32* work_out;       // lists that lost an entry in this round
  u32* n_work_out;
  int check_consistency;  // literal kernels: re-test the list first
  int dirty_bit;          // 2 in the first cut round, 4 in the second
  u64* req;               // multi-GPU: keys of mirror records to act on at their owner rank
  u32* n_req;
  u32 req_cap;
  ull* dcount;
};

// overhang b^y[ov:], left aligned, of the adjacency record in slot `slot`.  The join stored it next
// to the record, so the lists are read sequentially here and no neighbour read is gathered (with the
// node table replicated over several GPUs those gathers missed the L2: 1.5 ms of consistency_kernel
// on one GPU became 5.0 ms on four).
template <int NW>
__device__ __forceinline__ void load_overhang(const StringParams& P, size_t slot, u64 key, u64 (&h)[NW], int& len) {
  fetch_overhang<NW>(P.ovh + slot * (size_t)P.nwo, h, P.nwo);
  len = P.L - (int)P.lay.ov(key);
}

template <int NW>
__device__ __forceinline__ int base_at(const u64 (&h)[NW], int j) {
  u64 w = 0;
#pragma unroll
  for (int i = 0; i < NW; i++)
    if (i == (j >> 5)) w = h[i];
  return (int)((w >> (62 - 2 * (j & 31))) & 3);
}

// slot index of `key` in its (unordered) list or -1 (an 8-wide unrolled scan measured slower in
// tr_final_kernel: 2.71 vs 1.87 ms -- the early exit matters more than the loads in flight)
__device__ __forceinline__ long long find_edge(const StringParams& P, u64 key) {
  const u32 list = P.lay.list(key);
  const uint2 lr_ = P.adj[list]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
  for (u32 t = lo; t < hi; t++)
    if (P.edges[t] == key) return (long long)t;
  return -1;
}

constexpr int ST_WARPS = 8;

// ---------------------------------------------------------------------------------------------
// consistency of every list, one thread per list: one scan finds the longest overhang (minimal
// overlap, first such record) and the maximal overlap, a second pass compares every overhang
// with the longest one and precomputes what the fast path of TransitiveReduction needs
// (tr_final_kernel): bit0 maximal overlap, bit1 counts as transitive, bit2 ov == L.
// ---------------------------------------------------------------------------------------------
// `lists` (several GPUs): the non-empty lists, i.e. the ones this rank owns -- without it three
// of four lanes of a warp would idle on four GPUs
template <int NW>
__global__ void __launch_bounds__(256) consistency_kernel(StringParams P, u32* __restrict__ work_out,
                                                          u32* __restrict__ n_work_out, int force_slow,
                                                          const u32* __restrict__ lists,
                                                          const uint4* __restrict__ lists4, u32 n_iter) {
  for (u32 it = blockIdx.x * blockDim.x + threadIdx.x; it < n_iter; it += gridDim.x * blockDim.x) {
    // lists4 (bucket path): (list, first slot, slots) in slot order -- no gather of adj[]
    u32 lid, lo, hi;
    if (lists4) {
      const uint4 q = lists4[it];
      lid = q.x; lo = q.y; hi = q.y + q.z;
    } else {
      lid = lists ? lists[it] : it;
      const uint2 lr_ = P.adj[lid];
      lo = lr_.x; hi = lr_.x + lr_.y;
    }
    // ONE pass over the list: the overhangs nest iff every record nests with the longest overhang
    // seen so far (if the new one is longer it must extend that one, and then everything before it
    // is a prefix of the new one as well)
    u32 ovmax = 0, n = 0, nmax = 0;
    u32 m0 = 0xffffffffu, m1 = 0xffffffffu;  // (nbr | y << 31) of the first two records of maximal overlap
    u32 s0 = 0xffffffffu, s1 = 0xffffffffu;  // ... and their slots
    u64 H[NW];
    int lenH = -1;
    bool bad = false;
    for (u32 t = lo; t < hi; t++) {
      const u64 k = P.edges[t];
      if (k == ~0ull) continue;
      n++;
      const u32 o = P.lay.ov(k);
      const u32 nby = P.lay.nbr(k) | (P.lay.y(k) << 31);
      if (o > ovmax) { ovmax = o; nmax = 1; m0 = nby; m1 = 0xffffffffu; s0 = t; s1 = 0xffffffffu; }
      else if (o == ovmax) { if (nmax == 1) { m1 = nby; s1 = t; } nmax++; }
      if (!force_slow) {
        u64 h[NW];
        int len;
        load_overhang<NW>(P, t, k, h, len);
        if (lenH < 0) {
#pragma unroll
          for (int w = 0; w < NW; w++) H[w] = h[w];
          lenH = len;
        } else if (len > lenH) {
          bad = bad || !prefix_equal<NW>(H, h, lenH);
#pragma unroll
          for (int w = 0; w < NW; w++) H[w] = h[w];
          lenH = len;
        } else {
          bad = bad || !prefix_equal<NW>(h, H, len);
        }
      }
    }
    if (lists4) {
      // bucket path: the summary goes to the list's place in slot order (sequential), and what
      // TransitiveReduction will count on this list if it stays consistent is taken now, while the list is in
      // L1: records below the maximal overlap are transitive unless a kept record has their type and
      // neighbour (TransitiveReduction.java:325-334), records with ov == L are "contained" (:319-323)
      u32 trans = 0, contained = 0;
      for (u32 t = lo; t < hi && n > nmax; t++) {
        const u64 k = P.edges[t];
        if (k == ~0ull) continue;
        const u32 o = P.lay.ov(k);
        if (o == ovmax) continue;
        const u32 nby = P.lay.nbr(k) | (P.lay.y(k) << 31);
        bool step2 = nby == m0 || nby == m1;
        for (u32 q = lo; nmax > 2 && !step2 && q < hi; q++) {
          const u64 kq = P.edges[q];
          step2 = kq != ~0ull && P.lay.ov(kq) == ovmax && (P.lay.nbr(kq) | (P.lay.y(kq) << 31)) == nby;
        }
        trans += !step2;
      }
      if ((int)ovmax == P.L) contained = nmax;  // only the maximal overlap can be L
      P.lseq[it] = make_uint4(s0, s1, (n ? ovmax : 0u) | ((nmax > 255 ? 255u : nmax) << 8) |
                                          ((contained > 65535u ? 65535u : contained) << 16), trans);
      if (n == 0) continue;
    } else {
      if (n == 0) { P.lsumm[lid] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u); continue; }
      P.lsumm[lid] = make_uint4(m0, m1, ovmax | ((nmax > 255 ? 255u : nmax) << 8), 0u);
    }
    if (n > 1 && (bad || force_slow)) {  // only this thread writes the list's flag here
      P.list_flag[lid] = 0;
      work_out[atomicAdd(n_work_out, 1u)] = lid;
    }
  }
}

// non-empty lists in ascending order of their id within a warp (warp-aggregated append)
__global__ void __launch_bounds__(256) own_lists_kernel(const uint2* __restrict__ adj, u32 n_lists,
                                                        u32* __restrict__ out, u32* __restrict__ n_out) {
  const u32 lid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool have = lid < n_lists && adj[lid].y > 0;
  const u32 bal = __ballot_sync(0xffffffffu, have);
  if (bal == 0) return;
  u32 base = 0;
  if (lane == 0) base = atomicAdd(n_out, (u32)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (have) out[base + __popc(bal & ((1u << lane) - 1))] = lid;
}

// ---------------------------------------------------------------------------------------------
// K10: CutChimericLinks, literal path, one warp per queued list
// ---------------------------------------------------------------------------------------------
struct TopKey {  // order of Node.Consensus after the reducer's pre-sort: overhang length desc
  u32 ov;        // (= ov asc), then id asc for direction f / desc for direction r
  u64 hi, lo;
  u32 idx;
};
__device__ __forceinline__ bool top_less(const TopKey& a, const TopKey& b) {
  if (a.ov != b.ov) return a.ov < b.ov;
  if (a.hi != b.hi) return a.hi < b.hi;
  if (a.lo != b.lo) return a.lo < b.lo;
  return a.idx < b.idx;
}
__device__ __forceinline__ TopKey warp_top_min(TopKey v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    TopKey t;
    t.ov = __shfl_xor_sync(0xffffffffu, v.ov, o);
    t.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
    t.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
    t.idx = __shfl_xor_sync(0xffffffffu, v.idx, o);
    if (top_less(t, v)) v = t;
  }
  return v;
}

// warp version of the consistency test (lists that changed since consistency_kernel ran);
// also returns the number of records
template <int NW>
__device__ __forceinline__ bool list_consistent_warp(const StringParams& P, u32 lo, u32 hi, int lane, u32& n_out) {
  u32 hov = 0xffffffffu, hidx = 0xffffffffu, n = 0;
  for (u32 t = lo + lane; t < hi; t += 32) {
    const u64 q = P.edges[t];
    if (q == ~0ull) continue;
    n++;
    u32 o = P.lay.ov(q);
    if (o < hov) { hov = o; hidx = t; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u32 ov2 = __shfl_xor_sync(0xffffffffu, hov, o), ix2 = __shfl_xor_sync(0xffffffffu, hidx, o);
    if (ov2 < hov || (ov2 == hov && ix2 < hidx)) { hov = ov2; hidx = ix2; }
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  n_out = n;
  if (n <= 1) return true;
  u64 H[NW];
  int lenH;
  load_overhang<NW>(P, hidx, P.edges[hidx], H, lenH);
  bool good = true;
  for (u32 t = lo + lane; t < hi; t += 32) {
    const u64 q = P.edges[t];
    if (q == ~0ull) continue;
    u64 h[NW];
    int len;
    load_overhang<NW>(P, t, q, h, len);
    good = good && prefix_equal<NW>(h, H, len);
  }
  return __all_sync(0xffffffffu, good);
}

// sets this round's dirty bit of the list's flag byte; the first setter queues the list once.
// Any dirty bit makes the flag differ from 1, so TransitiveReduction takes the literal path.
__device__ __forceinline__ void mark_dirty(const StringParams& P, u32 lid) {
  unsigned int* word = reinterpret_cast<unsigned int*>(P.list_flag + (lid & ~3u));
  unsigned int sh = 8 * (lid & 3u);
  unsigned int bit = (unsigned int)P.dirty_bit << sh;
  unsigned int old = atomicOr(word, bit);
  if (!(old & bit)) P.work_out[atomicAdd(P.n_work_out, 1u)] = lid;
}

template <int NW>
__global__ void __launch_bounds__(ST_WARPS * 32) cut_kernel(StringParams P) {
  __shared__ unsigned char s_cons[ST_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L;
  for (u32 wi = warp_global; wi < P.n_work; wi += nwarps) {
    const u32 list = P.work[wi];
    const uint2 lr_ = P.adj[list]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
    u32 n = 0;
    if (P.check_consistency) {
      const bool cons = list_consistent_warp<NW>(P, lo, hi, lane, n);
      if (n <= 1) continue;  // CutChimericLinks.java:306 "flist.size() > 1"
      if (cons && P.majority < 1.0f) continue;
    } else {
      for (u32 t = lo + lane; t < hi; t += 32) n += P.edges[t] != ~0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (n <= 1) continue;
    }
    if (lane == 0) atomicAdd(&P.dcount[DC_SLOW_LISTS], 1ull);
    const u32 x = list & 1;
    // ---- entries 0, 1 of the Consensus order and the overlap of entry 2 (Node.java:1295-1320) ----
    TopKey best[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      TopKey mine;
      mine.ov = 0xffffffffu; mine.hi = ~0ull; mine.lo = ~0ull; mine.idx = 0xffffffffu;
      for (u32 t = lo + lane; t < hi; t += 32) {
        if (r >= 1 && t == best[0].idx) continue;
        if (r >= 2 && t == best[1].idx) continue;
        u64 k = P.edges[t];
        if (k == ~0ull) continue;
        IdKey ik = P.node_idkey[P.lay.nbr(k)];
        TopKey c;
        c.ov = P.lay.ov(k);
        c.hi = x ? ~ik.hi : ik.hi;
        c.lo = x ? ~ik.lo : ik.lo;
        c.idx = t;
        if (top_less(c, mine)) mine = c;
      }
      best[r] = warp_top_min(mine);
    }
    const u32 cov0 = P.node_cov[P.lay.nbr(P.edges[best[0].idx])];
    const u32 cov1 = P.node_cov[P.lay.nbr(P.edges[best[1].idx])];
    int clen;
    if (n == 2 || (float)cov0 + (float)cov1 > 2.0f) clen = L - (int)best[1].ov;
    else clen = L - (int)best[2].ov;
    // ---- position weight matrix, one lane per position (Node.java:1321-1364) ----
    int ncount = 0;
    for (int j0 = 0; j0 < clen; j0 += 32) {
      int j = j0 + lane;
      int cnt[4] = {0, 0, 0, 0};
      int sum = 0;
      for (u32 t = lo; t < hi; t++) {
        u64 k = P.edges[t];
        if (k == ~0ull) continue;
        u64 h[NW];
        int len;
        load_overhang<NW>(P, t, k, h, len);
        int w = (int)P.node_cov[P.lay.nbr(k)];
        if (j < len && j < clen) {
          int b = base_at<NW>(h, j);
          sum += w;
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (q == b) cnt[q] += w;
        }
      }
      int code = 4;  // N
      if (j < clen) {
        float fs = (float)sum;
        if (__fdiv_rn((float)cnt[0], fs) > P.majority) code = 0;       // A
        else if (__fdiv_rn((float)cnt[3], fs) > P.majority) code = 3;  // T
        else if (__fdiv_rn((float)cnt[1], fs) > P.majority) code = 1;  // C
        else if (__fdiv_rn((float)cnt[2], fs) > P.majority) code = 2;  // G
        s_cons[warp][j] = (unsigned char)code;
      }
      ncount += __popc(__ballot_sync(0xffffffffu, j < clen && code == 4));
    }
    __syncwarp();
    // Node.java:1369-1371 (0/0 = NaN compares false, as in Java)
    if (__fdiv_rn((float)ncount, (float)clen) > P.pwm_n) continue;
    // ---- entries that disagree with the consensus (CutChimericLinks.java:337-381) ----
    for (u32 t0 = lo; t0 < hi; t0 += 32) {
      u32 t = t0 + lane;
      u64 k = t < hi ? P.edges[t] : ~0ull;
      if (k != ~0ull) {
        u64 h[NW];
        int len;
        load_overhang<NW>(P, t, k, h, len);
        int m = len < clen ? len : clen;
        bool miss = false;
        for (int j = 0; j < m; j++) {
          int cc = s_cons[warp][j];
          if (cc != 4 && cc != base_at<NW>(h, j)) { miss = true; break; }
        }
        if (miss) {
          P.removed[t] = 1;
          mark_dirty(P, list);
          u64 mk = P.lay.mirror(k);
          if (P.req) {  // the mirror's list may live on another rank: request its removal there
            u32 ri = atomicAdd(P.n_req, 1u);
            if (ri < P.req_cap) P.req[ri] = mk;
          } else {
            long long mi = find_edge(P, mk);
            if (mi >= 0) { P.removed[mi] = 1; mark_dirty(P, P.lay.list(mk)); }
          }
          atomicAdd(&P.dcount[DC_EDGE_REMOVAL], 1ull);
        }
      }
    }
    __syncwarp();
  }
}

// EdgeRemoval (EdgeRemoval.java:121-202): a flagged record's slot becomes empty.  Only the lists
// that lost a record in this round are visited (one warp per list); the record and its slot are
// logged so that 01-overlap can be restored for an export.  Clears the flags it consumes.
__global__ void __launch_bounds__(256) remove_flagged_kernel(u64* __restrict__ keys, uint8_t* __restrict__ drop,
                                                             const uint2* __restrict__ adj,
                                                             const u32* __restrict__ lists, u32 n_lists_dirty,
                                                             u64* __restrict__ log, u32* __restrict__ n_log, u32 log_cap,
                                                             ull* __restrict__ dcount) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  u32 removed = 0;
  for (u32 wi = warp_global; wi < n_lists_dirty; wi += nwarps) {
    const u32 lid = lists[wi];
    const uint2 lr_ = adj[lid]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
    for (u32 t = lo + lane; t < hi; t += 32)
      if (drop[t]) {
        const u32 at = atomicAdd(n_log, 1u);
        if (at < log_cap) { log[2 * (size_t)at] = t; log[2 * (size_t)at + 1] = keys[t]; }
        keys[t] = ~0ull;
        drop[t] = 0;
        removed++;
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
  if (lane == 0 && removed) atomicAdd(&dcount[DC_EDGES_C], (ull)removed);
}

// writes logged records back into their slots (restores 01-overlap)
__global__ void __launch_bounds__(256) restore_log_kernel(u64* __restrict__ keys, const u64* __restrict__ log, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[log[2 * i]] = log[2 * i + 1];
}

// ---------------------------------------------------------------------------------------------
// K11: TransitiveReduction
// ---------------------------------------------------------------------------------------------
// lists whose flag is not "consistent" -> work list for the literal kernel; every list's summary
// byte for the final pass (lsum)
__global__ void __launch_bounds__(256) tr_worklist_kernel(const uint8_t* __restrict__ list_flag,
                                                          const uint2* __restrict__ adj,
                                                          const uint4* __restrict__ lsumm, u32 n_lists,
                                                          u32* __restrict__ work, u32* __restrict__ n_work,
                                                          uint8_t* __restrict__ lsum) {
  u32 lid = blockIdx.x * blockDim.x + threadIdx.x;
  if (lid >= n_lists) return;
  const bool nonempty = adj[lid].y > 0;
  const bool fast = list_flag[lid] == 1;
  lsum[lid] = (nonempty && fast) ? (uint8_t)(lsumm[lid].z & 0xffu) : (uint8_t)0;
  if (nonempty && !fast) work[atomicAdd(n_work, 1u)] = lid;
}

// bucket path: the same over the lists that own slots (list_order); every other list keeps lsum = 0 (memset)
__global__ void __launch_bounds__(256) tr_worklist_seq_kernel(const uint8_t* __restrict__ list_flag,
                                                              const uint4* __restrict__ list_order,
                                                              const uint4* __restrict__ lseq, u32 n_order,
                                                              u32* __restrict__ work, u32* __restrict__ n_work,
                                                              uint8_t* __restrict__ lsum) {
  const u32 it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_order) return;
  const u32 lid = list_order[it].x;
  const u32 om = lseq[it].z & 0xffu;  // 0: no record in the list
  const bool fast = list_flag[lid] == 1;
  if (om && fast) lsum[lid] = (uint8_t)om;
  if (om && !fast) work[atomicAdd(n_work, 1u)] = lid;
}

constexpr int TR_KC = 64;  // kept entries per list in the literal scan

// literal greedy scan (TransitiveReduction.java:300-409), one warp per queued list.  Entries are
// visited by overlap descending; entries of equal overlap skip each other (:341-343) and cannot
// be the same edge, so a whole overlap level is decided at once against the entries kept from
// larger overlaps.
template <int NW>
__global__ void __launch_bounds__(ST_WARPS * 32) tr_kernel(StringParams P) {
  __shared__ u64 s_kh[ST_WARPS][TR_KC][NW];
  __shared__ u64 s_kkey[ST_WARPS][TR_KC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L;
  u32 trans = 0, contained = 0;
  for (u32 wi = warp_global; wi < P.n_work; wi += nwarps) {
    const u32 list = P.work[wi];
    const uint2 lr_ = P.adj[list]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
    if (lane == 0) atomicAdd(&P.dcount[DC_SLOW_LISTS], 1ull);
    u32 nk = 0;
    int level = L + 1;  // overlap levels strictly below `level` are still to do
    for (;;) {
      // next level: the largest overlap below `level`
      int nxt = -1;
      for (u32 t = lo + lane; t < hi; t += 32) {
        const u64 q = P.edges[t];
        if (q == ~0ull) continue;
        int o = (int)P.lay.ov(q);
        if (o < level && o > nxt) nxt = o;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nxt = max(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
      if (nxt < 0) break;
      level = nxt;
      const u32 nk_level = nk;  // entries of this level are compared with earlier levels only
      for (u32 t0 = lo; t0 < hi; t0 += 32) {
        u32 t = t0 + lane;
        u64 k = t < hi ? P.edges[t] : ~0ull;
        bool mine = k != ~0ull && (int)P.lay.ov(k) == level;
        bool keep = false;
        u64 h[NW];
        if (mine) {
          int len;
          load_overhang<NW>(P, t, k, h, len);
          if (level == L) contained++;
          bool step2 = false, step3 = false;
          for (u32 q = 0; q < nk_level; q++) {  // :325-334 comes before the transitive test
            u64 kq = s_kkey[warp][q];
            if (P.lay.nbr(kq) == P.lay.nbr(k) && P.lay.y(kq) == P.lay.y(k)) { step2 = true; break; }
          }
          for (u32 q = 0; q < nk_level && !step2; q++) {
            u64 kq = s_kkey[warp][q];
            u64 hq[NW];
#pragma unroll
            for (int w = 0; w < NW; w++) hq[w] = s_kh[warp][q][w];
            if (prefix_equal<NW>(h, hq, L - (int)P.lay.ov(kq))) { step3 = true; break; }  // startsWith :345
          }
          if (step3) trans++;
          keep = !step2 && !step3;
        }
        u32 bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
          u32 pos = nk + __popc(bal & lt_mask);
          if (keep) {
            if (pos < TR_KC) {
#pragma unroll
              for (int w = 0; w < NW; w++) s_kh[warp][pos][w] = h[w];
              s_kkey[warp][pos] = k;
              P.kept[t] = 1;
            } else {
              atomicAdd(&P.dcount[DC_TR_OVERFLOW], 1ull);
            }
          }
          nk += __popc(bal);
          if (nk > TR_KC) nk = TR_KC;
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    trans += __shfl_xor_sync(0xffffffffu, trans, o);
    contained += __shfl_xor_sync(0xffffffffu, contained, o);
  }
  if (lane == 0) {
    if (trans) atomicAdd(&P.dcount[DC_TRANS_EDGE], (ull)trans);
    if (contained) atomicAdd(&P.dcount[DC_CONTAINED_EDGE], (ull)contained);
  }
}

// TransitiveReduction on the consistent lists fused with the EdgeRemoval that follows it
// (TransitiveReduction.java:300-409 + EdgeRemoval.java:121-202): an edge survives iff both
// endpoints kept it.  One thread per slot.  A record of a consistent list is kept iff its overlap
// is the list's maximum (precomputed by consistency_kernel: `pre`); whether the MIRROR was kept is
// read off the mirror list's summary byte -- lsum[list] = that list's maximal overlap if it is
// consistent -- so no list is searched and, on several GPUs, nothing is exchanged per edge.  Only a
// mirror in an inconsistent list (literal tr_kernel) is looked up (one GPU) or was voted for by its
// owner (VOTED, several GPUs: bit 1 of kept[slot]).
template <bool VOTED>
__global__ void __launch_bounds__(256) tr_final_kernel(StringParams P, u64* __restrict__ out) {
  __shared__ u64 s_out[8][64];
  u32 tot = 0, trans = 0, contained = 0;
  u32 wn = 0;  // survivors staged by this warp (warp-uniform)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  // block-uniform trip count: every warp reaches the ballot below with all its lanes
  for (size_t slot0 = (size_t)blockIdx.x * blockDim.x; slot0 < P.slots; slot0 += (size_t)gridDim.x * blockDim.x) {
    const size_t slot = slot0 + threadIdx.x;
    const u64 k = slot < P.slots ? P.edges[slot] : ~0ull;
    bool keep = false;
    if (k != ~0ull) {
      bool kept_here;
      const u32 lid = P.lay.list(k);
      const u32 om = P.lsum[lid];  // the list's maximal overlap if it is consistent and untouched, else 0
      if (om) {
        // TransitiveReduction on a consistent list: keep the maximal overlap (:341-343); anything
        // else is dropped, counted as transitive unless a kept record has my type and neighbour
        // (:325-334)
        const u32 ov = P.lay.ov(k);
        kept_here = ov == om;
        contained += (int)ov == P.L;  // TransitiveReduction.java:319-323 (equal-length reads)
        if (!kept_here) {
          const u32 nby = P.lay.nbr(k) | (P.lay.y(k) << 31);
          const uint4 mm = P.lsumm[lid];
          bool step2 = mm.x == nby || mm.y == nby;
          if (!step2 && ((mm.z >> 8) & 0xffu) > 2) {
            const uint2 lr_ = P.adj[lid]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
            for (u32 q = lo; q < hi && !step2; q++) {
              const u64 kq = P.edges[q];
              if (kq != ~0ull && P.lay.ov(kq) == om && (P.lay.nbr(kq) | (P.lay.y(kq) << 31)) == nby) step2 = true;
            }
          }
          trans += !step2;
        }
      } else {
        kept_here = (P.kept[slot] & 1) != 0;  // literal kernel
      }
      if (kept_here) {
        const u64 mk = P.lay.mirror(k);
        const uint8_t ms = P.lsum[P.lay.list(mk)];
        if (ms) keep = P.lay.ov(k) == (u32)ms;
        else if (VOTED) keep = (P.kept[slot] & 2) != 0;
        else {
          long long mi = find_edge(P, mk);
          keep = mi < 0 ? true : (P.kept[mi] != 0);  // no mirror: nobody can ask for the removal
        }
      }
    }
    // survivors are staged per warp in shared memory and appended back to back once 32 are waiting
    // (one same-address atomic per warp and round cost 2.2 ms); the export sorts them anyway
    const u32 bal = __ballot_sync(0xffffffffu, keep);
    if (bal) {
      if (keep) s_out[warp][wn + __popc(bal & lt_mask)] = k;
      wn += __popc(bal);
      __syncwarp();
      if (wn >= 32) {
        ull base = 0;
        if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_B], (ull)wn);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (u32 t = lane; t < wn; t += 32) out[base + t] = s_out[warp][t];
        wn = 0;
        __syncwarp();
      }
    }
  }
  if (wn) {
    ull base = 0;
    if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_B], (ull)wn);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (u32 t = lane; t < wn; t += 32) out[base + t] = s_out[warp][t];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    trans += __shfl_xor_sync(0xffffffffu, trans, o);
    contained += __shfl_xor_sync(0xffffffffu, contained, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (trans) atomicAdd(&P.dcount[DC_TRANS_EDGE], (ull)trans);
    if (contained) atomicAdd(&P.dcount[DC_CONTAINED_EDGE], (ull)contained);
  }
}

// Bucket path: TransitiveReduction + EdgeRemoval per LIST instead of per slot.  A consistent list keeps exactly its
// records of maximal overlap, whose slots consistency_kernel left in lseq, so the 146 M slots of config 2 are
// not streamed again: one thread per list (8.6 M), in slot order, looks at its one or two kept records and at
// the summary byte of their mirrors' lists.  What the list counts as transitive / contained was taken by
// consistency_kernel.  Lists that took the literal path are scanned for their kept flags.
template <bool VOTED>
__global__ void __launch_bounds__(256) tr_final_lists_kernel(StringParams P, const uint4* __restrict__ list_order,
                                                             u32 n_order, u64* __restrict__ out) {
  u32 trans = 0, contained = 0;
  const int lane = threadIdx.x & 31;
  const u32 lt_mask = (1u << lane) - 1;
  auto mirror_kept = [&](u64 k, size_t slot) {
    const u64 mk = P.lay.mirror(k);
    const uint8_t ms = P.lsum[P.lay.list(mk)];
    if (ms) return P.lay.ov(k) == (u32)ms;
    if (VOTED) return (P.kept[slot] & 2) != 0;
    const long long mi = find_edge(P, mk);
    return mi < 0 ? true : (P.kept[mi] != 0);  // no mirror: nobody can ask for the removal
  };
  auto append = [&](bool have, u64 k) {  // warp-aggregated append of the survivors
    const u32 bal = __ballot_sync(0xffffffffu, have);
    if (bal == 0) return;
    ull base = 0;
    if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_B], (ull)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (have) out[base + __popc(bal & lt_mask)] = k;
  };
  // block-uniform trip count: the appends are warp-collective
  for (u32 it0 = blockIdx.x * blockDim.x; it0 < n_order; it0 += gridDim.x * blockDim.x) {
    const u32 it = it0 + threadIdx.x;
    u64 k0 = ~0ull, k1 = ~0ull;
    bool keep0 = false, keep1 = false;
    bool scan = false;  // the list is walked slot by slot (more than two maximal records, or the literal path)
    uint4 q = make_uint4(0, 0, 0, 0), sq = make_uint4(0, 0, 0, 0);
    u32 om = 0;
    if (it < n_order) {
      q = list_order[it];
      sq = P.lseq[it];
      om = P.lsum[q.x];
      if (om) {
        trans += sq.w;
        contained += sq.z >> 16;
        const u32 nmax = (sq.z >> 8) & 0xffu;
        if (nmax <= 2) {
          k0 = P.edges[sq.x];
          keep0 = mirror_kept(k0, sq.x);
          if (nmax == 2) {
            k1 = P.edges[sq.y];
            keep1 = mirror_kept(k1, sq.y);
          }
        } else {
          scan = true;
        }
      } else if (sq.z & 0xffu) {
        scan = true;  // records, but not on the fast path: literal kernel
      }
    }
    append(keep0, k0);
    append(keep1, k1);
    // rare: lists walked slot by slot; the warp stays together for the collective appends
    u32 t = q.y;
    const u32 hi = q.y + q.z;
    while (__any_sync(0xffffffffu, scan)) {
      bool have = false;
      u64 k = ~0ull;
      while (scan && !have) {
        if (t >= hi) { scan = false; break; }
        k = P.edges[t];
        if (k != ~0ull) {
          const bool kept_here = om ? P.lay.ov(k) == om : (P.kept[t] & 1) != 0;
          have = kept_here && mirror_kept(k, t);
        }
        t++;
      }
      append(have, k);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    trans += __shfl_xor_sync(0xffffffffu, trans, o);
    contained += __shfl_xor_sync(0xffffffffu, contained, o);
  }
  if (lane == 0) {
    if (trans) atomicAdd(&P.dcount[DC_TRANS_EDGE], (ull)trans);
    if (contained) atomicAdd(&P.dcount[DC_CONTAINED_EDGE], (ull)contained);
  }
}

// ---- multi-GPU helpers: act on routed mirror keys at the rank that owns their list ----
__global__ void __launch_bounds__(256) apply_removal_requests_kernel(StringParams P, const u64* __restrict__ keys,
                                                                     size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long mi = find_edge(P, keys[i]);
  if (mi >= 0) { P.removed[mi] = 1; mark_dirty(P, P.lay.list(keys[i])); }
}

// several GPUs: every record the literal kernel kept (inconsistent lists only, a handful) votes
// "kept" for its mirror at the rank that owns the mirror's list.  One warp per queued list.
__global__ void __launch_bounds__(256) tr_votes_kernel(StringParams P) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 wi = warp_global; wi < P.n_work; wi += nwarps) {
    const u32 list = P.work[wi];
    const uint2 lr_ = P.adj[list]; const u32 lo = lr_.x, hi = lr_.x + lr_.y;
    for (u32 t = lo + lane; t < hi; t += 32)
      if (P.kept[t] & 1) {
        u32 ri = atomicAdd(P.n_req, 1u);
        if (ri < P.req_cap) P.req[ri] = P.lay.mirror(P.edges[t]);
      }
  }
}

__global__ void __launch_bounds__(256) apply_votes_kernel(StringParams P, const u64* __restrict__ keys, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long mi = find_edge(P, keys[i]);
  if (mi >= 0) P.kept[mi] |= 2;  // one vote per record at most: no other thread writes this byte
}

template <int NW>
__device__ __forceinline__ u32 list_owner_s(const u64* __restrict__ node_seq, u32 lid, int L, int K, int pbit,
                                            u32 wmask) {
  u64 r[NW];
  load_read_row<NW>(node_seq, lid >> 1, r);
  u64 km = extract_kmer<NW>(r, (lid & 1u) ? 0 : L - K, K);
  u64 rk = revcomp_kmer(km, K);
  u64 canon = km < rk ? km : rk;
  return (u32)(canon >> pbit) & wmask;
}

template <int NW>
__global__ void __launch_bounds__(256) key_owner_s_kernel(const u64* __restrict__ keys, size_t n, EdgeLayout lay,
                                                          const u64* __restrict__ node_seq, int L, int K, int pbit,
                                                          u32 wmask, u32* __restrict__ dest) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dest[i] = list_owner_s<NW>(node_seq, lay.list(keys[i]), L, K, pbit, wmask);
}

static StringParams make_params(cb_ctx* c, const u64* keys) {
  StringParams P;
  P.edges = keys;
  P.adj = c->adj.p;
  P.n_lists = 2 * c->n_nodes;
  P.slots = c->slots;
  P.lay = c->lay;
  P.node_seq = c->node_seq.p;
  P.ovh = c->ovh.p;
  P.nwo = c->NWo;
  P.node_idkey = c->node_idkey.p;
  P.node_cov = c->node_cov.p;
  P.L = c->L;
  P.K = c->K;
  P.majority = c->cfg.majority;
  P.pwm_n = c->cfg.pwm_n;
  P.removed = nullptr;
  P.kept = nullptr;
  P.list_flag = c->list_flag.p;
  P.lsumm = nullptr;
  P.lseq = nullptr;
  P.lsum = nullptr;
  P.work = nullptr;
  P.n_work = 0;
  P.work_out = nullptr;
  P.n_work_out = nullptr;
  P.check_consistency = 0;
  P.dirty_bit = 2;
  P.req = nullptr;
  P.n_req = nullptr;
  P.req_cap = 0;
  P.dcount = (ull*)c->dcount.p;
  return P;
}

static int warp_grid(size_t n_items) {
  int blocks = div_up(n_items, ST_WARPS);
  int cap = kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  return blocks > 0 ? blocks : 1;
}

static int slot_grid(size_t slots) {
  int blocks = div_up(slots, 256);
  int cap = kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  return blocks > 0 ? blocks : 1;
}

static u32 read_u32s(cb_ctx* c, const u32* p) {
  u32 v = 0;
  c->lc.rb.get(&v, p, 4, c->stream);
  c->lc.rb.sync(c->stream);
  return v;
}

// routes device keys to the owner rank of their adjacency list; returns the received keys
static size_t route_keys(cb_ctx* c, const u64* keys, size_t n, DevBuf<char>& out, u64* piggy = nullptr,
                         int n_piggy = 0) {
  cudaStream_t s = c->stream;
  DevBuf<u32> dest;
  guarded_defer(c, [&] { dest.alloc(n ? n : 1, s); });
  if (failed_locally(c)) n = 0;
  if (n) {
    CB_DISPATCH_NW(c->NW, (key_owner_s_kernel<NW_><<<div_up(n, 256), 256, 0, s>>>(
                              keys, n, c->lay, c->node_seq.p, c->L, c->K, c->pbit, (u32)c->cfg.world - 1, dest.p)));
    c->lc.n++;
  }
  size_t n_out = 0;
  route_records(c, keys, 8, dest.p, n, out, n_out, piggy, n_piggy);
  return n_out;
}

// a private copy of the slot array with everything CutChimericLinks removed put back (01-overlap)
void restore_overlap_copy(cb_ctx* c, DevBuf<u64>& out) {
  cudaStream_t s = c->stream;
  const size_t slots = c->slots ? c->slots : 1;
  out.alloc(slots, s);
  CB_CUDA(cudaMemcpyAsync(out.p, c->ckA.keys.p, slots * sizeof(u64), cudaMemcpyDeviceToDevice, s));
  if (c->n_cut_log) {
    restore_log_kernel<<<div_up(c->n_cut_log, 256), 256, 0, s>>>(out.p, c->cut_log.p, c->n_cut_log);
    c->lc.n++;
  }
}

void stage_string_graph(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const u32 n_lists = 2 * c->n_nodes;
  const size_t nl = n_lists ? n_lists : 1;
  const size_t slots = c->slots ? c->slots : 1;
  const bool dist = distributed(c);
  c->carried_code = 0;
  wait_node_table(c, true);  // id keys (CutChimericLinks tie-breaks) and SFA lines of every rank's nodes
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGE_REMOVAL, 0, (DC_COUNT - DC_EDGE_REMOVAL) * sizeof(u64), s));
  c->list_flag.alloc((nl + 3) & ~(size_t)3, s);
  CB_CUDA(cudaMemsetAsync(c->list_flag.p, 1, (nl + 3) & ~(size_t)3, s));
  DevBuf<u32> workA, workB, nwork;  // work lists: inconsistent lists / lists that lost an entry
  workA.alloc(nl, s);
  workB.alloc(nl, s);
  nwork.alloc(8, s);
  CB_CUDA(cudaMemsetAsync(nwork.p, 0, 32, s));
  DevBuf<uint8_t> flags;  // per slot: "removed" during the cut rounds, then "kept" for TR
  flags.alloc(slots, s);
  CB_CUDA(cudaMemsetAsync(flags.p, 0, slots, s));
  DevBuf<uint8_t> lsum;   // per list: summary byte of the final pass
  DevBuf<uint4> lsumm;    // per list: maximal overlap, its first two records and their number
  lsum.alloc(nl, s);
  const bool by_list = c->n_list_order > 0;  // bucket path: per-list summaries in slot order, TR per list
  DevBuf<uint4> lseq;
  if (by_list) lseq.alloc(c->n_list_order, s);
  else lsumm.alloc(nl, s);
  DevBuf<u64> req;  // multi-GPU: mirror keys to act on at their owner
  if (dist) req.alloc(slots, s);

  // ---- CutChimericLinks x <= 2 (BrushAssembler.java:347-367) ----
  timer_start(c, "cut");
  // a previous run of this stage group removed records in place: put them back first
  if (c->n_cut_log) {
    restore_log_kernel<<<div_up(c->n_cut_log, 256), 256, 0, s>>>(c->ckA.keys.p, c->cut_log.p, c->n_cut_log);
    c->lc.n++;
    c->n_cut_log = 0;
  }
  u64* const cur_keys = c->ckA.keys.p;  // 01-overlap, then (in place) 01-overlap.chl2
  size_t cur_n = c->ckA.n;
  u32 log_cap = 0;
  int64_t e1 = 0, e2 = 0;
  {
    StringParams P = make_params(c, cur_keys);
    P.lsumm = lsumm.p;
    P.lseq = lseq.p;
    if (c->ckA.n) {
      const u32* lists = nullptr;
      const uint4* lists4 = nullptr;
      u32 n_iter = n_lists;
      if (c->n_list_order) {  // bucket path: the lists that own slots, in slot order (coalesced list reads)
        lists4 = c->list_order.p;
        n_iter = (u32)c->n_list_order;
      } else if (dist) {  // workB is free until the first cut round writes its dirty list
        own_lists_kernel<<<div_up((size_t)nl, 256), 256, 0, s>>>(c->adj.p, n_lists, workB.p, nwork.p + 6);
        c->lc.n++;
        n_iter = read_u32s(c, nwork.p + 6);
        lists = workB.p;
      }
      c->lc.begin("consistency");
      if (n_iter)
        CB_DISPATCH_NW(c->NW, (consistency_kernel<NW_><<<min(div_up((size_t)n_iter, 256), kNumSMs * 16), 256, 0, s>>>(
                                  P, workA.p, nwork.p + 0, c->cfg.majority < 1.0f ? 0 : 1, lists, lists4, n_iter)));
      c->lc.end("consistency");
      c->lc.n++;
      CB_CUDA(cudaGetLastError());
    }
    u32 n_bad = read_u32s(c, nwork.p + 0);
    const u32* work = workA.p;
    u32 n_work = n_bad;
    for (int round = 1; round <= 2; round++) {
      u64 any_work = n_work;
      // every rank must take the same branches (collectives below); a cut-log failure of round 1 surfaces here too
      if (dist) any_work = ex_gather_checked(c, &any_work, 1).sum(0);
      if (any_work == 0) break;
      u32* dirty = (round == 1) ? workB.p : workA.p;  // round 2 reuses A's storage: its list is consumed
      u32* n_dirty = nwork.p + round;
      P = make_params(c, cur_keys);
      P.removed = flags.p;
      P.work = work;
      P.n_work = n_work;
      P.work_out = dirty;
      P.n_work_out = n_dirty;
      P.check_consistency = round == 2;
      P.dirty_bit = round == 1 ? 2 : 4;
      if (dist) {
        P.req = req.p;
        P.n_req = nwork.p + 4;
        P.req_cap = (u32)(slots > 0xffffffffu ? 0xffffffffu : slots);
        CB_CUDA(cudaMemsetAsync(nwork.p + 4, 0, 4, s));
      }
      CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGE_REMOVAL, 0, sizeof(u64), s));
      if (n_work) {
        c->lc.begin("cut");
        CB_DISPATCH_NW(c->NW, (cut_kernel<NW_><<<warp_grid(n_work), ST_WARPS * 32, 0, s>>>(P)));
        c->lc.end("cut");
        c->lc.n++;
        CB_CUDA(cudaGetLastError());
      }
      size_t n_req_applied = 0;
      u64 cuts = 0;
      u32 nd = 0;
      if (dist) {  // removal requests for mirrors -> owner rank of the mirror's list
        u32 nr = 0;
        c->lc.rb.get(&nr, nwork.p + 4, 4, s);
        download_counters(c);  // (the cut kernel is the only one that counts DC_EDGE_REMOVAL)
        if (nr > P.req_cap) {
          defer_error(c, CB_E_CAPACITY, "removal request buffer overflow");
          nr = P.req_cap;
        }
        // the exchange's own gather also sums the cuts of the round and carries the failure flags
        cuts = c->hcount[DC_EDGE_REMOVAL];
        DevBuf<char> got;
        size_t n_got = route_keys(c, req.p, nr, got, &cuts, 1);
        n_req_applied = n_got;
        if (n_got) {
          apply_removal_requests_kernel<<<div_up(n_got, 256), 256, 0, s>>>(P, (const u64*)got.p, n_got);
          c->lc.n++;
          CB_CUDA(cudaGetLastError());
        }
      } else {
        c->lc.rb.get(&nd, n_dirty, 4, s);  // (one round trip with the counters)
        download_counters(c);
        cuts = c->hcount[DC_EDGE_REMOVAL];
      }
      (round == 1 ? e1 : e2) = (int64_t)cuts;
      if (cuts == 0) break;
      // EdgeRemoval: the flagged slots of the lists that lost a record are emptied in place and logged
      if (dist) nd = read_u32s(c, n_dirty);  // (requests applied here have added to the list)
      {  // room for this round's removals: every local cut flags its record, plus the mirrors flagged here
        const size_t add = dist ? (size_t)c->hcount[DC_EDGE_REMOVAL] + n_req_applied : 2 * (size_t)c->hcount[DC_EDGE_REMOVAL];
        const size_t need = c->n_cut_log + add;
        if (c->cut_log.n < 2 * need) {
          DevBuf<u64> bigger;
          bigger.alloc(2 * need + 64, s);
          if (c->n_cut_log)
            CB_CUDA(cudaMemcpyAsync(bigger.p, c->cut_log.p, 2 * c->n_cut_log * sizeof(u64), cudaMemcpyDeviceToDevice, s));
          c->cut_log = std::move(bigger);
        }
        log_cap = (u32)(c->cut_log.n / 2);
      }
      CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGES_C, 0, sizeof(u64), s));
      if (nd) {
        remove_flagged_kernel<<<warp_grid(nd), ST_WARPS * 32, 0, s>>>(cur_keys, flags.p, c->adj.p, dirty, nd,
                                                                     c->cut_log.p, nwork.p + 7, log_cap,
                                                                     (ull*)c->dcount.p);
        c->lc.n++;
        CB_CUDA(cudaGetLastError());
      }
      u32 n_log = 0;
      c->lc.rb.get(&n_log, nwork.p + 7, 4, s);
      download_counters(c);
      cur_n -= (size_t)c->hcount[DC_EDGES_C];
      c->n_cut_log = n_log;
      if (c->n_cut_log > log_cap) {  // (cannot happen: the log was sized from the exact counts above)
        c->n_cut_log = log_cap;
        c->overlapped = false;  // the slot array no longer restores to 01-overlap: cb_build_overlap must run again
        defer_error(c, CB_E_CAPACITY, "cut log overflow");
      }
      // round 2 only has to look at the lists that changed: an unchanged list cuts nothing again
      work = dirty;
      n_work = nd;
    }
  }
  c->counters["edge_removal_1"] = e1;
  c->counters["edge_removal_2"] = e2;
  c->counters["edge_removal"] = e1 > 0 ? e2 : e1;
  timer_stop(c, "cut");

  // ---- TransitiveReduction + EdgeRemoval (BrushAssembler.java:370-379) ----
  timer_start(c, "tr");
  c->ckB.keys.alloc(cur_n ? cur_n : 1, s);  // compact: the survivors back to back (no more than what is left)
  c->ckB.compact = true;
  {
    // `flags` is all zero again here (remove_flagged_kernel clears what it consumed)
    StringParams P = make_params(c, cur_keys);
    P.kept = flags.p;
    P.lsumm = lsumm.p;
    P.lseq = lseq.p;
    P.lsum = lsum.p;
    // literal path first: lists that are inconsistent or lost an entry in a cut round
    CB_CUDA(cudaMemsetAsync(nwork.p + 3, 0, 4, s));
    if (by_list) {
      CB_CUDA(cudaMemsetAsync(lsum.p, 0, nl, s));
      tr_worklist_seq_kernel<<<div_up(c->n_list_order, 256), 256, 0, s>>>(c->list_flag.p, c->list_order.p, lseq.p,
                                                                         (u32)c->n_list_order, workA.p, nwork.p + 3, lsum.p);
    } else {
      tr_worklist_kernel<<<div_up((size_t)nl, 256), 256, 0, s>>>(c->list_flag.p, c->adj.p, lsumm.p, n_lists,
                                                                workA.p, nwork.p + 3, lsum.p);
    }
    c->lc.n++;
    u32 n_slow = read_u32s(c, nwork.p + 3);
    if (n_slow) {
      P.work = workA.p;
      P.n_work = n_slow;
      c->lc.begin("tr");
      CB_DISPATCH_NW(c->NW, (tr_kernel<NW_><<<warp_grid(n_slow), ST_WARPS * 32, 0, s>>>(P)));
      c->lc.end("tr");
      c->lc.n++;
    }
    CB_CUDA(cudaGetLastError());
    if (dist) {
      // every rank needs the summary byte of every list (each list has one owner; the others hold 0)
      timer_start(c, "tr_votes");
      ex_allreduce_max_u8(c, lsum.p, nl);
      // mirrors that live in an inconsistent list: its owner votes for the records it kept
      P.req = req.p;
      P.n_req = nwork.p + 5;
      P.req_cap = (u32)(slots > 0xffffffffu ? 0xffffffffu : slots);
      CB_CUDA(cudaMemsetAsync(nwork.p + 5, 0, 4, s));
      if (n_slow) {
        tr_votes_kernel<<<warp_grid(n_slow), ST_WARPS * 32, 0, s>>>(P);
        c->lc.n++;
      }
      u32 nv = read_u32s(c, nwork.p + 5);
      if (nv > P.req_cap) {
        defer_error(c, CB_E_CAPACITY, "vote buffer overflow");
        nv = P.req_cap;
      }
      u64 any_votes = nv;
      any_votes = ex_gather_checked(c, &any_votes, 1).sum(0);  // (and every failure deferred since the cut rounds)
      if (any_votes) {
        DevBuf<char> got;
        size_t n_got = route_keys(c, req.p, nv, got);
        if (n_got) {
          apply_votes_kernel<<<div_up(n_got, 256), 256, 0, s>>>(P, (const u64*)got.p, n_got);
          c->lc.n++;
        }
      }
      timer_stop(c, "tr_votes");
    }
    if (cur_n) {
      c->lc.begin("tr_final");
      if (by_list) {
        const int grid = slot_grid(c->n_list_order);
        if (dist) tr_final_lists_kernel<true><<<grid, 256, 0, s>>>(P, c->list_order.p, (u32)c->n_list_order, c->ckB.keys.p);
        else tr_final_lists_kernel<false><<<grid, 256, 0, s>>>(P, c->list_order.p, (u32)c->n_list_order, c->ckB.keys.p);
      } else if (dist) tr_final_kernel<true><<<slot_grid(slots), 256, 0, s>>>(P, c->ckB.keys.p);
      else tr_final_kernel<false><<<slot_grid(slots), 256, 0, s>>>(P, c->ckB.keys.p);
      c->lc.end("tr_final");
      c->lc.n++;
    }
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "tr");
  download_counters(c);
  {
    const u64 left = cur_n;
    const Gathered G = reduce_counters(c, &left, 1);
    c->counters["edges_chl2"] = (int64_t)G.sum(DC_COUNT);
  }
  c->ckB.n = (size_t)c->hcount[DC_EDGES_B];
  c->ckB.valid = true;
  c->n_edges_c = cur_n;
  if (c->gcount[DC_TR_OVERFLOW]) throw Error(CB_E_CAPACITY, "TransitiveReduction kept-list capacity exceeded");
  c->counters["trans_edge"] = (int64_t)c->gcount[DC_TRANS_EDGE];
  c->counters["contained_edge"] = (int64_t)c->gcount[DC_CONTAINED_EDGE];
  c->counters["edges_re"] = (int64_t)c->gcount[DC_EDGES_B];
  c->counters["slow_lists"] = (int64_t)c->gcount[DC_SLOW_LISTS];
  c->counters["host_collectives"] = c->n_host_coll;
}

}  // namespace cb
