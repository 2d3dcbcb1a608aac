// stage_string.cu -- per-node adjacency pruning: CutChimericLinks, EdgeRemoval and
// TransitiveReduction on the sorted edge set, one warp per (node, direction) list.
//
// Replaces (reference):
//   CutChimericLinksMapper/Reducer  CutChimericLinks.java:87-114, :258-409  cut_kernel
//   Node.Consensus                  Node.java:1293-1377                     cut_kernel (PWM part)
//   EdgeRemovalMapper/Reducer       EdgeRemoval.java:58-87, :121-202        removed flags + compaction
//   TransitiveReductionReducer      TransitiveReduction.java:239-547        tr_kernel
//   driver loop                     BrushAssembler.java:337-379             stage_string_graph
//
// The graph of checkpoint A is symmetric (S u mirror(S)), so the neighbour messages each
// reducer rebuilds its lists from (DARKMSG / OVALMSG, sent by the mirror edge) are exactly the
// node's own adjacency list; a list is the contiguous key range [adj_off[2a+x], adj_off[2a+x+1])
// ordered by overlap descending.  The overhang of an entry is b^y[ov:], the bases of the
// neighbour that extend past the node.
//
// Both kernels first test whether every overhang of the list is a prefix of the longest one
// ("consistent list").  Then (a) the PWM consensus equals that longest overhang wherever it is
// defined, has no N, and cuts nothing (needs MAJORITY < 1); (b) the greedy scan of
// TransitiveReduction keeps exactly the entries of maximal overlap.  Lists that are not
// consistent (sequencing errors, repeats) take the literal path.
#include "ctx.cuh"

namespace cb {

template <int NW>
__device__ __forceinline__ void load_read_c(const u64* __restrict__ reads, u32 idx, u64 (&r)[NW]) {
  const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    uint4 v = __ldg(row + i);
    r[2 * i] = (u64)v.x | ((u64)v.y << 32);
    r[2 * i + 1] = (u64)v.z | ((u64)v.w << 32);
  }
}

struct StringParams {
  const u64* edges;    // gapped CSR keys
  const u32* adj_off;  // [n_lists + 1]
  const u32* cnt;      // [n_lists]
  u32 n_lists;         // 2 * n_nodes
  EdgeLayout lay;
  const u64* node_seq;
  const IdKey* node_idkey;
  const u32* node_cov;
  int L, K;
  float majority, pwm_n;
  uint8_t* removed;    // cut_kernel output, per slot
  uint8_t* kept;       // tr_kernel output, per slot
  uint8_t* list_flag;  // per list: 1 consistent, 0 not, 2 unknown
  ull* dcount;
};

// overhang b^y[ov:], left aligned, of an adjacency record
template <int NW>
__device__ __forceinline__ void load_overhang(const StringParams& P, u64 key, u64 (&h)[NW], int& len) {
  u64 r[NW];
  load_read_c<NW>(P.node_seq, P.lay.nbr(key), r);
  if (P.lay.y(key)) revcomp_read<NW>(r, h, P.L);
  else {
#pragma unroll
    for (int w = 0; w < NW; w++) h[w] = r[w];
  }
  int ov = (int)P.lay.ov(key);
  shl_bases<NW>(h, ov);
  len = P.L - ov;
}

template <int NW>
__device__ __forceinline__ int base_at(const u64 (&h)[NW], int j) {
  u64 w = 0;
#pragma unroll
  for (int i = 0; i < NW; i++)
    if (i == (j >> 5)) w = h[i];
  return (int)((w >> (62 - 2 * (j & 31))) & 3);
}

// slot index of `key` in its list or -1.  Lists are ordered by (L - ov) ascending; entries of
// equal overlap are in no particular order, so the equal-overlap run is scanned.
__device__ __forceinline__ long long find_edge(const StringParams& P, u64 key) {
  const u32 list = P.lay.list(key);
  const u32 base = P.adj_off[list], n = P.cnt[list];
  const u32 sh = P.lay.nb + 1, msk = (1u << P.lay.ob) - 1;
  const u32 target = (u32)(key >> sh) & msk;
  u32 lo = 0, hi = n;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (((u32)(P.edges[base + mid] >> sh) & msk) < target) lo = mid + 1;
    else hi = mid;
  }
  for (u32 t = lo; t < n; t++) {
    u64 k = P.edges[base + t];
    if (((u32)(k >> sh) & msk) != target) break;
    if (k == key) return (long long)(base + t);
  }
  return -1;
}

// are all overhangs of [lo, hi) prefixes of the longest one (the last entry)?
template <int NW>
__device__ __forceinline__ bool list_consistent(const StringParams& P, u32 lo, u32 hi, int lane) {
  u64 H[NW];
  int lenH;
  load_overhang<NW>(P, P.edges[hi - 1], H, lenH);
  bool good = true;
  for (u32 t0 = lo; t0 < hi; t0 += 32) {
    u32 t = t0 + lane;
    if (t < hi) {
      u64 h[NW];
      int len;
      load_overhang<NW>(P, P.edges[t], h, len);
      good = good && prefix_equal<NW>(h, H, len);
    }
  }
  return __all_sync(0xffffffffu, good);
}

constexpr int ST_WARPS = 8;

// ---------------------------------------------------------------------------------------------
// K10: CutChimericLinks
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

template <int NW>
__global__ void __launch_bounds__(ST_WARPS * 32) cut_kernel(StringParams P) {
  __shared__ unsigned char s_cons[ST_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L;
  for (u32 list = warp_global; list < P.n_lists; list += nwarps) {
    const u32 n = P.cnt[list];
    const u32 lo = P.adj_off[list], hi = lo + n;
    if (n <= 1) {  // CutChimericLinks.java:306 "flist.size() > 1"
      if (lane == 0) P.list_flag[list] = 1;
      continue;
    }
    const bool cons = list_consistent<NW>(P, lo, hi, lane);
    if (lane == 0) P.list_flag[list] = cons ? 1 : 0;
    if (cons && P.majority < 1.0f) continue;
    if (lane == 0) atomicAdd(&P.dcount[DC_SLOW_LISTS], 1ull);
    const u32 x = list & 1;
    // ---- entries 0 and 1 of the Consensus order (Node.java:1295-1320) ----
    TopKey best[2];
    for (int r = 0; r < 2; r++) {
      TopKey mine;
      mine.ov = 0xffffffffu; mine.hi = ~0ull; mine.lo = ~0ull; mine.idx = 0xffffffffu;
      for (u32 t = lo + lane; t < hi; t += 32) {
        if (r == 1 && t == best[0].idx) continue;
        u64 k = P.edges[t];
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
    if (n == 2 || (float)cov0 + (float)cov1 > 2.0f) clen = L - (int)P.lay.ov(P.edges[hi - 2]);
    else clen = L - (int)P.lay.ov(P.edges[hi - 3]);
    // ---- position weight matrix, one lane per position (Node.java:1321-1364) ----
    int ncount = 0;
    for (int j0 = 0; j0 < clen; j0 += 32) {
      int j = j0 + lane;
      int cnt[4] = {0, 0, 0, 0};
      int sum = 0;
      for (u32 t = lo; t < hi; t++) {
        u64 k = P.edges[t];
        u64 h[NW];
        int len;
        load_overhang<NW>(P, k, h, len);
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
      if (t < hi) {
        u64 k = P.edges[t];
        u64 h[NW];
        int len;
        load_overhang<NW>(P, k, h, len);
        int m = len < clen ? len : clen;
        bool miss = false;
        for (int j = 0; j < m; j++) {
          int cc = s_cons[warp][j];
          if (cc != 4 && cc != base_at<NW>(h, j)) { miss = true; break; }
        }
        if (miss) {
          P.removed[t] = 1;
          long long mi = find_edge(P, P.lay.mirror(k));
          if (mi >= 0) P.removed[mi] = 1;
          atomicAdd(&P.dcount[DC_EDGE_REMOVAL], 1ull);
        }
      }
    }
    __syncwarp();
  }
}

// EdgeRemoval (EdgeRemoval.java:121-202): per list, copy the entries that are not flagged.
// A list that loses an entry has its consistency flag reset to "unknown".
__global__ void __launch_bounds__(ST_WARPS * 32) filter_lists_kernel(const u64* __restrict__ in,
                                                                     const u32* __restrict__ adj_off,
                                                                     const u32* __restrict__ cnt_in,
                                                                     const uint8_t* __restrict__ drop, u32 n_lists,
                                                                     u64* __restrict__ out, u32* __restrict__ cnt_out,
                                                                     uint8_t* __restrict__ list_flag,
                                                                     ull* __restrict__ dcount, int counter) {
  const int lane = threadIdx.x & 31;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  ull total = 0;
  for (u32 list = warp_global; list < n_lists; list += nwarps) {
    const u32 n = cnt_in[list], lo = adj_off[list];
    u32 w = 0;
    for (u32 t0 = 0; t0 < n; t0 += 32) {
      u32 t = t0 + lane;
      bool keep = t < n && !drop[lo + t];
      u64 k = keep ? in[lo + t] : 0;
      u32 bal = __ballot_sync(0xffffffffu, keep);
      if (keep) out[lo + w + __popc(bal & lt_mask)] = k;
      w += __popc(bal);
    }
    if (lane == 0) {
      cnt_out[list] = w;
      if (w != n && list_flag) list_flag[list] = 2;
      total += w;
    }
  }
  if (lane == 0 && total) atomicAdd(&dcount[counter], total);
}

// ---------------------------------------------------------------------------------------------
// K11: TransitiveReduction
// ---------------------------------------------------------------------------------------------
constexpr int TR_KC = 64;  // kept entries per list in the literal scan

template <int NW>
__global__ void __launch_bounds__(ST_WARPS * 32) tr_kernel(StringParams P) {
  __shared__ u64 s_kh[ST_WARPS][TR_KC][NW];
  __shared__ u64 s_kkey[ST_WARPS][TR_KC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L;
  u32 trans = 0, contained = 0;
  for (u32 list = warp_global; list < P.n_lists; list += nwarps) {
    const u32 n = P.cnt[list];
    if (n == 0) continue;
    const u32 lo = P.adj_off[list], hi = lo + n;
    const uint8_t fl = P.list_flag[list];
    const bool cons = fl == 1 ? true : list_consistent<NW>(P, lo, hi, lane);
    if (cons) {
      // all overhangs nest: the scan keeps the entries of maximal overlap and drops the rest
      const u32 ovmax = P.lay.ov(P.edges[lo]);
      for (u32 t0 = lo; t0 < hi; t0 += 32) {
        u32 t = t0 + lane;
        if (t < hi) {
          u64 k = P.edges[t];
          u32 ov = P.lay.ov(k);
          if ((int)ov == L) contained++;  // TransitiveReduction.java:319-323 (equal-length reads)
          if (ov == ovmax) P.kept[t] = 1;
          else {
            // :325-334 an entry of the same type to the same neighbour already kept -> dropped
            // without counting as transitive
            bool step2 = false;
            for (u32 q = lo; q < hi && P.lay.ov(P.edges[q]) == ovmax; q++) {
              u64 kq = P.edges[q];
              if (P.lay.nbr(kq) == P.lay.nbr(k) && P.lay.y(kq) == P.lay.y(k)) { step2 = true; break; }
            }
            if (!step2) trans++;
          }
        }
      }
      continue;
    }
    if (lane == 0) atomicAdd(&P.dcount[DC_SLOW_LISTS], 1ull);
    // ---- literal greedy scan (TransitiveReduction.java:300-409) ----
    u32 nk = 0;
    for (u32 t = lo; t < hi; t++) {
      u64 k = P.edges[t];
      u64 h[NW];
      int len;
      load_overhang<NW>(P, k, h, len);
      const u32 ov = P.lay.ov(k), y = P.lay.y(k), nbr = P.lay.nbr(k);
      if ((int)ov == L && lane == 0) contained++;
      bool step2 = false, step3 = false;
      for (u32 q = lane; q < nk; q += 32) {
        u64 kq = s_kkey[warp][q];
        if (P.lay.nbr(kq) == nbr && P.lay.y(kq) == y) step2 = true;
        u32 ovq = P.lay.ov(kq);
        if (ovq != ov) {  // :341-343 (equal length reads: same ov <=> same ov and same length)
          u64 hq[NW];
#pragma unroll
          for (int w = 0; w < NW; w++) hq[w] = s_kh[warp][q][w];
          if (prefix_equal<NW>(h, hq, L - (int)ovq)) step3 = true;  // prefix.startsWith(suffix) :345
        }
      }
      step2 = __any_sync(0xffffffffu, step2);
      step3 = __any_sync(0xffffffffu, step3);
      if (step2) continue;
      if (step3) { if (lane == 0) trans++; continue; }
      if (nk >= TR_KC) { if (lane == 0) atomicAdd(&P.dcount[DC_TR_OVERFLOW], 1ull); continue; }
      if (lane < NW) {
#pragma unroll
        for (int w = 0; w < NW; w++)
          if (w == lane) s_kh[warp][nk][w] = h[w];
      }
      if (lane == 0) { s_kkey[warp][nk] = k; P.kept[t] = 1; }
      nk++;
      __syncwarp();
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

// EdgeRemoval after TransitiveReduction: an edge survives iff both endpoints kept it.
// Writes the surviving entries of every list (checkpoint B) at the list's own offset.
__global__ void __launch_bounds__(ST_WARPS * 32) tr_final_kernel(StringParams P, u64* __restrict__ out,
                                                                 u32* __restrict__ cnt_out) {
  const int lane = threadIdx.x & 31;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  ull total = 0;
  for (u32 list = warp_global; list < P.n_lists; list += nwarps) {
    const u32 n = P.cnt[list], lo = P.adj_off[list];
    u32 w = 0;
    for (u32 t0 = 0; t0 < n; t0 += 32) {
      u32 t = t0 + lane;
      bool keep = false;
      u64 k = 0;
      if (t < n && P.kept[lo + t]) {
        k = P.edges[lo + t];
        long long mi = find_edge(P, P.lay.mirror(k));
        keep = (mi >= 0) ? (P.kept[mi] != 0) : true;  // no mirror: nobody can ask for the removal
      }
      u32 bal = __ballot_sync(0xffffffffu, keep);
      if (keep) out[lo + w + __popc(bal & lt_mask)] = k;
      w += __popc(bal);
    }
    if (lane == 0) { cnt_out[list] = w; total += w; }
  }
  if (lane == 0 && total) atomicAdd(&P.dcount[DC_EDGES_B], total);
}

#define CB_DISPATCH_NW(nw, CALL)                                             \
  switch (nw) {                                                              \
    case 2: { constexpr int NW_ = 2; CALL; } break;                          \
    case 4: { constexpr int NW_ = 4; CALL; } break;                          \
    case 6: { constexpr int NW_ = 6; CALL; } break;                          \
    case 8: { constexpr int NW_ = 8; CALL; } break;                          \
    default: throw Error(CB_E_UNSUPPORTED, "read length above 256 bases");   \
  }

static StringParams make_params(cb_ctx* c, const EdgeSet& es) {
  StringParams P;
  P.edges = es.keys.p;
  P.adj_off = c->adj_off.p;
  P.cnt = es.cnt.p;
  P.n_lists = 2 * c->n_nodes;
  P.lay = c->lay;
  P.node_seq = c->node_seq.p;
  P.node_idkey = c->node_idkey.p;
  P.node_cov = c->node_cov.p;
  P.L = c->L;
  P.K = c->K;
  P.majority = c->cfg.majority;
  P.pwm_n = c->cfg.pwm_n;
  P.removed = nullptr;
  P.kept = nullptr;
  P.list_flag = c->list_flag.p;
  P.dcount = (ull*)c->dcount.p;
  return P;
}

static int list_grid(u32 n_lists) {
  int blocks = div_up((size_t)n_lists, ST_WARPS);
  int cap = kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  return blocks > 0 ? blocks : 1;
}

// one CutChimericLinks job + the EdgeRemoval that follows it; returns edge_removal
static int64_t cut_round(cb_ctx* c, const EdgeSet& in, EdgeSet& out, bool& changed) {
  cudaStream_t s = c->stream;
  changed = false;
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGE_REMOVAL, 0, sizeof(u64), s));
  if (in.n == 0) return 0;
  const u32 n_lists = 2 * c->n_nodes;
  DevBuf<uint8_t> removed;
  removed.alloc(c->slots, s);
  CB_CUDA(cudaMemsetAsync(removed.p, 0, c->slots, s));
  StringParams P = make_params(c, in);
  P.removed = removed.p;
  c->lc.begin("cut");
  CB_DISPATCH_NW(c->NW, (cut_kernel<NW_><<<list_grid(P.n_lists), ST_WARPS * 32, 0, s>>>(P)));
  c->lc.end("cut");
  c->lc.n++;
  CB_CUDA(cudaGetLastError());
  download_counters(c);
  int64_t cuts = (int64_t)c->hcount[DC_EDGE_REMOVAL];
  if (cuts > 0) {
    out.keys.alloc(c->slots, s);
    out.cnt.alloc(n_lists, s);
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGES_C, 0, sizeof(u64), s));
    filter_lists_kernel<<<list_grid(n_lists), ST_WARPS * 32, 0, s>>>(in.keys.p, c->adj_off.p, in.cnt.p, removed.p,
                                                                    n_lists, out.keys.p, out.cnt.p, c->list_flag.p,
                                                                    (ull*)c->dcount.p, (int)DC_EDGES_C);
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    download_counters(c);
    out.n = (size_t)c->hcount[DC_EDGES_C];
    out.valid = true;
    changed = true;
  }
  return cuts;
}

void stage_string_graph(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const u32 n_lists = 2 * c->n_nodes;
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_EDGE_REMOVAL, 0, (DC_COUNT - DC_EDGE_REMOVAL) * sizeof(u64), s));
  c->list_flag.alloc(n_lists ? n_lists : 1, s);
  CB_CUDA(cudaMemsetAsync(c->list_flag.p, 2, n_lists ? n_lists : 1, s));
  // ---- CutChimericLinks x <= 2 (BrushAssembler.java:347-367) ----
  timer_start(c, "cut");
  EdgeSet tmp1, tmp2;
  const EdgeSet* cur = &c->ckA;
  bool ch1 = false, ch2 = false;
  int64_t e1 = cut_round(c, *cur, tmp1, ch1), e2 = 0;
  if (ch1) cur = &tmp1;
  c->counters["edge_removal_1"] = e1;
  c->counters["edge_removal"] = e1;
  if (e1 > 0) {
    e2 = cut_round(c, *cur, tmp2, ch2);
    if (ch2) cur = &tmp2;
    c->counters["edge_removal"] = e2;
  }
  c->counters["edge_removal_2"] = e2;
  timer_stop(c, "cut");
  c->ckC_is_A = (cur == &c->ckA);
  if (cur == &tmp1) c->ckC = std::move(tmp1);
  else if (cur == &tmp2) c->ckC = std::move(tmp2);
  const EdgeSet& C = c->ckC_is_A ? c->ckA : c->ckC;
  c->counters["edges_chl2"] = (int64_t)C.n;

  // ---- TransitiveReduction + EdgeRemoval (BrushAssembler.java:370-379) ----
  timer_start(c, "tr");
  c->ckB.keys.alloc(c->slots ? c->slots : 1, s);
  c->ckB.cnt.alloc(n_lists ? n_lists : 1, s);
  if (C.n == 0) {
    CB_CUDA(cudaMemsetAsync(c->ckB.cnt.p, 0, (size_t)(n_lists ? n_lists : 1) * 4, s));
    c->ckB.n = 0;
  } else {
    DevBuf<uint8_t> kept;
    kept.alloc(c->slots, s);
    CB_CUDA(cudaMemsetAsync(kept.p, 0, c->slots, s));
    StringParams P = make_params(c, C);
    P.kept = kept.p;
    c->lc.begin("tr");
    CB_DISPATCH_NW(c->NW, (tr_kernel<NW_><<<list_grid(P.n_lists), ST_WARPS * 32, 0, s>>>(P)));
    c->lc.end("tr");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    c->lc.begin("tr_final");
    tr_final_kernel<<<list_grid(P.n_lists), ST_WARPS * 32, 0, s>>>(P, c->ckB.keys.p, c->ckB.cnt.p);
    c->lc.end("tr_final");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "tr");
  download_counters(c);
  c->ckB.n = (size_t)c->hcount[DC_EDGES_B];
  c->ckB.valid = true;
  if (c->hcount[DC_TR_OVERFLOW]) throw Error(CB_E_CAPACITY, "TransitiveReduction kept-list capacity exceeded");
  c->counters["trans_edge"] = (int64_t)c->hcount[DC_TRANS_EDGE];
  c->counters["contained_edge"] = (int64_t)c->hcount[DC_CONTAINED_EDGE];
  c->counters["edges_re"] = (int64_t)c->ckB.n;
  c->counters["slow_lists"] = (int64_t)c->hcount[DC_SLOW_LISTS];
}

}  // namespace cb
