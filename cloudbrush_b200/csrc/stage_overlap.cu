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
                                                     const u32* __restrict__ node_cov, u32 n_nodes, int L, int K,
                                                     int W, int pb, u64* __restrict__ keys, u32* __restrict__ vals,
                                                     ull* __restrict__ dcount) {
  const int lane = threadIdx.x & 31;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  u32 invalid = 0;
  for (u32 node = warp_global; node < n_nodes; node += nwarps) {
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
        size_t slot = (size_t)node * W + i;
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
// K6+K7+K8: segmented join fused with the packed-word overlap check.
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
  u64* edges_out;
  u64 edges_cap;
  ull* dcount;
};

constexpr int JN_WARPS = 8;
constexpr int JN_EB = 256;  // staged edge records per warp

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
  if ((pos == 0 || pos == L - K) && sel) ov = K;
}

template <int NW>
__global__ void __launch_bounds__(JN_WARPS * 32) join_verify_kernel(JoinParams P) {
  __shared__ u64 s_B[JN_WARPS][32][NW];
  __shared__ u32 s_Bmeta[JN_WARPS][32];
  __shared__ u64 s_edges[JN_WARPS][JN_EB];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 lt_mask = (1u << lane) - 1;
  const u32 warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  const int L = P.L, K = P.K;
  const u32 posmask = (1u << P.pb) - 1;
  u32 ebuf = 0;        // staged records (warp uniform)
  u32 n_verified = 0;  // lane-local count of maximal verified candidates

  auto flush = [&]() {
    if (ebuf == 0) return;
    ull base = 0;
    if (lane == 0) base = atomicAdd(&P.dcount[DC_EDGES_RAW], (ull)ebuf);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (u32 t = lane; t < ebuf; t += 32)
      if (base + t < P.edges_cap) P.edges_out[base + t] = s_edges[warp][t];
    __syncwarp();
    ebuf = 0;
  };

  const u32 nchunks = (P.T + 31) / 32;
  for (u32 chunk = warp_global; chunk < nchunks; chunk += nwarps) {
    u32 i = chunk * 32 + lane;
    bool valid = i < P.T;
    u64 ki = valid ? (P.keys[i] & P.kmask) : 0;
    u64 kprev = (valid && i > 0) ? (P.keys[i - 1] & P.kmask) : ~ki;
    u32 heads = __ballot_sync(0xffffffffu, valid && ki != kprev);
    while (heads) {
      int h = __ffs(heads) - 1;
      heads &= heads - 1;
      const u32 s = chunk * 32 + h;
      const u64 kseg = __shfl_sync(0xffffffffu, ki, h);
      u32 e = s + 1;
      for (;;) {
        u32 j = e + lane;
        bool diff = (j >= P.T) || ((P.keys[j < P.T ? j : 0] & P.kmask) != kseg);
        u32 b = __ballot_sync(0xffffffffu, diff);
        if (b) { e += __ffs(b) - 1; break; }
        e += 32;
      }
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
              // the same (a, x) with the same (b, y) wins
              for (int dirn = -1; dirn <= 1 && ok; dirn += 2) {
                long long jj = (long long)j + dirn;
                while (jj >= (long long)s && jj < (long long)e) {
                  u64 m2 = tuple_meta(P, (u32)jj);
                  if ((u32)(m2 >> (P.pb + 1)) != a) break;
                  int x2, ov2;
                  entry_of(L, K, (int)((m2 >> 1) & posmask), (int)(m2 & 1), sel, x2, ov2);
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
              if (ebuf + 2 * cnt > JN_EB) flush();
              if (ok) {
                u32 r = ebuf + 2 * __popc(bal & lt_mask);
                s_edges[warp][r] = P.lay.encode(a, (u32)x, (u32)ov, (u32)y, b);
                s_edges[warp][r + 1] = P.lay.encode(b, (u32)(y ^ 1), (u32)ov, (u32)(x ^ 1), a);  // GenReverseEdge
                n_verified++;
              }
              ebuf += 2 * cnt;
              __syncwarp();
            }
          }
        }
        __syncwarp();
        if (scan_pos >= e) break;
      }
    }
  }
  flush();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_verified += __shfl_xor_sync(0xffffffffu, n_verified, o);
  if (lane == 0 && n_verified) atomicAdd(&P.dcount[DC_CAND_VERIFIED], (ull)n_verified);
}

// ---------------------------------------------------------------------------------------------
// unique + compaction of a sorted key array
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unique_flags_kernel(const u64* __restrict__ k, size_t n,
                                                           u32* __restrict__ flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || k[i] != k[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) compact_kernel(const u64* __restrict__ k, const u32* __restrict__ flag,
                                                      const u32* __restrict__ idx, size_t n, u64* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[idx[i]] = k[i];
}

// keeps k[i] where flag[i] != 0; returns the number kept (synchronises the stream)
size_t compact_by_flags(cb_ctx* c, const u64* k, const u32* flag, size_t n, DevBuf<u64>& out) {
  cudaStream_t s = c->stream;
  if (n == 0) { out.alloc(0, s); return 0; }
  DevBuf<u32> idx;
  idx.alloc(n, s);
  exclusive_scan_u32(flag, idx.p, n, s, c->lc);
  u32 last_idx = 0, last_flag = 0;
  CB_CUDA(cudaMemcpyAsync(&last_idx, idx.p + n - 1, 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(&last_flag, flag + n - 1, 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  size_t m = (size_t)last_idx + (last_flag ? 1 : 0);
  out.alloc(m, s);
  compact_kernel<<<div_up(n, 256), 256, 0, s>>>(k, flag, idx.p, n, out.p);
  c->lc.n++;
  CB_CUDA(cudaGetLastError());
  return m;
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

void stage_build_overlap(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const u32 nn = c->n_nodes;
  const int L = c->L, K = c->K, W = c->W;
  const int pb = bits_for((u64)W);
  const int nbits = bits_for((u64)nn);
  if (nbits > 27) throw Error(CB_E_CAPACITY, "more than 2^27 nodes per context");
  if (nbits + pb + 1 > 32 + (64 - 2 * K))
    throw Error(CB_E_CAPACITY, "node/window payload does not fit the 12-byte tuple for this k");
  const size_t Tmax = (size_t)nn * W;
  if (Tmax >= (1ull << 32)) throw Error(CB_E_CAPACITY, "more than 2^32 k-mer tuples per context");
  c->lay.nb = nbits;
  c->lay.ob = bits_for((u64)(L - K + 1));
  c->lay.L = L;

  // ---- keygen ----
  timer_start(c, "keygen");
  DevBuf<u64> keys[2];
  DevBuf<u32> vals[2];
  for (int i = 0; i < 2; i++) { keys[i].alloc(Tmax, s); vals[i].alloc(Tmax, s); }
  CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_TUPLES_INVALID, 0, (DC_COUNT - DC_TUPLES_INVALID) * sizeof(u64), s));
  {
    int blocks = div_up((size_t)nn * 32, 256);
    int cap = kNumSMs * 8 * 4;
    if (blocks > cap) blocks = cap;
    c->lc.begin("keygen");
    CB_DISPATCH_NW(c->NW, (keygen_kernel<NW_><<<blocks, 256, 0, s>>>(c->node_seq.p, c->node_cov.p, nn, L, K, W, pb,
                                                                    keys[0].p, vals[0].p, (ull*)c->dcount.p)));
    c->lc.end("keygen");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  timer_stop(c, "keygen");

  // ---- shuffle replacement ----
  timer_start(c, "sort_tuples");
  int where = radix_sort(keys[0].p, keys[1].p, vals[0].p, vals[1].p, Tmax, 0, 2 * K, s, c->lc, "tuples");
  timer_stop(c, "sort_tuples");
  download_counters(c);
  const u32 T = (u32)(Tmax - c->hcount[DC_TUPLES_INVALID]);
  c->counters["tuples"] = (int64_t)T;
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
  DevBuf<u64> raw;
  u64 cap = (u64)T * 4 + 1024;
  u64 n_raw = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    raw.alloc(cap, s);
    P.edges_out = raw.p;
    P.edges_cap = cap;
    CB_CUDA(cudaMemsetAsync(c->dcount.p + DC_HKMER, 0, (DC_EDGE_REMOVAL - DC_HKMER) * sizeof(u64), s));
    int blocks = kNumSMs * 8;
    int need = div_up(((size_t)T + 31) / 32, JN_WARPS);
    if (need < blocks) blocks = need > 0 ? need : 1;
    c->lc.begin("join_verify");
    CB_DISPATCH_NW(c->NW, (join_verify_kernel<NW_><<<blocks, JN_WARPS * 32, 0, s>>>(P)));
    c->lc.end("join_verify");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    download_counters(c);
    n_raw = c->hcount[DC_EDGES_RAW];
    if (n_raw <= cap) break;
    if (attempt == 1) throw Error(CB_E_CAPACITY, "edge buffer overflow");
    cap = n_raw;  // exact size known now: run again
  }
  timer_stop(c, "join_verify");
  keys[where].release();
  vals[where].release();

  // BuildHighKmerList's counter (poly-A/T k-mer handled through its global weight)
  int64_t hkmer = (int64_t)c->hcount[DC_HKMER] + ((long long)c->hcount[DC_POLYA_WEIGHT] > c->cfg.up_kmer ? 1 : 0);
  c->counters["hkmer"] = hkmer;
  c->counters["candidates_verified"] = (int64_t)c->hcount[DC_CAND_VERIFIED];
  c->counters["upkmer_truncated_lists"] = (int64_t)c->hcount[DC_UPKMER_LISTS];
  c->counters["order_dependent"] = (int64_t)c->hcount[DC_UPKMER_LISTS];
  if (hkmer > 0)
    throw Error(CB_E_UNSUPPORTED,
                "non-empty high-frequency k-mer list: the reference's stopword loader (MatchPrefix.java:81) is "
                "ill-defined for it, parity undefined");

  // ---- S u mirror(S): sort + unique ----
  timer_start(c, "sort_edges");
  DevBuf<u64> raw2;
  raw2.alloc(n_raw, s);
  int w2 = radix_sort(raw.p, raw2.p, nullptr, nullptr, n_raw, 0, c->lay.total_bits(), s, c->lc, "edges");
  u64* sorted = w2 ? raw2.p : raw.p;
  DevBuf<u32> flag;
  flag.alloc(n_raw, s);
  if (n_raw) {
    unique_flags_kernel<<<div_up(n_raw, 256), 256, 0, s>>>(sorted, n_raw, flag.p);
    c->lc.n++;
  }
  c->ckA.n = compact_by_flags(c, sorted, flag.p, n_raw, c->ckA.keys);
  c->ckA.valid = true;
  timer_stop(c, "sort_edges");
  build_adjacency(c, c->ckA);
  c->counters["edges_overlap"] = (int64_t)c->ckA.n;
}

}  // namespace cb
