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

__device__ __forceinline__ u32 thread_term_mask(const char* __restrict__ text, size_t base, size_t n) {
  u32 m = 0;
  if (base + PT_BYTES < n) {  // whole chunk and the look-ahead byte are in range
    uint4 v = *reinterpret_cast<const uint4*>(text + base);
    u32 w[4] = {v.x, v.y, v.z, v.w};
    char nxt = text[base + PT_BYTES];
#pragma unroll
    for (int j = 0; j < PT_BYTES; j++) {
      char c = (char)((w[j >> 2] >> ((j & 3) * 8)) & 0xff);
      char c1 = (j + 1 < PT_BYTES) ? (char)((w[(j + 1) >> 2] >> (((j + 1) & 3) * 8)) & 0xff) : nxt;
      if (c == '\n' || (c == '\r' && c1 != '\n')) m |= 1u << j;
    }
  } else {
    for (int j = 0; j < PT_BYTES; j++)
      if (base + j < n && is_term(text, base + j, n)) m |= 1u << j;
  }
  return m;
}

__global__ void __launch_bounds__(PT_THREADS) count_terms_kernel(const char* __restrict__ text, size_t n,
                                                                 u32* __restrict__ block_counts) {
  __shared__ u32 tmp[PT_THREADS / 32 + 1];
  size_t base = (size_t)blockIdx.x * PT_BLOCK + (size_t)threadIdx.x * PT_BYTES;
  u32 cnt = base < n ? __popc(thread_term_mask(text, base, n)) : 0;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 v = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) tmp[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 total = 0;
    for (int w = 0; w < PT_THREADS / 32; w++) total += tmp[w];
    block_counts[blockIdx.x] = total;
  }
}

__global__ void __launch_bounds__(PT_THREADS) fill_line_starts_kernel(const char* __restrict__ text, size_t n,
                                                                      const u32* __restrict__ block_offs,
                                                                      u32* __restrict__ line_start) {
  __shared__ u32 tmp[PT_THREADS / 32 + 1];
  size_t base = (size_t)blockIdx.x * PT_BLOCK + (size_t)threadIdx.x * PT_BYTES;
  u32 m = base < n ? thread_term_mask(text, base, n) : 0;
  u32 cnt = __popc(m);
  // block exclusive scan of cnt
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) tmp[warp] = inc;
  __syncthreads();
  u32 woff = 0;
  for (int w = 0; w < warp; w++) woff += tmp[w];
  u32 rank = block_offs[blockIdx.x] + woff + inc - cnt;
  while (m) {
    int j = __ffs(m) - 1;
    m &= m - 1;
    line_start[++rank] = (u32)(base + j + 1);  // line (rank) starts after terminator #(rank-1)
  }
}

__global__ void set_u32_kernel(u32* p, u32 v) { *p = v; }

// One thread per SFA line: Java split("\t") semantics, uppercase, ACGT filter, length filter,
// 2-bit packing (row written with 16-byte vector stores), 16-byte id key, and the 64-bit hash
// of the canonical read min(s, rc(s)) that keys the duplicate detection.
template <int NW>
__global__ void __launch_bounds__(128) parse_lines_kernel(const char* __restrict__ text,
                                                          const u32* __restrict__ line_start, u32 n_lines, int K,
                                                          int L, uint8_t* __restrict__ status,
                                                          u64* __restrict__ reads, IdKey* __restrict__ idkey,
                                                          u64* __restrict__ hkeys, u32* __restrict__ hvals,
                                                          ull* __restrict__ dcount) {
  u32 line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  u32 s = line_start[line], e = line_start[line + 1];
  if (e > s && text[e - 1] == '\n') e--;
  if (e > s && text[e - 1] == '\r') e--;
  // field boundaries; String.split drops trailing empty fields
  u32 tab1 = e, tab2 = e;
  int field = 0, last_nonempty = -1;
  for (u32 p = s; p < e; p++) {
    char c = text[p];
    if (c == '\t') {
      if (field == 0) tab1 = p;
      else if (field == 1) tab2 = p;
      field++;
    } else {
      last_nonempty = field;
    }
  }
  int nf = last_nonempty + 1;
  u64 w[NW];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = 0;
  uint8_t st = 0;
  u64 hk = ~0ull;
  if (nf != 2 && nf != 3) {  // GenNonContainedReads.java:64-69
    st = 1;
    atomicAdd(&dcount[DC_LINES_INVALID], 1ull);
  } else {
    u32 q0 = tab1 + 1, q1 = tab2;
    int len = (int)(q1 - q0);
    bool bad = false;
    for (int i = 0; i < len; i++) {
      char c = text[q0 + i];
      if (c >= 'a' && c <= 'z') c = (char)(c - 32);  // toUpperCase :79
      u64 code;
      if (c == 'A') code = 0;
      else if (c == 'C') code = 1;
      else if (c == 'G') code = 2;
      else if (c == 'T') code = 3;
      else { bad = true; break; }
      if (i < NW * 32) {
        int wi = i >> 5;
#pragma unroll
        for (int j = 0; j < NW; j++)
          if (j == wi) w[j] |= code << (62 - 2 * (i & 31));
      }
    }
    if (bad) {  // :102-107
      st = 2;
      atomicAdd(&dcount[DC_READS_SKIPPED], 1ull);
    } else if (len <= K) {  // :110-115
      st = 3;
      atomicAdd(&dcount[DC_READS_SHORT], 1ull);
    } else {
      atomicAdd(&dcount[DC_READS_GOOD], 1ull);
      atomicAdd(&dcount[DC_READS_GOODBP], (ull)len);
      if (len != L) atomicAdd(&dcount[DC_LEN_MISMATCH], 1ull);
      u32 idlen = tab1 - s;
      if (idlen > 16) atomicAdd(&dcount[DC_LONG_ID], 1ull);
      IdKey ik;
      ik.hi = 0;
      ik.lo = 0;
      for (u32 i = 0; i < idlen && i < 16; i++) {
        u64 b = (u64)(unsigned char)text[s + i];
        if (i < 8) ik.hi |= b << (56 - 8 * i);
        else ik.lo |= b << (56 - 8 * (i - 8));
      }
      idkey[line] = ik;
      u64 rc[NW];
      revcomp_read<NW>(w, rc, L);
      bool fwd = cmp_words<NW>(w, rc) <= 0;
      u64 h = 0x9e3779b97f4a7c15ull;
#pragma unroll
      for (int i = 0; i < NW; i++) h = mix64(h ^ (fwd ? w[i] : rc[i])) + 0x9e3779b97f4a7c15ull;
      hk = h & ~1ull;  // bit 0 clear: never collides with the all-ones key of non-reads
    }
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
}

template <int NW>
__device__ __forceinline__ void load_read(const u64* __restrict__ reads, u32 idx, u64 (&r)[NW]) {
  const uint4* row = reinterpret_cast<const uint4*>(reads + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) {
    uint4 v = __ldg(row + i);
    r[2 * i] = (u64)v.x | ((u64)v.y << 32);
    r[2 * i + 1] = (u64)v.z | ((u64)v.w << 32);
  }
}

template <int NW>
__device__ __forceinline__ void load_canon(const u64* __restrict__ reads, u32 idx, int L, u64 (&canon)[NW],
                                           bool& selfrc) {
  u64 r[NW], rc[NW];
  load_read<NW>(reads, idx, r);
  revcomp_read<NW>(r, rc, L);
  int c = cmp_words<NW>(r, rc);
  selfrc = (c == 0);
#pragma unroll
  for (int i = 0; i < NW; i++) canon[i] = c <= 0 ? r[i] : rc[i];
}

// Sorted by canonical-read hash: the head of every run of equal (masked) hashes resolves the
// duplicate classes inside its run by exact comparison of the canonical reads.
template <int NW>
__global__ void __launch_bounds__(128) dedupe_runs_kernel(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                                          u32 n_good, u64 keymask, const u64* __restrict__ reads,
                                                          const IdKey* __restrict__ idkey, int L,
                                                          u32* __restrict__ surv_flag, u32* __restrict__ cov,
                                                          ull* __restrict__ dcount) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_good) return;
  u64 k = keys[i] & keymask;
  if (i > 0 && (keys[i - 1] & keymask) == k) return;
  u32 e = i + 1;
  while (e < n_good && (keys[e] & keymask) == k) e++;
  for (u32 a = i; a < e; a++) {
    u64 ca[NW], cb_[NW];
    bool self, self2;
    load_canon<NW>(reads, vals[a], L, ca, self);
    bool seen = false;
    for (u32 b = i; b < a; b++) {
      load_canon<NW>(reads, vals[b], L, cb_, self2);
      if (cmp_words<NW>(ca, cb_) == 0) { seen = true; break; }
    }
    if (seen) continue;
    u32 cnt = 1, best = vals[a];
    IdKey bk = idkey[best];
    for (u32 b = a + 1; b < e; b++) {
      load_canon<NW>(reads, vals[b], L, cb_, self2);
      if (cmp_words<NW>(ca, cb_) == 0) {
        cnt++;
        IdKey kb = idkey[vals[b]];
        if (idkey_less(kb, bk)) { bk = kb; best = vals[b]; }
      }
    }
    u32 cv = 1 + (cnt - 1) * (self ? 2u : 1u);  // GenNonContainedReads.java:176-201
    surv_flag[best] = 1;
    cov[best] = cv;
    atomicAdd(&dcount[DC_SURVIVORS], 1ull);
    atomicMax(&dcount[DC_MAX_COV], (ull)cv);
  }
}

template <int NW>
__global__ void __launch_bounds__(256) build_nodes_kernel(const u32* __restrict__ surv_flag,
                                                          const u32* __restrict__ surv_idx, u32 n_lines,
                                                          const u64* __restrict__ reads,
                                                          const IdKey* __restrict__ idkey, const u32* __restrict__ cov,
                                                          u32* __restrict__ node_line, u64* __restrict__ node_seq,
                                                          IdKey* __restrict__ node_idkey, u32* __restrict__ node_cov) {
  u32 line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines || !surv_flag[line]) return;
  u32 idx = surv_idx[line];
  node_line[idx] = line;
  node_idkey[idx] = idkey[line];
  node_cov[idx] = cov[line];
  const uint4* src = reinterpret_cast<const uint4*>(reads + (size_t)line * NW);
  uint4* dst = reinterpret_cast<uint4*>(node_seq + (size_t)idx * NW);
#pragma unroll
  for (int i = 0; i < NW / 2; i++) dst[i] = src[i];
}

#define CB_DISPATCH_NW(nw, CALL)                                             \
  switch (nw) {                                                              \
    case 2: { constexpr int NW_ = 2; CALL; } break;                          \
    case 4: { constexpr int NW_ = 4; CALL; } break;                          \
    case 6: { constexpr int NW_ = 6; CALL; } break;                          \
    case 8: { constexpr int NW_ = 8; CALL; } break;                          \
    default: throw Error(CB_E_UNSUPPORTED, "read length above 256 bases");   \
  }

void stage_parse_and_pack(cb_ctx* c) {
  cudaStream_t s = c->stream;
  const size_t n = c->text_n;
  if (n == 0) throw Error(CB_E_NOREADS, "No good reads");
  if (n >= (1ull << 32)) throw Error(CB_E_CAPACITY, "SFA input of 4 GiB or more per context");
  c->K = c->cfg.k;
  c->L = c->cfg.readlen;
  if (c->K < 1 || c->K > 31 || (c->K % 2) == 0)
    throw Error(CB_E_UNSUPPORTED, "k must be odd and <= 31 (even k lets palindromic k-mers drop nodes; SURVEY 8c)");
  if (c->L <= c->K) throw Error(CB_E_INVALID, "readlen must exceed k");
  if (c->L > 255) throw Error(CB_E_UNSUPPORTED, "readlen above 255");
  c->NW = ((c->L + 31) / 32 + 1) & ~1;
  c->W = c->L - c->K + 1;
  c->dcount.alloc(DC_COUNT, s);
  CB_CUDA(cudaMemsetAsync(c->dcount.p, 0, DC_COUNT * sizeof(u64), s));

  timer_start(c, "parse");
  unsigned nb = (unsigned)div_up(n, PT_BLOCK);
  DevBuf<u32> bcnt, boff;
  bcnt.alloc(nb, s);
  boff.alloc(nb, s);
  count_terms_kernel<<<nb, PT_THREADS, 0, s>>>(c->text_ptr, n, bcnt.p);
  c->lc.n++;
  exclusive_scan_u32(bcnt.p, boff.p, nb, s, c->lc);
  u32 last_cnt = 0, last_off = 0;
  char last_byte = 0;
  CB_CUDA(cudaMemcpyAsync(&last_cnt, bcnt.p + nb - 1, 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(&last_off, boff.p + nb - 1, 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(&last_byte, c->text_ptr + n - 1, 1, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  u32 terms = last_cnt + last_off;
  bool last_is_term = (last_byte == '\n' || last_byte == '\r');
  c->n_lines = terms + (last_is_term ? 0 : 1);
  c->line_start.alloc((size_t)c->n_lines + 1, s);
  set_u32_kernel<<<1, 1, 0, s>>>(c->line_start.p, 0u);
  fill_line_starts_kernel<<<nb, PT_THREADS, 0, s>>>(c->text_ptr, n, boff.p, c->line_start.p);
  set_u32_kernel<<<1, 1, 0, s>>>(c->line_start.p + c->n_lines, (u32)n);
  c->lc.n += 3;

  const u32 nl = c->n_lines;
  c->status.alloc(nl, s);
  c->reads.alloc((size_t)nl * c->NW, s);
  c->idkey.alloc(nl, s);
  // the hash keys / values of the duplicate detection live until stage_dedupe
  DevBuf<u64> hk[2];
  DevBuf<u32> hv[2];
  {
    for (int i = 0; i < 2; i++) { hk[i].alloc(nl, s); hv[i].alloc(nl, s); }
    c->lc.begin("parse_lines");
    CB_DISPATCH_NW(c->NW, (parse_lines_kernel<NW_><<<div_up(nl, 128), 128, 0, s>>>(
                              c->text_ptr, c->line_start.p, nl, c->K, c->L, c->status.p, c->reads.p, c->idkey.p,
                              hk[0].p, hv[0].p, (ull*)c->dcount.p)));
    c->lc.end("parse_lines");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    timer_stop(c, "parse");
    download_counters(c);
    if (c->hcount[DC_READS_GOOD] == 0) throw Error(CB_E_NOREADS, "No good reads");
    if (c->hcount[DC_LEN_MISMATCH])
      throw Error(CB_E_UNSUPPORTED, "reads of mixed length: " + std::to_string(c->hcount[DC_LEN_MISMATCH]) +
                                        " good reads differ from -readlen " + std::to_string(c->L));
    if (c->hcount[DC_LONG_ID]) throw Error(CB_E_UNSUPPORTED, "read ids longer than 16 bytes");

    // ---- duplicate detection ----
    timer_start(c, "dedupe");
    const u32 n_good = (u32)c->hcount[DC_READS_GOOD];
    int bits = 0;
    while ((1ull << bits) < (u64)n_good) bits++;
    int hb = ((bits + 16 + 7) / 8) * 8;
    if (hb > 64) hb = 64;
    const u64 keymask = hb == 64 ? ~0ull : ((1ull << hb) - 1);
    int where = radix_sort(hk[0].p, hk[1].p, hv[0].p, hv[1].p, nl, 0, hb, s, c->lc, "dedupe");
    DevBuf<u32> surv_flag, surv_idx, cov;
    surv_flag.alloc(nl, s);
    surv_idx.alloc(nl, s);
    cov.alloc(nl, s);
    CB_CUDA(cudaMemsetAsync(surv_flag.p, 0, (size_t)nl * 4, s));
    c->lc.begin("dedupe_runs");
    CB_DISPATCH_NW(c->NW, (dedupe_runs_kernel<NW_><<<div_up(n_good, 128), 128, 0, s>>>(
                              hk[where].p, hv[where].p, n_good, keymask, c->reads.p, c->idkey.p, c->L, surv_flag.p,
                              cov.p, (ull*)c->dcount.p)));
    c->lc.end("dedupe_runs");
    c->lc.n++;
    exclusive_scan_u32(surv_flag.p, surv_idx.p, nl, s, c->lc);
    download_counters(c);
    c->n_nodes = (u32)c->hcount[DC_SURVIVORS];
    const u32 nn = c->n_nodes;
    c->node_line.alloc(nn, s);
    c->node_seq.alloc((size_t)nn * c->NW, s);
    c->node_idkey.alloc(nn, s);
    c->node_cov.alloc(nn, s);
    CB_DISPATCH_NW(c->NW, (build_nodes_kernel<NW_><<<div_up(nl, 256), 256, 0, s>>>(
                              surv_flag.p, surv_idx.p, nl, c->reads.p, c->idkey.p, cov.p, c->node_line.p,
                              c->node_seq.p, c->node_idkey.p, c->node_cov.p)));
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    timer_stop(c, "dedupe");
  }

  // counters with the reference's names (BrushAssembler.java:266-289)
  auto& C = c->counters;
  C["reads_good"] = (int64_t)c->hcount[DC_READS_GOOD];
  C["reads_goodbp"] = (int64_t)c->hcount[DC_READS_GOODBP];
  C["reads_short"] = (int64_t)c->hcount[DC_READS_SHORT];
  C["reads_skipped"] = (int64_t)c->hcount[DC_READS_SKIPPED];
  C["input_lines_invalid"] = (int64_t)c->hcount[DC_LINES_INVALID];
  C["nodecount"] = (int64_t)c->hcount[DC_READS_GOOD];
  C["contained"] = (int64_t)(c->hcount[DC_READS_GOOD] - c->hcount[DC_SURVIVORS]);
  C["redundant"] = C["contained"];
  C["nodes"] = (int64_t)c->n_nodes;
  // free what later stages do not need (node arrays carry everything forward)
  c->reads.release();
  c->idkey.release();
  c->status.release();
}

void stage_dedupe(cb_ctx*) {}

}  // namespace cb
