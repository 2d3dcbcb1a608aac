// capi.cu -- the C ABI of include/cloudbrush_gpu.h: context, input, counters, export, node text.
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "ctx.cuh"

namespace cb {

void timer_start(cb_ctx* c, const char* name) {
  Timer& t = c->timers[name];
  if (!t.a) {
    cudaEventCreate(&t.a);
    cudaEventCreate(&t.b);
  }
  t.done = false;
  cudaEventRecord(t.a, c->stream);
}
void timer_stop(cb_ctx* c, const char* name) {
  Timer& t = c->timers[name];
  cudaEventRecord(t.b, c->stream);
  t.done = true;
}

void download_counters(cb_ctx* c) {
  c->lc.rb.get(c->hcount, c->dcount.p, DC_COUNT * sizeof(u64), c->stream);
  c->lc.rb.sync(c->stream);  // (delivers every small result enqueued before it as well)
}

// canonical export order (node, type, nbr, ov): compact the occupied slots, re-encode, sort, decode
__global__ void __launch_bounds__(256) export_flags_kernel(const u64* __restrict__ keys, size_t n,
                                                           u32* __restrict__ flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = keys[i] != ~0ull ? 1u : 0u;
}
__global__ void __launch_bounds__(256) export_gather_kernel(const u64* __restrict__ keys, const u32* __restrict__ flag,
                                                            const u32* __restrict__ idx, size_t n, EdgeLayout lay,
                                                            int K, u64* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  u64 k = keys[i];
  u64 type = (u64)lay.x(k) * 2 + lay.y(k);
  u64 ovf = (u64)(lay.ov(k) - (u32)K);  // K <= ov <= L, so ov - K fits ob bits
  out[idx[i]] = ((u64)lay.node(k) << (lay.nb + lay.ob + 2)) | (type << (lay.nb + lay.ob)) |
                ((u64)lay.nbr(k) << lay.ob) | ovf;
}
__global__ void __launch_bounds__(256) export_decode_kernel(const u64* __restrict__ in, size_t n, EdgeLayout lay,
                                                            int K, const u32* __restrict__ node_line,
                                                            cb_edge* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 k = in[i];
  cb_edge e;
  e.node = node_line[(u32)(k >> (lay.nb + lay.ob + 2))];
  e.type = (uint8_t)((k >> (lay.nb + lay.ob)) & 3);
  e.nbr = node_line[(u32)((k >> lay.ob) & ((1ull << lay.nb) - 1))];
  e.ov = (uint16_t)((u32)(k & ((1ull << lay.ob) - 1)) + (u32)K);
  e.pad = 0;
  out[i] = e;
}

__global__ void __launch_bounds__(256) export_nodes_kernel(const u32* __restrict__ node_line,
                                                           const u32* __restrict__ node_cov, u32 n,
                                                           cb_node* __restrict__ out) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cb_node v;
  v.line = node_line[i];
  v.cov = node_cov[i];
  out[i] = v;
}

// records of a checkpoint, or -1 if it is not available yet
static long long edge_count(cb_ctx* c, int ck) {
  switch (ck) {
    case CB_CK_OVERLAP: return c->overlapped ? (long long)c->ckA.n : -1;
    case CB_CK_CHL2: return c->reduced ? (long long)c->n_edges_c : -1;
    case CB_CK_RE: return c->reduced ? (long long)c->ckB.n : -1;
  }
  return -1;
}

__global__ void __launch_bounds__(256) export_recode_kernel(const u64* __restrict__ keys, size_t n, EdgeLayout lay,
                                                            int K, u64* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u64 k = keys[i];
  u64 type = (u64)lay.x(k) * 2 + lay.y(k);
  u64 ovf = (u64)(lay.ov(k) - (u32)K);
  out[i] = ((u64)lay.node(k) << (lay.nb + lay.ob + 2)) | (type << (lay.nb + lay.ob)) | ((u64)lay.nbr(k) << lay.ob) | ovf;
}

// sorted cb_edge records of a checkpoint straight into caller memory (pinned memory makes the
// device-to-host copy run at PCIe speed).  01-overlap.re is a compact record list; 01-overlap and
// 01-overlap.chl2 share the slot array (ctx.cuh): once CutChimericLinks has removed records in
// place, 01-overlap is exported from a restored private copy.
// the records of a checkpoint this context holds, re-encoded as (node | type | nbr | ov - K) so that an
// ascending sort gives the canonical export order; unsorted, on the device
static void export_encoded(cb_ctx* c, int ck, DevBuf<u64>& a) {
  cudaStream_t s = c->stream;
  const size_t n = (size_t)edge_count(c, ck);
  a.alloc(n ? n : 1, s);
  if (n == 0) return;
  if (ck == CB_CK_RE) {
    export_recode_kernel<<<div_up(n, 256), 256, 0, s>>>(c->ckB.keys.p, n, c->lay, c->K, a.p);
    c->lc.n++;
  } else {
    const size_t slots = c->slots;
    DevBuf<u64> restored;
    const u64* keys = c->ckA.keys.p;
    if (ck == CB_CK_OVERLAP && c->n_cut_log) {
      restore_overlap_copy(c, restored);
      keys = restored.p;
    }
    DevBuf<u32> flag, idx;
    flag.alloc(slots, s);
    idx.alloc(slots, s);
    export_flags_kernel<<<div_up(slots, 256), 256, 0, s>>>(keys, slots, flag.p);
    exclusive_scan_u32(flag.p, idx.p, slots, s, c->lc);
    export_gather_kernel<<<div_up(slots, 256), 256, 0, s>>>(keys, flag.p, idx.p, slots, c->lay, c->K, a.p);
    c->lc.n += 2;
  }
}

// the n > 0 records of a checkpoint in export order, on the device
static void export_edges_dev(cb_ctx* c, int ck, DevBuf<cb_edge>& d) {
  cudaStream_t s = c->stream;
  const size_t n = (size_t)edge_count(c, ck);
  DevBuf<u64> a, b;
  export_encoded(c, ck, a);
  b.alloc(n, s);
  int w = radix_sort(a.p, b.p, nullptr, nullptr, n, 0, c->lay.total_bits(), s, c->lc, "export");
  d.alloc(n, s);
  export_decode_kernel<<<div_up(n, 256), 256, 0, s>>>(w ? b.p : a.p, n, c->lay, c->K, c->node_line.p, d.p);
  c->lc.n++;
}

static void export_edges_to(cb_ctx* c, int ck, cb_edge* out) {
  cudaStream_t s = c->stream;
  const size_t n = (size_t)edge_count(c, ck);
  if (n == 0) return;
  DevBuf<cb_edge> d;
  export_edges_dev(c, ck, d);
  CB_CUDA(cudaMemcpyAsync(out, d.p, n * sizeof(cb_edge), cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
}

// cb_export_*_begin: the copy of a staged result to the caller's memory, behind whatever the context's stream does next
static void d2h_begin(cb_ctx* c, void* out, const void* dev, size_t bytes) {
  if (!c->d2h_stream) {
    CB_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming));
  }
  CB_CUDA(cudaEventRecord(c->ev_d2h, c->stream));
  CB_CUDA(cudaStreamWaitEvent(c->d2h_stream, c->ev_d2h, 0));
  // (the stage groups' own small read-backs do not queue behind this transfer: they do not use the copy engine, ReadBack)
  CB_CUDA(cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, c->d2h_stream));
  c->d2h_pending = true;
}


// ---------------------------------------------------------------------------------------------
// Node text on the device (Node.toNodeMsg, Node.java:1757-1982): the host formatter
// spent 9 s on the 1.4 GB of 01-overlap at config 2 (tools/dropin_e2e.py).  One thread per node: the
// record's length, an exclusive scan, then the bytes:
//   id \t N \t *s \t <pair-coded dna> \t *v \t <cov>.00 [\t *ff \t id!ov ...] [*fr] [*rf] [*rr] \n
// Records arrive sorted by (node, type, nbr, ov) in the export encoding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 idkey_len(const IdKey& k) {
  if (k.lo) return 16u - (u32)(__ffsll((long long)k.lo) - 1) / 8u;
  if (k.hi) return 8u - (u32)(__ffsll((long long)k.hi) - 1) / 8u;
  return 0u;
}
__device__ __forceinline__ u32 dec_digits(u32 v) {
  u32 d = 1;
  while (v >= 10u) { v /= 10u; d++; }
  return d;
}
__device__ __forceinline__ char* put_id(char* p, const IdKey& k) {
  const u32 n = idkey_len(k);
  for (u32 i = 0; i < n; i++) *p++ = (char)((i < 8 ? k.hi >> (56 - 8 * i) : k.lo >> (56 - 8 * (i - 8))) & 0xff);
  return p;
}
__device__ __forceinline__ char* put_dec(char* p, u32 v) {
  const u32 d = dec_digits(v);
  for (u32 i = d; i-- > 0;) { p[i] = (char)('0' + v % 10u); v /= 10u; }
  return p + d;
}

// first record of node lo + i in the sorted records (lower bound); first[nn] = n
__global__ void __launch_bounds__(256) text_first_kernel(const u64* __restrict__ recs, size_t n, int shift, u32 lo, u32 nn,
                                                         u32* __restrict__ first) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nn) return;
  const u64 want = (u64)lo + i;
  size_t a = 0, b = n;
  while (a < b) {
    const size_t m = a + ((b - a) >> 1);
    if ((recs[m] >> shift) < want) a = m + 1;
    else b = m;
  }
  first[i] = (u32)a;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) text_nodes_kernel(const u64* __restrict__ recs, const u32* __restrict__ first,
                                                         u32 lo, u32 nn, EdgeLayout lay, int K, int L, int NW,
                                                         const u64* __restrict__ node_seq, const u32* __restrict__ node_cov,
                                                         const IdKey* __restrict__ idkey, u32* __restrict__ len,
                                                         const u32* __restrict__ off, char* __restrict__ out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nn) return;
  const u32 node = lo + i;
  const int nb = lay.nb, ob = lay.ob;
  const u64 nmask = (1ull << nb) - 1, omask = (1ull << ob) - 1;
  const u32 e0 = first[i], e1 = first[i + 1];
  if (!WRITE) {
    u32 n = idkey_len(idkey[node]) + 6u + (u32)(L + 1) / 2u + 4u + dec_digits(node_cov[node]) + 3u + 1u;
    int cur = -1;
    for (u32 e = e0; e < e1; e++) {
      const u64 k = recs[e];
      const int type = (int)((k >> (nb + ob)) & 3);
      if (type != cur) { cur = type; n += 4u; }
      n += 1u + idkey_len(idkey[(u32)((k >> ob) & nmask)]) + 1u + dec_digits((u32)(k & omask) + (u32)K);
    }
    len[i] = n;
    return;
  }
  char* p = out + off[i];
  p = put_id(p, idkey[node]);
  *p++ = '\t'; *p++ = 'N'; *p++ = '\t'; *p++ = '*'; *p++ = 's'; *p++ = '\t';
  const u64* row = node_seq + (size_t)node * NW;
  for (int b = 0; b < L; b += 2) {  // Node.str2dna (Node.java:186-211): two bases per letter, a lone last base alone
    const int c0 = (int)((row[b >> 5] >> (62 - 2 * (b & 31))) & 3);
    if (b + 1 < L) {
      const int c1 = (int)((row[(b + 1) >> 5] >> (62 - 2 * ((b + 1) & 31))) & 3);
      *p++ = (char)('A' + 5 * c0 + 1 + c1);
    } else {
      *p++ = (char)('A' + 5 * c0);
    }
  }
  *p++ = '\t'; *p++ = '*'; *p++ = 'v'; *p++ = '\t';
  p = put_dec(p, node_cov[node]);
  *p++ = '.'; *p++ = '0'; *p++ = '0';  // DecimalFormat("0.00") of an integral coverage (Node.java:1761,1775)
  int cur = -1;
  for (u32 e = e0; e < e1; e++) {
    const u64 k = recs[e];
    const int type = (int)((k >> (nb + ob)) & 3);
    if (type != cur) {
      cur = type;
      *p++ = '\t'; *p++ = '*'; *p++ = (type & 2) ? 'r' : 'f'; *p++ = (type & 1) ? 'r' : 'f';
    }
    *p++ = '\t';
    p = put_id(p, idkey[(u32)((k >> ob) & nmask)]);
    *p++ = '!';
    p = put_dec(p, (u32)(k & omask) + (u32)K);
  }
  *p++ = '\n';
}

__global__ void __launch_bounds__(256) sum_u32_kernel(const u32* __restrict__ v, u32 n, ull* __restrict__ out) {
  ull acc = 0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// owner rank of the node of an export-encoded record: the rank whose SFA shard the node was built from
__global__ void __launch_bounds__(256) node_owner_kernel(const u64* __restrict__ recs, size_t n, int shift,
                                                         const u64* __restrict__ node_bases, int world,
                                                         u32* __restrict__ dest) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 node = recs[i] >> shift;
  int d = 0;
  while (d + 1 < world && node >= node_bases[d + 1]) d++;
  dest[i] = (u32)d;
}

static void export_edges_host(cb_ctx* c, int ck, std::vector<cb_edge>& out) {
  out.resize((size_t)edge_count(c, ck));
  export_edges_to(c, ck, out.data());
}

static void read_file(const std::string& path, std::vector<char>& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(CB_E_IO, "cannot open " + path);
  f.seekg(0, std::ios::end);
  std::streamoff sz = f.tellg();
  f.seekg(0);
  size_t at = out.size();
  out.resize(at + (size_t)sz);
  if (sz) f.read(out.data() + at, sz);
  if (!out.empty() && out.back() != '\n' && out.back() != '\r') out.push_back('\n');  // files are split separately
}

// a blocking load replaces whatever was prefetched and not consumed
static void drop_prefetched(cb_ctx* c) {
  if (c->next_q.empty()) return;
  cudaStreamSynchronize(c->copy_stream);
  for (auto& e : c->next_q)
    if (e.done) cudaEventDestroy(e.done);
  c->next_q.clear();
}

static void reset_after_load(cb_ctx* c) {
  c->loaded = true;
  c->preprocessed = c->overlapped = c->reduced = false;
  c->counters.clear();
  c->ckA = EdgeSet();
  c->ckB = EdgeSet();
  c->n_cut_log = 0;
  c->n_edges_c = 0;
}

}  // namespace cb

using namespace cb;

#define CB_API_BEGIN(ctx)            \
  if (!(ctx)) return CB_E_INVALID;   \
  try {                              \
    cudaSetDevice((ctx)->cfg.device); \
    cb::current_arena() = &(ctx)->arena;

#define CB_API_END(ctx)                        \
  }                                            \
  catch (const cb::Error& e) {                 \
    (ctx)->err = e.what();                     \
    return e.code;                             \
  }                                            \
  catch (const std::bad_alloc&) {              \
    (ctx)->err = "host allocation failed";     \
    return CB_E_NOMEM;                         \
  }                                            \
  catch (const std::exception& e) {            \
    (ctx)->err = e.what();                     \
    return CB_E_INVALID;                       \
  }                                            \
  return CB_OK;

extern "C" {

const char* cb_version(void) { return "cloudbrush-b200 0.1 (sm_100a)"; }

void cb_default_config(cb_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = CB_ABI_VERSION;
  cfg->k = 21;
  cfg->readlen = 36;
  cfg->up_kmer = 2000;  // BrushConfig.java:84
  cfg->low_kmer = 1;    // BrushConfig.java:83
  cfg->majority = 0.6f; // BrushConfig.java:62
  cfg->pwm_n = 0.1f;    // BrushConfig.java:63
  cfg->device = 0;
  cfg->rank = 0;
  cfg->world = 1;
}

int cb_create(const cb_config* cfg, cb_ctx** out) {
  if (!cfg || !out) return CB_E_INVALID;
  *out = nullptr;
  if (cfg->abi_version != CB_ABI_VERSION) return CB_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CB_E_NODEVICE;  // no CPU fallback
  if (cfg->device < 0 || cfg->device >= ndev) return CB_E_NODEVICE;
  cb_ctx* c = new (std::nothrow) cb_ctx();
  if (!c) return CB_E_NOMEM;
  c->cfg = *cfg;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return CB_E_CUDA;
  }
  c->lc.stream = c->stream;
  *out = c;
  return CB_OK;
}

void cb_destroy(cb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  for (auto& kv : c->timers) {
    if (kv.second.a) cudaEventDestroy(kv.second.a);
    if (kv.second.b) cudaEventDestroy(kv.second.b);
  }
  cb::nccl_shutdown(c);
  if (c->d2h_stream) {
    cudaStreamSynchronize(c->d2h_stream);
    cudaStreamDestroy(c->d2h_stream);
    cudaEventDestroy(c->ev_d2h);
  }
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
    for (auto& e : c->next_q)
      if (e.done) cudaEventDestroy(e.done);
    cudaEventDestroy(c->ev_main);
  }
  cudaStream_t s = c->stream;
  delete c;  // DevBufs return to the arena, the arena returns everything to the driver
  cudaStreamDestroy(s);
}

const char* cb_last_error(const cb_ctx* c) { return c ? c->err.c_str() : "null context"; }

int cb_load_sfa_device(cb_ctx* c, const void* dev_buf, size_t n) {
  CB_API_BEGIN(c)
  if (!dev_buf && n) throw Error(CB_E_INVALID, "null buffer");
  c->text.alloc(n + 32, c->stream);
  c->text_n = n;
  if (n) CB_CUDA(cudaMemcpyAsync(c->text.p, dev_buf, n, cudaMemcpyDeviceToDevice, c->stream));
  CB_CUDA(cudaMemsetAsync(c->text.p + n, 0, 32, c->stream));
  CB_CUDA(cudaStreamSynchronize(c->stream));  // the caller's buffer is only borrowed for the call
  c->text_ptr = c->text.p;
  drop_prefetched(c);
  reset_after_load(c);
  CB_API_END(c)
}

int cb_borrow_sfa_device(cb_ctx* c, const void* dev_buf, size_t n) {
  CB_API_BEGIN(c)
  if (!dev_buf && n) throw Error(CB_E_INVALID, "null buffer");
  // the parser reads the text in aligned 16-byte vectors (stage_preprocess.cu: thread_term_mask, parse_lines)
  if (((uintptr_t)dev_buf & 15) != 0)
    throw Error(CB_E_INVALID, "cb_borrow_sfa_device: the buffer must be 16-byte aligned (use cb_load_sfa_device to copy)");
  c->text.release();
  c->text_ptr = (const char*)dev_buf;
  c->text_n = n;
  drop_prefetched(c);
  reset_after_load(c);
  CB_API_END(c)
}

int cb_set_exchange(cb_ctx* c, const cb_exchange* ex) {
  if (!c || !ex || !ex->exchange_counts || !ex->alltoallv || !ex->allgatherv || !ex->allreduce_u64 ||
      !ex->allreduce_max_u8_dev) return CB_E_INVALID;
  c->ex = *ex;
  c->has_ex = true;
  return CB_OK;
}

static std::string& nccl_error_string() {
  static thread_local std::string e;
  return e;
}

int cb_nccl_get_unique_id(void* id128) {
  if (!id128) return CB_E_INVALID;
  nccl_error_string().clear();
  return cb::nccl_unique_id(id128, nccl_error_string());
}

const char* cb_nccl_last_error(void) { return nccl_error_string().c_str(); }

int cb_init_nccl(cb_ctx* c, const void* id128) {
  CB_API_BEGIN(c)
  if (!id128) throw Error(CB_E_INVALID, "null id");
  nccl_init(c, id128);
  CB_API_END(c)
}

int cb_set_profile(cb_ctx* c, int on) {
  if (!c) return CB_E_INVALID;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  c->lc.profile = on != 0;
  c->lc.reset();
  return CB_OK;
}

int cb_get_kernel_stat(cb_ctx* c, const char* name, double* total_ms, int64_t* launches) {
  if (!c || !name || !total_ms || !launches) return CB_E_INVALID;
  cudaSetDevice(c->cfg.device);
  if (!c->lc.query(name, *total_ms, *launches)) {
    c->err = std::string("no such kernel stat: ") + name;
    return CB_E_INVALID;
  }
  return CB_OK;
}

int cb_load_sfa_buffer(cb_ctx* c, const char* buf, size_t n) {
  CB_API_BEGIN(c)
  if (!buf && n) throw Error(CB_E_INVALID, "null buffer");
  c->text.alloc(n + 32, c->stream);
  c->text_n = n;
  if (n) CB_CUDA(cudaMemcpyAsync(c->text.p, buf, n, cudaMemcpyHostToDevice, c->stream));
  CB_CUDA(cudaMemsetAsync(c->text.p + n, 0, 32, c->stream));
  CB_CUDA(cudaStreamSynchronize(c->stream));  // the caller's buffer is only borrowed for the call
  c->text_ptr = c->text.p;
  drop_prefetched(c);
  reset_after_load(c);
  CB_API_END(c)
}

int cb_prefetch_sfa_buffer(cb_ctx* c, const char* buf, size_t n) {
  CB_API_BEGIN(c)
  if (!buf && n) throw Error(CB_E_INVALID, "null buffer");
  if (!c->copy_stream) {
    CB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
  }
  if (c->next_q.size() >= 2) throw Error(CB_E_INVALID, "cb_prefetch_sfa_buffer: two prefetched inputs are waiting already");
  c->next_q.emplace_back();
  cb_ctx::Prefetched& e = c->next_q.back();
  try {
    e.buf.alloc(n + 32, c->stream);
    e.n = n;
    CB_CUDA(cudaEventCreateWithFlags(&e.done, cudaEventDisableTiming));
    // the block may have been used by work still queued on the context's stream: the copy starts after it
    CB_CUDA(cudaEventRecord(c->ev_main, c->stream));
    CB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_main, 0));
    if (n) CB_CUDA(cudaMemcpyAsync(e.buf.p, buf, n, cudaMemcpyHostToDevice, c->copy_stream));
    CB_CUDA(cudaMemsetAsync(e.buf.p + n, 0, 32, c->copy_stream));
    CB_CUDA(cudaEventRecord(e.done, c->copy_stream));
  } catch (...) {
    cudaStreamSynchronize(c->copy_stream);
    if (e.done) cudaEventDestroy(e.done);
    c->next_q.pop_back();
    throw;
  }
  CB_API_END(c)
}

int cb_load_sfa(cb_ctx* c, const char* path) {
  CB_API_BEGIN(c)
  if (!path) throw Error(CB_E_INVALID, "null path");
  struct stat st;
  if (stat(path, &st) != 0) throw Error(CB_E_IO, std::string("cannot stat ") + path);
  std::vector<char> all;
  if (S_ISDIR(st.st_mode)) {
    std::vector<std::string> names;
    DIR* d = opendir(path);
    if (!d) throw Error(CB_E_IO, std::string("cannot open directory ") + path);
    while (dirent* de = readdir(d)) {
      std::string nm = de->d_name;
      if (nm.empty() || nm[0] == '.' || nm[0] == '_') continue;  // Hadoop hides . and _ files
      std::string full = std::string(path) + "/" + nm;
      struct stat fs;
      if (stat(full.c_str(), &fs) == 0 && S_ISREG(fs.st_mode)) names.push_back(full);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string& f : names) read_file(f, all);
  } else {
    read_file(path, all);
  }
  int rc = cb_load_sfa_buffer(c, all.data(), all.size());
  if (rc != CB_OK) return rc;
  CB_API_END(c)
}

int cb_preprocess(cb_ctx* c) {
  CB_API_BEGIN(c)
  if (!c->next_q.empty()) {  // cb_prefetch_sfa_buffer: the oldest prefetched input becomes the current one
    cb_ctx::Prefetched& e = c->next_q.front();
    c->text = std::move(e.buf);  // (the previous input's block goes back to the arena: stream-ordered)
    c->text_n = e.n;
    c->text_ptr = c->text.p;
    reset_after_load(c);
    CB_CUDA(cudaStreamWaitEvent(c->stream, e.done, 0));
    cudaEventDestroy(e.done);  // (released once the wait above has been satisfied)
    c->next_q.pop_front();
  }
  if (!c->loaded) throw Error(CB_E_INVALID, "cb_preprocess: no input loaded");
  c->preprocessed = c->overlapped = c->reduced = false;
  timer_start(c, "preprocess");
  stage_parse_and_pack(c);
  timer_stop(c, "preprocess");
  c->preprocessed = true;
  CB_API_END(c)
}

int cb_build_overlap(cb_ctx* c) {
  CB_API_BEGIN(c)
  if (!c->preprocessed) throw Error(CB_E_INVALID, "cb_build_overlap: call cb_preprocess first");
  c->overlapped = c->reduced = false;
  timer_start(c, "overlap");
  stage_build_overlap(c);
  timer_stop(c, "overlap");
  c->overlapped = true;
  CB_API_END(c)
}

int cb_string_graph(cb_ctx* c) {
  CB_API_BEGIN(c)
  if (!c->overlapped) throw Error(CB_E_INVALID, "cb_string_graph: call cb_build_overlap first");
  c->reduced = false;
  timer_start(c, "string");
  stage_string_graph(c);
  timer_stop(c, "string");
  c->reduced = true;
  CB_API_END(c)
}

int cb_get_counter(cb_ctx* c, const char* name, int64_t* value) {
  if (!c || !name || !value) return CB_E_INVALID;
  if (!strcmp(name, "pool_reserved_bytes")) { *value = (int64_t)c->arena.reserved; return CB_OK; }
  if (!strcmp(name, "pool_used_bytes")) { *value = (int64_t)c->arena.in_use; return CB_OK; }
  auto it = c->counters.find(name);
  if (it == c->counters.end()) {
    c->err = std::string("unknown counter ") + name;
    return CB_E_INVALID;
  }
  *value = it->second;
  return CB_OK;
}

int cb_export_edges(cb_ctx* c, int checkpoint, cb_edge* out, size_t cap, size_t* n) {
  CB_API_BEGIN(c)
  if (!n) throw Error(CB_E_INVALID, "null size pointer");
  const long long cnt = edge_count(c, checkpoint);
  if (cnt < 0) throw Error(CB_E_INVALID, "checkpoint not available");
  wait_node_table(c, true);  // node -> SFA line of every rank's nodes (sharded contexts gather it in the background)
  *n = (size_t)cnt;
  if (out && cap >= (size_t)cnt && cnt) export_edges_to(c, checkpoint, out);
  CB_API_END(c)
}

int cb_export_nodes(cb_ctx* c, cb_node* out, size_t cap, size_t* n) {
  CB_API_BEGIN(c)
  if (!n) throw Error(CB_E_INVALID, "null size pointer");
  if (!c->preprocessed) throw Error(CB_E_INVALID, "nodes not available");
  // a sharded context exports the nodes built from ITS shard of the SFA lines (the union over the
  // ranks, in rank order, is the node table sorted by line)
  const u32 lo = distributed(c) ? c->node_lo : 0, hi = distributed(c) ? c->node_hi : c->n_nodes;
  const u32 cnt = hi - lo;
  *n = cnt;
  if (out && cap >= cnt && cnt) {
    // While the replicated table is still being gathered (or its gather has not even been enqueued: that happens
    // collectively, in cb_build_overlap) the rank's own nodes are read from the gather's send buffers, so this
    // call never depends on the other ranks.
    const u32* lines = c->node_line.p + lo;
    const u32* covs = c->node_cov.p + lo;
    if (c->gather_pending_rest && c->own_line_tmp.p && c->own_cov.p) {
      lines = c->own_line_tmp.p;
      covs = c->own_cov.p;
    } else {
      wait_node_table(c, true);
    }
    DevBuf<cb_node> d;
    d.alloc(cnt, c->stream);
    export_nodes_kernel<<<div_up(cnt, 256), 256, 0, c->stream>>>(lines, covs, cnt, d.p);
    c->lc.n++;
    CB_CUDA(cudaMemcpyAsync(out, d.p, (size_t)cnt * sizeof(cb_node), cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
  }
  CB_API_END(c)
}

int cb_export_nodes_begin(cb_ctx* c, cb_node* out, size_t cap, size_t* n) {
  CB_API_BEGIN(c)
  if (!n) throw Error(CB_E_INVALID, "null size pointer");
  if (!c->preprocessed) throw Error(CB_E_INVALID, "nodes not available");
  const u32 lo = distributed(c) ? c->node_lo : 0, hi = distributed(c) ? c->node_hi : c->n_nodes;
  const u32 cnt = hi - lo;
  *n = cnt;
  if (out && cap >= cnt && cnt) {
    // (a sharded context whose node table is still being gathered reads its own nodes from the send buffers
    // instead of making the stream wait for the gather)
    const u32* lines = c->node_line.p + lo;
    const u32* covs = c->node_cov.p + lo;
    if (c->gather_pending_rest && c->own_line_tmp.p && c->own_cov.p) {
      lines = c->own_line_tmp.p;
      covs = c->own_cov.p;
    } else {
      wait_node_table(c, true);
    }
    // a transfer out of the staging buffer may still be running: it is reused only after that
    if (c->d2h_pending && c->stage_nodes.p) CB_CUDA(cudaStreamSynchronize(c->d2h_stream));
    c->stage_nodes.alloc(cnt, c->stream);
    export_nodes_kernel<<<div_up(cnt, 256), 256, 0, c->stream>>>(lines, covs, cnt, c->stage_nodes.p);
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
    d2h_begin(c, out, c->stage_nodes.p, (size_t)cnt * sizeof(cb_node));
  }
  CB_API_END(c)
}

int cb_export_edges_begin(cb_ctx* c, int checkpoint, cb_edge* out, size_t cap, size_t* n) {
  CB_API_BEGIN(c)
  if (!n) throw Error(CB_E_INVALID, "null size pointer");
  const long long cnt = edge_count(c, checkpoint);
  if (cnt < 0) throw Error(CB_E_INVALID, "checkpoint not available");
  *n = (size_t)cnt;
  if (out && cap >= (size_t)cnt && cnt) {
    wait_node_table(c, true);
    if (c->d2h_pending && c->stage_edges.p) CB_CUDA(cudaStreamSynchronize(c->d2h_stream));
    export_edges_dev(c, checkpoint, c->stage_edges);
    CB_CUDA(cudaGetLastError());
    d2h_begin(c, out, c->stage_edges.p, (size_t)cnt * sizeof(cb_edge));
  }
  CB_API_END(c)
}

int cb_export_wait(cb_ctx* c) {
  CB_API_BEGIN(c)
  if (c->d2h_pending) {
    CB_CUDA(cudaStreamSynchronize(c->d2h_stream));
    c->d2h_pending = false;
  }
  CB_API_END(c)
}

int cb_write_nodes(cb_ctx* c, int checkpoint, const char* out_dir) {
  CB_API_BEGIN(c)
  if (!out_dir) throw Error(CB_E_INVALID, "null path");
  if (!c->preprocessed) throw Error(CB_E_INVALID, "nodes not available");
  const bool with_edges = checkpoint != CB_CK_PREPROCESS;
  if (with_edges && edge_count(c, checkpoint) < 0) throw Error(CB_E_INVALID, "checkpoint not available");
  cudaStream_t s = c->stream;
  wait_node_table(c, true);
  const bool dist = distributed(c);
  const int world = c->cfg.world, rank = c->cfg.rank;
  // A sharded context writes the nodes built from ITS shard of the SFA lines as part-000<rank> (a Hadoop
  // output directory is multi-part anyway, BrushAssembler.java:75-89): the adjacency records of a node's two
  // lists may be owned by other ranks, so every rank routes the records it holds to the rank of the record's
  // node.  Read ids come from the replicated 16-byte id keys (ids longer than that are refused at preprocess).
  const u32 lo = dist ? c->node_lo : 0, hi = dist ? c->node_hi : c->n_nodes;
  const u32 nn = hi - lo;
  // export-encoded records of the nodes [lo, hi), sorted, on the device
  DevBuf<u64> enc, ka, kb;
  DevBuf<char> got;
  const u64* d_recs = nullptr;
  size_t n_recs = 0;
  if (with_edges) {
    export_encoded(c, checkpoint, enc);
    size_t n = (size_t)edge_count(c, checkpoint);
    const u64* src = enc.p;
    if (dist) {
      DevBuf<u64> bases;
      DevBuf<u32> dest;
      bases.alloc(world + 1, s);
      dest.alloc(n ? n : 1, s);
      CB_CUDA(cudaMemcpyAsync(bases.p, c->node_bases.data(), (world + 1) * sizeof(u64), cudaMemcpyHostToDevice, s));
      if (n) {
        node_owner_kernel<<<div_up(n, 256), 256, 0, s>>>(enc.p, n, c->lay.nb + c->lay.ob + 2, bases.p, world, dest.p);
        c->lc.n++;
      }
      size_t n_got = 0;
      route_records(c, enc.p, 8, dest.p, n, got, n_got);
      n = n_got;
      src = (const u64*)got.p;
    }
    if (n) {
      ka.alloc(n, s);
      kb.alloc(n, s);
      CB_CUDA(cudaMemcpyAsync(ka.p, src, n * 8, cudaMemcpyDeviceToDevice, s));
      const int w = radix_sort(ka.p, kb.p, nullptr, nullptr, n, 0, c->lay.total_bits(), s, c->lc, "export");
      d_recs = w ? kb.p : ka.p;
    }
    n_recs = n;
  }
  // the text of this rank's part, formatted on the device
  std::vector<char> text;
  {
    DevBuf<u32> first, len, off;
    DevBuf<ull> sum64;
    first.alloc((size_t)nn + 2, s);
    len.alloc((size_t)nn + 1, s);
    off.alloc((size_t)nn + 1, s);
    sum64.alloc(1, s);
    text_first_kernel<<<div_up((size_t)nn + 1, 256), 256, 0, s>>>(d_recs, n_recs, c->lay.nb + c->lay.ob + 2, lo, nn, first.p);
    CB_CUDA(cudaMemsetAsync(len.p + nn, 0, 4, s));
    CB_CUDA(cudaMemsetAsync(sum64.p, 0, sizeof(ull), s));
    if (nn) {
      text_nodes_kernel<false><<<div_up((size_t)nn, 256), 256, 0, s>>>(d_recs, first.p, lo, nn, c->lay, c->K, c->L, c->NW,
                                                                      c->node_seq.p, c->node_cov.p, c->node_idkey.p, len.p,
                                                                      nullptr, nullptr);
      int sb = div_up((size_t)nn, 256 * 8);
      if (sb > kNumSMs * 8) sb = kNumSMs * 8;
      sum_u32_kernel<<<sb, 256, 0, s>>>(len.p, nn, sum64.p);
    }
    exclusive_scan_u32(len.p, off.p, (size_t)nn + 1, s, c->lc);
    ull total64 = 0;
    CB_CUDA(cudaMemcpyAsync(&total64, sum64.p, sizeof(ull), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    // the offsets are 32-bit: a part of 4 GiB or more is refused (shard the run over more ranks)
    sync_error(c, total64 >= (1ull << 32) ? CB_E_CAPACITY : 0, "node text of one part is 4 GiB or more");
    const size_t total = (size_t)total64;
    DevBuf<char> dtext;
    dtext.alloc(total ? total : 1, s);
    if (nn)
      text_nodes_kernel<true><<<div_up((size_t)nn, 256), 256, 0, s>>>(d_recs, first.p, lo, nn, c->lay, c->K, c->L, c->NW,
                                                                     c->node_seq.p, c->node_cov.p, c->node_idkey.p, nullptr,
                                                                     off.p, dtext.p);
    c->lc.n += 4;
    CB_CUDA(cudaGetLastError());
    text.resize(total);
    if (total) CB_CUDA(cudaMemcpyAsync(text.data(), dtext.p, total, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
  }
  // FileSystem.delete(outputPath, true), then the part files (e.g. VerifyOverlap.java:379); rank 0 clears the
  // directory, the others wait for it
  std::string dir(out_dir);
  if (rank == 0) {
    DIR* d = opendir(dir.c_str());
    if (d) {
      while (dirent* de = readdir(d)) {
        std::string nm = de->d_name;
        if (nm == "." || nm == "..") continue;
        unlink((dir + "/" + nm).c_str());
      }
      closedir(d);
    }
    mkdir(dir.c_str(), 0777);
  }
  if (dist) {
    u64 token = 1;
    ex_allreduce(c, &token, 1, 0);  // barrier: the directory is clean
  }
  char part[32];
  snprintf(part, sizeof(part), "/part-%05d", rank);
  int io_err = 0;
  {
    FILE* f = fopen((dir + part).c_str(), "wb");
    if (!f) io_err = 1;
    else {
      if (!text.empty() && fwrite(text.data(), 1, text.size(), f) != text.size()) io_err = 1;
      if (fclose(f) != 0) io_err = 1;
    }
  }
  sync_error(c, io_err ? CB_E_IO : 0, "cannot write " + dir + part);
  CB_API_END(c)
}

int cb_write_fasta(cb_ctx* c, const char* out_dir) {
  CB_API_BEGIN(c)
  if (!out_dir) throw Error(CB_E_INVALID, "null path");
  if (!c->preprocessed) throw Error(CB_E_INVALID, "nodes not available");
  cudaStream_t s = c->stream;
  wait_node_table(c, true);
  const bool dist = distributed(c);
  const u32 lo = dist ? c->node_lo : 0, hi = dist ? c->node_hi : c->n_nodes;
  const u32 nn = hi - lo;
  std::vector<u32> cov(nn ? nn : 1);
  std::vector<u64> seq((size_t)(nn ? nn : 1) * c->NW);
  std::vector<IdKey> ids(nn ? nn : 1);
  if (nn) {
    CB_CUDA(cudaMemcpyAsync(cov.data(), c->node_cov.p + lo, (size_t)nn * 4, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaMemcpyAsync(seq.data(), c->node_seq.p + (size_t)lo * c->NW, (size_t)nn * c->NW * 8, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaMemcpyAsync(ids.data(), c->node_idkey.p + lo, (size_t)nn * sizeof(IdKey), cudaMemcpyDeviceToHost, s));
  }
  CB_CUDA(cudaStreamSynchronize(s));
  std::string dir(out_dir);
  if (c->cfg.rank == 0) {
    DIR* d = opendir(dir.c_str());
    if (d) {
      while (dirent* de = readdir(d)) {
        std::string nm = de->d_name;
        if (nm == "." || nm == "..") continue;
        unlink((dir + "/" + nm).c_str());
      }
      closedir(d);
    }
    mkdir(dir.c_str(), 0777);
  }
  if (dist) {
    u64 token = 1;
    ex_allreduce(c, &token, 1, 0);  // barrier: the directory is clean
  }
  char part[32];
  snprintf(part, sizeof(part), "/part-%05d", c->cfg.rank);
  int io_err = 0;
  FILE* f = fopen((dir + part).c_str(), "wb");
  if (!f) io_err = 1;
  std::string rec;
  const int L = c->L;
  for (u32 i = 0; i < nn && !io_err; i++) {
    // Graph2Fasta.java:49-75: ">id \t len=<L> \t cov=<Float.toString(cov)>", then the sequence in lines of 60
    rec.assign(">");
    char buf[16];
    const IdKey& k = ids[i];
    for (int b = 0; b < 8; b++) { buf[b] = (char)(k.hi >> (56 - 8 * b)); buf[8 + b] = (char)(k.lo >> (56 - 8 * b)); }
    size_t len = 0;
    while (len < 16 && buf[len]) len++;
    rec.append(buf, len);
    rec += "\tlen=" + std::to_string(L) + "\tcov=";
    if (cov[i] < 10000000u) rec += std::to_string(cov[i]) + ".0";  // Float.toString of an integral value below 1e7
    else {
      char e[32];
      snprintf(e, sizeof(e), "%.7gE7", (double)cov[i] / 1e7);  // (never reached by real inputs)
      rec += e;
    }
    rec += "\n";
    const u64* row = seq.data() + (size_t)i * c->NW;
    for (int b = 0; b < L; b++) {
      rec.push_back("ACGT"[(row[b >> 5] >> (62 - 2 * (b & 31))) & 3]);
      if ((b + 1) % 60 == 0 || b + 1 == L) rec.push_back('\n');
    }
    if (fwrite(rec.data(), 1, rec.size(), f) != rec.size()) io_err = 1;
  }
  if (f && fclose(f) != 0) io_err = 1;
  sync_error(c, io_err ? CB_E_IO : 0, "cannot write " + dir + part);
  CB_API_END(c)
}

int cb_get_time_ms(cb_ctx* c, const char* what, double* ms) {
  CB_API_BEGIN(c)
  if (!what || !ms) throw Error(CB_E_INVALID, "null argument");
  auto it = c->timers.find(what);
  if (it == c->timers.end() || !it->second.done) throw Error(CB_E_INVALID, std::string("no such timer: ") + what);
  CB_CUDA(cudaEventSynchronize(it->second.b));
  float f = 0;
  CB_CUDA(cudaEventElapsedTime(&f, it->second.a, it->second.b));
  *ms = f;
  CB_API_END(c)
}

int cb_debug_sort(cb_ctx* c, uint64_t* keys, uint32_t* vals, size_t n, int begin_bit, int end_bit) {
  CB_API_BEGIN(c)
  if (!keys && n) throw Error(CB_E_INVALID, "null keys");
  if (begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) throw Error(CB_E_INVALID, "bad bit range");
  cudaStream_t s = c->stream;
  const size_t n1 = n ? n : 1;
  DevBuf<u64> ka, kb;
  DevBuf<u32> va, vb;
  ka.alloc(n1, s);
  kb.alloc(n1, s);
  if (vals) { va.alloc(n1, s); vb.alloc(n1, s); }
  if (n) {
    CB_CUDA(cudaMemcpyAsync(ka.p, keys, n * 8, cudaMemcpyHostToDevice, s));
    if (vals) CB_CUDA(cudaMemcpyAsync(va.p, vals, n * 4, cudaMemcpyHostToDevice, s));
  }
  const int w = radix_sort(ka.p, kb.p, vals ? va.p : nullptr, vals ? vb.p : nullptr, n, begin_bit, end_bit, s, c->lc, "debug");
  if (n) {
    CB_CUDA(cudaMemcpyAsync(keys, w ? kb.p : ka.p, n * 8, cudaMemcpyDeviceToHost, s));
    if (vals) CB_CUDA(cudaMemcpyAsync(vals, w ? vb.p : va.p, n * 4, cudaMemcpyDeviceToHost, s));
  }
  CB_CUDA(cudaStreamSynchronize(s));
  CB_API_END(c)
}

int cb_debug_bucket_plan(size_t n_tuples, int32_t* b1, int32_t* b2, uint32_t* n_buckets, int32_t* fits) {
  if (!b1 || !b2 || !n_buckets || !fits) return CB_E_INVALID;
  const cb::BucketPlan pl = cb::make_bucket_plan(n_tuples);
  *b1 = pl.b1;
  *b2 = pl.b2;
  *n_buckets = pl.NB;
  *fits = pl.fits ? 1 : 0;
  return CB_OK;
}

int cb_debug_partition(cb_ctx* c, uint64_t* keys, uint32_t* vals, size_t n, int kmer_bits, uint32_t* bstart, size_t bstart_cap,
                       uint32_t* n_buckets, size_t* n_kept) {
  CB_API_BEGIN(c)
  if ((!keys || !vals) && n) throw Error(CB_E_INVALID, "null arrays");
  if (!n_buckets || !n_kept) throw Error(CB_E_INVALID, "null output");
  if (kmer_bits < 2 || kmer_bits > 64) throw Error(CB_E_INVALID, "bad k-mer width");
  cudaStream_t s = c->stream;
  const BucketPlan pl = make_bucket_plan(n);
  *n_buckets = pl.NB;
  if (!bstart || bstart_cap < (size_t)pl.NB + 1) { *n_kept = 0; return CB_OK; }  // size query
  const size_t n1 = n ? n : 1;
  DevBuf<u64> ka, kb;
  DevBuf<u32> va, vb, bs, tot;
  ka.alloc(n1, s); kb.alloc(n1, s); va.alloc(n1, s); vb.alloc(n1, s);
  bs.alloc((size_t)pl.NB + 1, s);
  tot.alloc(1, s);
  if (n) {
    CB_CUDA(cudaMemcpyAsync(ka.p, keys, n * 8, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(va.p, vals, n * 4, cudaMemcpyHostToDevice, s));
  }
  const u64 kmask = kmer_bits == 64 ? ~0ull : ((1ull << kmer_bits) - 1);
  bucket_partition(ka.p, va.p, kb.p, vb.p, n, kmask, pl, nullptr, bs.p, tot.p, s, c->lc);
  u32 kept = 0;
  CB_CUDA(cudaMemcpyAsync(&kept, tot.p, 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaMemcpyAsync(bstart, bs.p, ((size_t)pl.NB + 1) * 4, cudaMemcpyDeviceToHost, s));
  CB_CUDA(cudaStreamSynchronize(s));
  if (kept) {
    CB_CUDA(cudaMemcpyAsync(keys, ka.p, (size_t)kept * 8, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaMemcpyAsync(vals, va.p, (size_t)kept * 4, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
  }
  *n_kept = kept;
  CB_API_END(c)
}

int cb_get_launch_count(cb_ctx* c, int64_t* n) {
  if (!c || !n) return CB_E_INVALID;
  *n = c->lc.n;
  return CB_OK;
}

}  // extern "C"
