// dist.cu -- multi-GPU plumbing: device-side partitioning of records by destination rank and the
// calls into the host program's transport (cb_exchange, include/cloudbrush_gpu.h).
//
// This replaces Hadoop's HashPartitioner + shuffle (every JobClient.runJob of the path, e.g.
// MatchPrefix.java:468) with: partition on the device -> exchange byte counts -> one all-to-all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

#include "ctx.cuh"

namespace cb {

static void ex_fail(const char* what, int rc) {
  throw Error(CB_E_CUDA, std::string("exchange callback ") + what + " failed with " + std::to_string(rc));
}

// ---------------------------------------------------------------------------------------------
// Native transport: NCCL, located at run time (the library has no link-time dependency on it, so
// the single-GPU path and the CPU-side symbol tests load without NCCL).  Every transfer is
// enqueued on the context's stream: no host round trip per exchange.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

static std::string g_nccl_why;  // the loader's words when libnccl could not be used

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("CB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
    api.why = dlerror();
  }
  if (!api.lib) { g_nccl_why = api.why; return nullptr; }
#define CB_NCCL_SYM(field, name)                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name)); \
  if (!api.field) { api.why = std::string("missing symbol ") + name; g_nccl_why = api.why; api.lib = nullptr; return nullptr; }
  CB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  CB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  CB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  CB_NCCL_SYM(GroupStart, "ncclGroupStart")
  CB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  CB_NCCL_SYM(Send, "ncclSend")
  CB_NCCL_SYM(Recv, "ncclRecv")
  CB_NCCL_SYM(AllReduce, "ncclAllReduce")
  CB_NCCL_SYM(AllGather, "ncclAllGather")
  CB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CB_NCCL_SYM
  return &api;
}

static std::string nccl_load_error();
#define CB_NCCL(expr)                                                                                  \
  do {                                                                                                 \
    ncclResult_t r_ = (expr);                                                                          \
    if (r_ != ncclSuccess)                                                                             \
      throw Error(CB_E_CUDA, std::string(#expr) + ": " + nccl_api()->GetErrorString(r_));             \
  } while (0)

static std::string nccl_load_error() { return g_nccl_why.empty() ? std::string("libnccl.so.2 could not be loaded") : g_nccl_why; }

int nccl_unique_id(void* id128, std::string& err) {
  NcclApi* A = nccl_api();
  if (!A) {  // the loader's own words (dlopen / dlsym), so that a missing library can be diagnosed
    err = "NCCL is not available: " + nccl_load_error();
    return CB_E_UNSUPPORTED;
  }
  ncclUniqueId id;
  ncclResult_t r = A->GetUniqueId(&id);
  if (r != ncclSuccess) { err = A->GetErrorString(r); return CB_E_CUDA; }
  memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
  return CB_OK;
}

void nccl_init(cb_ctx* c, const void* id128) {
  NcclApi* A = nccl_api();
  if (!A) throw Error(CB_E_UNSUPPORTED, "NCCL is not available (libnccl.so.2 could not be loaded)");
  if (c->cfg.world < 2) throw Error(CB_E_INVALID, "cb_init_nccl needs cfg.world > 1");
  if (c->nccl_comm) nccl_shutdown(c);
  // the all-to-alls are grouped ncclSend/ncclRecv: with NCCL's default channel count per peer they
  // reach ~200 GB/s per GPU on 4 B200s, with 64 channels ~300 GB/s (measured; existing settings win)
  setenv("NCCL_MIN_P2P_NCHANNELS", "64", 0);
  setenv("NCCL_MAX_P2P_NCHANNELS", "64", 0);
  setenv("NCCL_MAX_NCHANNELS", "64", 0);
  ncclUniqueId id;
  memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  CB_NCCL(A->CommInitRank(&comm, c->cfg.world, id, c->cfg.rank));
  c->nccl_comm = comm;
  if (!c->side_stream) {
    CB_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_seq, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_rest, cudaEventDisableTiming));
  }
  const size_t W = (size_t)c->cfg.world;
  const size_t words = W * W * 4 + 4 * W + (size_t)DC_COUNT + 64;
  CB_CUDA(cudaMalloc(&c->ex_scratch, words * sizeof(u64)));
  c->ex_scratch_words = words;
}

void nccl_shutdown(cb_ctx* c) {
  if (c->nccl_comm && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
  if (c->side_stream) {
    cudaStreamSynchronize(c->side_stream);
    cudaStreamDestroy(c->side_stream);
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_seq);
    cudaEventDestroy(c->ev_rest);
    c->side_stream = nullptr;
  }
  if (c->ex_scratch) cudaFree(c->ex_scratch);
  c->ex_scratch = nullptr;
}

static inline ncclComm_t comm_of(cb_ctx* c) { return (ncclComm_t)c->nccl_comm; }
static inline cudaStream_t xs(cb_ctx* c) { return c->ex_stream ? c->ex_stream : c->stream; }

static void need_ex(cb_ctx* c) {
  if (!c->has_ex && !c->nccl_comm)
    throw Error(CB_E_INVALID, "cfg.world > 1 but no transport was set (cb_init_nccl or cb_set_exchange)");
}

void ex_allreduce(cb_ctx* c, u64* vals, int n, int op) {
  if (!distributed(c)) return;
  need_ex(c);
  if (c->nccl_comm) {
    if ((size_t)n > c->ex_scratch_words) throw Error(CB_E_INVALID, "ex_allreduce: too many values");
    NcclApi* A = nccl_api();
    CB_CUDA(cudaMemcpyAsync(c->ex_scratch, vals, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CB_NCCL(A->AllReduce(c->ex_scratch, c->ex_scratch, (size_t)n, ncclUint64, op == 0 ? ncclSum : ncclMax, comm_of(c),
                         c->stream));
    CB_CUDA(cudaMemcpyAsync(vals, c->ex_scratch, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    return;
  }
  int rc = c->ex.allreduce_u64(c->ex.user, (uint64_t*)vals, n, op);
  if (rc) ex_fail("allreduce_u64", rc);
}

void sync_error(cb_ctx* c, int local_code, const std::string& msg) {
  if (!distributed(c)) {
    if (local_code) throw Error(local_code, msg);
    return;
  }
  u64 v = local_code ? 1 : 0;
  ex_allreduce(c, &v, 1, 1);
  if (local_code) throw Error(local_code, msg);
  if (v) throw Error(CB_E_CAPACITY, "another rank of the sharded run failed (its cb_last_error has the reason)");
}

void ex_allreduce_max_u8(cb_ctx* c, uint8_t* dev, size_t n) {
  if (!distributed(c)) return;
  need_ex(c);
  if (c->nccl_comm) {
    CB_NCCL(nccl_api()->AllReduce(dev, dev, n, ncclUint8, ncclMax, comm_of(c), c->stream));
    return;
  }
  if (!c->ex.allreduce_max_u8_dev) throw Error(CB_E_INVALID, "cb_exchange.allreduce_max_u8_dev is not set");
  CB_CUDA(cudaStreamSynchronize(c->stream));
  int rc = c->ex.allreduce_max_u8_dev(c->ex.user, dev, (uint64_t)n);
  if (rc) ex_fail("allreduce_max_u8_dev", rc);
}

void ex_counts(cb_ctx* c, const u64* send, u64* recv, int n_per_rank) {
  need_ex(c);
  if (c->nccl_comm) {
    // every rank's whole send vector is gathered; this rank's column is picked out on the host
    const size_t W = (size_t)c->cfg.world, m = W * (size_t)n_per_rank;
    if (m + W * m > c->ex_scratch_words) throw Error(CB_E_INVALID, "ex_counts: too many values");
    u64* d_send = c->ex_scratch;
    u64* d_all = c->ex_scratch + m;
    CB_CUDA(cudaMemcpyAsync(d_send, send, m * 8, cudaMemcpyHostToDevice, c->stream));
    CB_NCCL(nccl_api()->AllGather(d_send, d_all, m, ncclUint64, comm_of(c), c->stream));
    std::vector<u64> all(W * m);
    CB_CUDA(cudaMemcpyAsync(all.data(), d_all, W * m * 8, cudaMemcpyDeviceToHost, c->stream));
    CB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t r = 0; r < W; r++)
      for (int j = 0; j < n_per_rank; j++)
        recv[r * n_per_rank + j] = all[r * m + (size_t)c->cfg.rank * n_per_rank + j];
    return;
  }
  int rc = c->ex.exchange_counts(c->ex.user, (const uint64_t*)send, (uint64_t*)recv, n_per_rank);
  if (rc) ex_fail("exchange_counts", rc);
}

void ex_alltoallv(cb_ctx* c, const void* send, const u64* send_bytes, void* recv, const u64* recv_bytes) {
  need_ex(c);
  if (c->nccl_comm) {
    NcclApi* A = nccl_api();
    const int W = c->cfg.world, me = c->cfg.rank;
    size_t so = 0, ro = 0;
    CB_NCCL(A->GroupStart());
    for (int r = 0; r < W; r++) {
      const char* sp = (const char*)send + so;
      char* rp = (char*)recv + ro;
      if (r == me) {
        if (send_bytes[r]) CB_CUDA(cudaMemcpyAsync(rp, sp, send_bytes[r], cudaMemcpyDeviceToDevice, xs(c)));
      } else {
        if (send_bytes[r]) CB_NCCL(A->Send(sp, send_bytes[r], ncclUint8, r, comm_of(c), xs(c)));
        if (recv_bytes[r]) CB_NCCL(A->Recv(rp, recv_bytes[r], ncclUint8, r, comm_of(c), xs(c)));
      }
      so += send_bytes[r];
      ro += recv_bytes[r];
    }
    CB_NCCL(A->GroupEnd());
    return;
  }
  CB_CUDA(cudaStreamSynchronize(xs(c)));
  int rc = c->ex.alltoallv(c->ex.user, send, (const uint64_t*)send_bytes, recv, (const uint64_t*)recv_bytes);
  if (rc) ex_fail("alltoallv", rc);
}

void ex_allgatherv(cb_ctx* c, const void* send, u64 send_bytes, void* recv, const u64* recv_bytes) {
  need_ex(c);
  if (c->nccl_comm) {
    NcclApi* A = nccl_api();
    const int W = c->cfg.world, me = c->cfg.rank;
    size_t ro = 0;
    CB_NCCL(A->GroupStart());
    for (int r = 0; r < W; r++) {
      char* rp = (char*)recv + ro;
      if (r == me) {
        if (send_bytes) CB_CUDA(cudaMemcpyAsync(rp, send, send_bytes, cudaMemcpyDeviceToDevice, xs(c)));
      } else {
        if (send_bytes) CB_NCCL(A->Send(send, send_bytes, ncclUint8, r, comm_of(c), xs(c)));
        if (recv_bytes[r]) CB_NCCL(A->Recv(rp, recv_bytes[r], ncclUint8, r, comm_of(c), xs(c)));
      }
      ro += recv_bytes[r];
    }
    CB_NCCL(A->GroupEnd());
    return;
  }
  CB_CUDA(cudaStreamSynchronize(xs(c)));
  int rc = c->ex.allgatherv(c->ex.user, send, send_bytes, recv, (const uint64_t*)recv_bytes);
  if (rc) ex_fail("allgatherv", rc);
}

void ex_alltoallv2(cb_ctx* c, const void* send_a, void* recv_a, size_t elem_a, const void* send_b, void* recv_b,
                   size_t elem_b, const u64* send_cnt, const u64* recv_cnt) {
  need_ex(c);
  const int W = c->cfg.world, me = c->cfg.rank;
  if (c->nccl_comm) {
    NcclApi* A = nccl_api();
    size_t so = 0, ro = 0;
    CB_NCCL(A->GroupStart());
    for (int r = 0; r < W; r++) {
      const void* sp[2] = {(const char*)send_a + so * elem_a, (const char*)send_b + so * elem_b};
      void* rp[2] = {(char*)recv_a + ro * elem_a, (char*)recv_b + ro * elem_b};
      const size_t el[2] = {elem_a, elem_b};
      for (int q = 0; q < 2; q++) {
        if (r == me) {
          if (send_cnt[r]) CB_CUDA(cudaMemcpyAsync(rp[q], sp[q], send_cnt[r] * el[q], cudaMemcpyDeviceToDevice, xs(c)));
        } else {
          if (send_cnt[r]) CB_NCCL(A->Send(sp[q], send_cnt[r] * el[q], ncclUint8, r, comm_of(c), xs(c)));
          if (recv_cnt[r]) CB_NCCL(A->Recv(rp[q], recv_cnt[r] * el[q], ncclUint8, r, comm_of(c), xs(c)));
        }
      }
      so += send_cnt[r];
      ro += recv_cnt[r];
    }
    CB_NCCL(A->GroupEnd());
    return;
  }
  std::vector<u64> sb(W), rb(W);
  for (int r = 0; r < W; r++) { sb[r] = send_cnt[r] * elem_a; rb[r] = recv_cnt[r] * elem_a; }
  ex_alltoallv(c, send_a, sb.data(), recv_a, rb.data());
  for (int r = 0; r < W; r++) { sb[r] = send_cnt[r] * elem_b; rb[r] = recv_cnt[r] * elem_b; }
  ex_alltoallv(c, send_b, sb.data(), recv_b, rb.data());
}

void wait_node_table(cb_ctx* c, bool rest) {
  if (c->gather_pending_seq) {
    CB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_seq, 0));
    c->gather_pending_seq = false;
  }
  if (rest && c->gather_pending_rest) {
    CB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rest, 0));
    c->gather_pending_rest = false;
  }
}

void reduce_counters(cb_ctx* c) {
  for (int i = 0; i < DC_COUNT; i++) c->gcount[i] = c->hcount[i];
  if (!distributed(c)) return;
  ex_allreduce(c, (u64*)c->gcount, DC_COUNT, 0);
  u64 mx[2] = {c->hcount[DC_MAX_COV], c->hcount[DC_MAX_LIST]};
  ex_allreduce(c, mx, 2, 1);
  c->gcount[DC_MAX_COV] = mx[0];
  c->gcount[DC_MAX_LIST] = mx[1];
}

__global__ void __launch_bounds__(256) dest_keys_kernel(const u32* __restrict__ dest, size_t n,
                                                        u64* __restrict__ keys, u32* __restrict__ vals) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = dest[i]; vals[i] = (u32)i; }
}

// counts[d] = number of keys equal to d (d <= world), block-local histogram first
__global__ void __launch_bounds__(256) dest_hist_kernel(const u64* __restrict__ keys, size_t n, int nbins,
                                                        ull* __restrict__ counts) {
  __shared__ u32 h[257];
  for (int t = threadIdx.x; t < nbins; t += blockDim.x) h[t] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u32 d = (u32)keys[i];
    if ((int)d < nbins) atomicAdd(&h[d], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nbins; t += blockDim.x)
    if (h[t]) atomicAdd(&counts[t], (ull)h[t]);
}

__global__ void __launch_bounds__(256) gather_records_kernel(const u32* __restrict__ recs, int words,
                                                             const u32* __restrict__ perm, size_t n,
                                                             u32* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = n * (size_t)words;
  if (i >= total) return;
  size_t r = i / words;
  int w = (int)(i - r * words);
  out[i] = recs[(size_t)perm[r] * words + w];
}

void route_records(cb_ctx* c, const void* recs, size_t elem_bytes, const u32* dest, size_t n, DevBuf<char>& out,
                   size_t& n_out) {
  cudaStream_t s = c->stream;
  const int world = c->cfg.world;
  const int words = (int)(elem_bytes / 4);
  if (elem_bytes % 4) throw Error(CB_E_INVALID, "route_records: record size must be a multiple of 4");
  std::vector<u64> cnt(world + 1, 0), send_bytes(world), recv_bytes(world);
  DevBuf<char> send;
  guarded(c, [&] {  // every rank reaches the exchange below, whatever happens locally
    if (!n) { send.alloc(elem_bytes, s); return; }
    DevBuf<u64> ka, kb;
    DevBuf<u32> va, vb;
    DevBuf<u64> dcnt;
    ka.alloc(n, s); kb.alloc(n, s); va.alloc(n, s); vb.alloc(n, s);
    dcnt.alloc(world + 1, s);
    CB_CUDA(cudaMemsetAsync(dcnt.p, 0, (world + 1) * sizeof(u64), s));
    dest_keys_kernel<<<div_up(n, 256), 256, 0, s>>>(dest, n, ka.p, va.p);
    int bits = 1;
    while ((1 << bits) < world + 1) bits++;
    int w = radix_sort(ka.p, kb.p, va.p, vb.p, n, 0, bits, s, c->lc, "route");
    int blocks = div_up(n, 256 * 16);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    dest_hist_kernel<<<blocks, 256, 0, s>>>(w ? kb.p : ka.p, n, world + 1, (ull*)dcnt.p);
    CB_CUDA(cudaMemcpyAsync(cnt.data(), dcnt.p, (world + 1) * sizeof(u64), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    size_t n_send = n - (size_t)cnt[world];
    send.alloc((n_send ? n_send : 1) * elem_bytes, s);
    if (n_send)
      gather_records_kernel<<<div_up(n_send * words, 256), 256, 0, s>>>((const u32*)recs, words, w ? vb.p : va.p,
                                                                       n_send, (u32*)send.p);
    c->lc.n += 3;
    CB_CUDA(cudaGetLastError());
  });
  std::vector<u64> rcnt(world);
  ex_counts(c, cnt.data(), rcnt.data(), 1);
  n_out = 0;
  for (int r = 0; r < world; r++) {
    send_bytes[r] = cnt[r] * elem_bytes;
    recv_bytes[r] = rcnt[r] * elem_bytes;
    n_out += (size_t)rcnt[r];
  }
  out.alloc((n_out ? n_out : 1) * elem_bytes, s);
  ex_alltoallv(c, send.p, send_bytes.data(), out.p, recv_bytes.data());
}

void gather_all(cb_ctx* c, const void* local, size_t elem_bytes, size_t count, const std::vector<u64>& counts,
                void* out) {
  std::vector<u64> rb(c->cfg.world);
  for (int r = 0; r < c->cfg.world; r++) rb[r] = counts[r] * elem_bytes;
  ex_allgatherv(c, local, (u64)count * elem_bytes, out, rb.data());
}

}  // namespace cb
