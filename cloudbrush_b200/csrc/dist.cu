// dist.cu -- multi-GPU plumbing: device-side partitioning of records by destination rank and the
// calls into the host program's transport (cb_exchange, include/cloudbrush_gpu.h).
//
// This replaces Hadoop's HashPartitioner + shuffle (every JobClient.runJob of the path, e.g.
// MatchPrefix.java:468) with: partition on the device -> exchange byte counts -> one all-to-all.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <algorithm>
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
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;  // optional (NCCL >= 2.18)
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
  api.CommSplit = reinterpret_cast<decltype(api.CommSplit)>(dlsym(api.lib, "ncclCommSplit"));
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
  static const bool one_comm = getenv("CB_ONE_COMM") != nullptr;  // experiment knob
  if (A->CommSplit && !one_comm) {
    ncclComm_t side = nullptr;
    if (A->CommSplit(comm, 0, c->cfg.rank, &side, nullptr) == ncclSuccess && side) c->nccl_comm_side = side;
  }
  if (!c->side_stream) {
    CB_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_seq, cudaEventDisableTiming));
    CB_CUDA(cudaEventCreateWithFlags(&c->ev_rest, cudaEventDisableTiming));
  }
  const size_t W = (size_t)c->cfg.world;
  const size_t words = (W + 1) * (4 * W + (size_t)DC_COUNT + 64);  // a gather of that many values per rank fits
  CB_CUDA(cudaMalloc(&c->ex_scratch, words * sizeof(u64)));
  c->ex_scratch_words = words;
}

void nccl_shutdown(cb_ctx* c) {
  if (c->nccl_comm_side && nccl_api()) nccl_api()->CommDestroy((ncclComm_t)c->nccl_comm_side);
  c->nccl_comm_side = nullptr;
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
  p2p_release(c);
  if (c->ex_scratch) cudaFree(c->ex_scratch);
  c->ex_scratch = nullptr;
}

static inline ncclComm_t comm_of(cb_ctx* c) { return (ncclComm_t)c->nccl_comm; }
// the communicator of the transfers enqueued on xs(c): the second one while they go to the side stream
static inline ncclComm_t comm_x(cb_ctx* c) {
  return (ncclComm_t)((c->ex_stream && c->ex_stream == c->side_stream && c->nccl_comm_side) ? c->nccl_comm_side : c->nccl_comm);
}
static inline cudaStream_t xs(cb_ctx* c) { return c->ex_stream ? c->ex_stream : c->stream; }

static void need_ex(cb_ctx* c) {
  if (!c->has_ex && !c->nccl_comm)
    throw Error(CB_E_INVALID, "cfg.world > 1 but no transport was set (cb_init_nccl or cb_set_exchange)");
}

void ex_allreduce(cb_ctx* c, u64* vals, int n, int op) {
  if (!distributed(c)) return;
  need_ex(c);
  c->n_host_coll++;
  if (c->nccl_comm) {
    if ((size_t)n > c->ex_scratch_words) throw Error(CB_E_INVALID, "ex_allreduce: too many values");
    NcclApi* A = nccl_api();
    // (the few words travel through mapped pinned memory, not the copy engines: primitives.cuh, ReadBack)
    c->lc.rb.put(c->ex_scratch, vals, (size_t)n * 8, c->stream);
    CB_NCCL(A->AllReduce(c->ex_scratch, c->ex_scratch, (size_t)n, ncclUint64, op == 0 ? ncclSum : ncclMax, comm_of(c),
                         c->stream));
    c->lc.rb.get(vals, c->ex_scratch, (size_t)n * 8, c->stream);
    c->lc.rb.sync(c->stream);
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

void defer_error(cb_ctx* c, int code, const std::string& msg) {
  if (!code) return;
  if (!distributed(c)) throw Error(code, msg);
  if (!c->carried_code) { c->carried_code = code; c->carried_msg = msg; }
}

Gathered ex_gather_checked(cb_ctx* c, const u64* mine, int n) {
  Gathered G;
  G.world = c->cfg.world;
  G.n = n;
  if (!distributed(c)) {
    G.v.assign(mine, mine + n);
    return G;
  }
  const int code = c->carried_code;
  const std::string msg = c->carried_msg;
  c->carried_code = 0;
  c->carried_msg.clear();
  need_ex(c);
  c->n_host_coll++;
  const size_t W = (size_t)G.world, m = (size_t)n + 1;  // the last word is the failure flag
  std::vector<u64> all(W * m);
  if (c->nccl_comm) {
    if (m + W * m > c->ex_scratch_words) throw Error(CB_E_INVALID, "ex_gather_checked: too many values");
    // (the few words travel through mapped pinned memory, not the copy engines -- a copy would queue behind the
    // input prefetch / result transfers of a pipelined caller: primitives.cuh, ReadBack)
    std::vector<u64> snd(mine, mine + n);
    snd.push_back(code ? 1 : 0);
    c->lc.rb.put(c->ex_scratch, snd.data(), m * 8, c->stream);
    CB_NCCL(nccl_api()->AllGather(c->ex_scratch, c->ex_scratch + m, m, ncclUint64, comm_of(c), c->stream));
    c->lc.rb.get(all.data(), c->ex_scratch + m, W * m * 8, c->stream);
    c->lc.rb.sync(c->stream);
  } else {
    // the callbacks have an all-to-all of counts: the same m values go to every rank
    std::vector<u64> snd(W * m);
    for (size_t r = 0; r < W; r++) {
      for (int j = 0; j < n; j++) snd[r * m + j] = mine[j];
      snd[r * m + n] = code ? 1 : 0;
    }
    int rc = c->ex.exchange_counts(c->ex.user, (const uint64_t*)snd.data(), (uint64_t*)all.data(), (int)m);
    if (rc) ex_fail("exchange_counts", rc);
  }
  bool other = false;
  G.v.resize(W * (size_t)n);
  for (size_t r = 0; r < W; r++) {
    for (int j = 0; j < n; j++) G.v[r * n + j] = all[r * m + j];
    if (all[r * m + n]) other = true;
  }
  if (code) throw Error(code, msg);
  if (other) throw Error(CB_E_CAPACITY, "another rank of the sharded run failed (its cb_last_error has the reason)");
  return G;
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
        if (send_bytes[r]) CB_NCCL(A->Send(sp, send_bytes[r], ncclUint8, r, comm_x(c), xs(c)));
        if (recv_bytes[r]) CB_NCCL(A->Recv(rp, recv_bytes[r], ncclUint8, r, comm_x(c), xs(c)));
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
        if (send_bytes) CB_NCCL(A->Send(send, send_bytes, ncclUint8, r, comm_x(c), xs(c)));
        if (recv_bytes[r]) CB_NCCL(A->Recv(rp, recv_bytes[r], ncclUint8, r, comm_x(c), xs(c)));
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
          if (send_cnt[r]) CB_NCCL(A->Send(sp[q], send_cnt[r] * el[q], ncclUint8, r, comm_x(c), xs(c)));
          if (recv_cnt[r]) CB_NCCL(A->Recv(rp[q], recv_cnt[r] * el[q], ncclUint8, r, comm_x(c), xs(c)));
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

// ---------------------------------------------------------------------------------------------
// One-sided exchange over NVLink.  NCCL's grouped send/recv all-to-all reaches about 300 GB/s per GPU
// here (it stages through its own FIFOs); the exchange below is one kernel of plain coalesced stores from
// the send buffer into the peers' mailboxes, every peer at once, followed by a stream-ordered barrier.
// Safety of the mailbox: a rank writes into a peer's mailbox only after a host-synchronous gather of the
// same exchange (the count matrix), which the peer enters in stream order after everything that read
// the mailbox's previous contents.
// ---------------------------------------------------------------------------------------------
static void p2p_unmap(cb_ctx* c) {
  for (int r = 0; r < (int)c->p2p_peer.size(); r++)
    if (r != c->cfg.rank && c->p2p_peer[r]) {
      if (cudaIpcCloseMemHandle(c->p2p_peer[r]) != cudaSuccess) cudaGetLastError();
    }
  c->p2p_peer.clear();
}

void p2p_release(cb_ctx* c) {
  p2p_unmap(c);
  if (c->p2p_box) cudaFree(c->p2p_box);
  c->p2p_box = nullptr;
  c->p2p_cap = 0;
  c->p2p_on = false;
  c->p2p_tried = false;
}

// collective: allocates this rank's mailbox of `cap` bytes and maps every peer's; false (on every rank) if any step
// fails on any rank
static bool p2p_map(cb_ctx* c, size_t cap) {
  const int W = c->cfg.world, me = c->cfg.rank;
  u64 info[11] = {0};
  char host[256] = {0};
  gethostname(host, sizeof(host) - 1);
  u64 hh = 1469598103934665603ull;
  for (const char* q = host; *q; q++) hh = (hh ^ (unsigned char)*q) * 1099511628211ull;
  int dev = 0;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  bool ok = cudaGetDevice(&dev) == cudaSuccess;
  if (ok && cudaMalloc(&c->p2p_box, cap) != cudaSuccess) { cudaGetLastError(); c->p2p_box = nullptr; ok = false; }
  if (ok && cudaIpcGetMemHandle(&h, c->p2p_box) != cudaSuccess) { cudaGetLastError(); ok = false; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  info[0] = ok ? 1 : 0;
  info[1] = hh;
  info[2] = (u64)dev;
  memcpy(info + 3, &h, 64);
  const Gathered G = ex_gather_checked(c, info, 11);
  bool all = true;
  for (int r = 0; r < W; r++) {
    if (!G.at(r, 0) || G.at(r, 1) != hh) all = false;
    for (int q = 0; q < r; q++)
      if (G.at(q, 2) == G.at(r, 2)) all = false;  // two ranks on one device: not the layout this path is for
  }
  u64 ok2 = all ? 1 : 0;
  if (all) {
    c->p2p_peer.assign(W, nullptr);
    c->p2p_peer[me] = c->p2p_box;
    for (int r = 0; r < W && ok2; r++) {
      if (r == me) continue;
      cudaIpcMemHandle_t hr;
      u64 words[8];
      for (int j = 0; j < 8; j++) words[j] = G.at(r, 3 + j);
      memcpy(&hr, words, 64);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok2 = 0;
      } else {
        c->p2p_peer[r] = (char*)ptr;
      }
    }
  }
  const Gathered G2 = ex_gather_checked(c, &ok2, 1);
  bool every = true;
  for (int r = 0; r < W; r++)
    if (!G2.at(r, 0)) every = false;
  if (!every) {
    p2p_unmap(c);
    // (the peers may still hold this mailbox mapped until they pass the same point: one more gather orders the free)
    u64 z = 0;
    ex_gather_checked(c, &z, 1);
    if (c->p2p_box) cudaFree(c->p2p_box);
    c->p2p_box = nullptr;
    return false;
  }
  c->p2p_cap = cap;
  return true;
}

// collective (the same decision on every rank: need_max is computed from gathered data)
static bool p2p_ensure(cb_ctx* c, u64 need_max) {
  static const bool off = getenv("CB_NO_P2P") != nullptr;
  if (off || !c->nccl_comm || c->cfg.world > 16) return false;
  if (c->p2p_tried && !c->p2p_on) return false;
  if (c->p2p_on && c->p2p_cap >= need_max) return true;
  size_t cap = (size_t)need_max + (size_t)need_max / 8 + (1u << 20);
  cap = (cap + (2u << 20) - 1) / (2u << 20) * (2u << 20);
  if (c->p2p_on) {  // grow: nobody may still have the old mailbox mapped when it is freed
    CB_CUDA(cudaStreamSynchronize(c->stream));
    p2p_unmap(c);
    u64 z = 0;
    ex_gather_checked(c, &z, 1);
    cudaFree(c->p2p_box);
    c->p2p_box = nullptr;
    c->p2p_on = false;
  }
  c->p2p_tried = true;
  c->p2p_on = p2p_map(c, cap);
  return c->p2p_on;
}

constexpr int kPutSegs = 48;  // 16 ranks x 3 arrays
struct PutTable {
  int n;
  unsigned long long chunks_max;
  const char* src[kPutSegs];
  char* dst[kPutSegs];
  unsigned long long bytes[kPutSegs];
};
constexpr u64 kPutChunk = 64 << 10;

// Chunk v of the exchange is chunk v / n of segment v % n: consecutive blocks store to different peers, so every
// NVLink destination is busy from the first wave on.  The access width is what source, destination and length allow.
template <class V>
__device__ __forceinline__ void put_chunk(const char* __restrict__ src, char* __restrict__ dst, u64 bytes) {
  const V* s = reinterpret_cast<const V*>(src);
  V* d = reinterpret_cast<V*>(dst);
  const u64 n = bytes / sizeof(V);
  u64 i = threadIdx.x;
  for (; i + 3 * (u64)blockDim.x < n; i += 4 * (u64)blockDim.x) {
    const V a = s[i], b = s[i + blockDim.x], e = s[i + 2 * (u64)blockDim.x], f = s[i + 3 * (u64)blockDim.x];
    d[i] = a;
    d[i + blockDim.x] = b;
    d[i + 2 * (u64)blockDim.x] = e;
    d[i + 3 * (u64)blockDim.x] = f;
  }
  for (; i < n; i += blockDim.x) d[i] = s[i];
}
__global__ void __launch_bounds__(512) p2p_put_kernel(const PutTable T) {
  const unsigned long long total = T.chunks_max * (unsigned long long)T.n;
  for (unsigned long long v = blockIdx.x; v < total; v += gridDim.x) {
    const int sg = (int)(v % (unsigned long long)T.n);
    const u64 off = (v / (unsigned long long)T.n) * kPutChunk;
    if (off >= T.bytes[sg]) continue;
    const u64 len = T.bytes[sg] - off < kPutChunk ? T.bytes[sg] - off : kPutChunk;
    const char* src = T.src[sg] + off;
    char* dst = T.dst[sg] + off;
    const u64 al = (u64)(uintptr_t)src | (u64)(uintptr_t)dst | len;
    if ((al & 15) == 0) put_chunk<uint4>(src, dst, len);
    else if ((al & 7) == 0) put_chunk<uint2>(src, dst, len);
    else put_chunk<u32>(src, dst, len);  // every record of the path is a multiple of 4 bytes
  }
}

bool p2p_exchange(cb_ctx* c, const Gathered& M, int narr, const void* const* send, const size_t* elem, void** recv) {
  if (!distributed(c) || !c->nccl_comm) return false;
  const int W = c->cfg.world, me = c->cfg.rank;
  if (W * narr > kPutSegs) return false;
  auto align = [](u64 b) { return (b + 255) & ~255ull; };
  // what every rank receives, and the mailbox bytes that takes
  std::vector<u64> trecv(W, 0);
  u64 need_max = 0;
  for (int d = 0; d < W; d++) {
    for (int q = 0; q < W; q++) trecv[d] += M.at(q, d);
    u64 need = 0;
    for (int a = 0; a < narr; a++) need += align(trecv[d] * elem[a]);
    if (need > need_max) need_max = need;
  }
  if (!p2p_ensure(c, need_max)) return false;
  std::vector<u64> so(W + 1, 0);  // elements before destination d in the send arrays
  for (int d = 0; d < W; d++) so[d + 1] = so[d] + M.at(me, d);
  PutTable T;
  T.n = 0;
  T.chunks_max = 0;
  for (int i = 0; i < W; i++) {
    const int d = (me + 1 + i) % W;  // the ranks do not all start on rank 0; this rank's own part goes last
    const u64 cnt = M.at(me, d);
    u64 before = 0;  // elements the ranks below this one put into d's mailbox
    for (int q = 0; q < me; q++) before += M.at(q, d);
    u64 base = 0;
    for (int a = 0; a < narr; a++) {
      if (cnt) {
        T.src[T.n] = (const char*)send[a] + so[d] * elem[a];
        T.dst[T.n] = c->p2p_peer[d] + base + before * elem[a];
        T.bytes[T.n] = cnt * elem[a];
        const u64 ch = (T.bytes[T.n] + kPutChunk - 1) / kPutChunk;
        if (ch > T.chunks_max) T.chunks_max = ch;
        T.n++;
      }
      base += align(trecv[d] * elem[a]);
    }
  }
  {
    u64 base = 0;
    for (int a = 0; a < narr; a++) {
      recv[a] = c->p2p_box + base;
      base += align(trecv[me] * elem[a]);
    }
  }
  if (T.n) {
    static const int per_sm = getenv("CB_P2P_BLOCKS") ? atoi(getenv("CB_P2P_BLOCKS")) : 4;  // experiment knobs
    static const bool use_ce = getenv("CB_P2P_CE") != nullptr;
    const unsigned long long total = T.chunks_max * (unsigned long long)T.n;
    int grid = kNumSMs * (per_sm > 0 ? per_sm : 4);
    if ((unsigned long long)grid > total) grid = (int)total;
    c->lc.begin("p2p_put");
    if (use_ce) {  // the copy engines instead of the SMs
      for (int j = 0; j < T.n; j++)
        CB_CUDA(cudaMemcpyAsync(T.dst[j], T.src[j], T.bytes[j], cudaMemcpyDeviceToDevice, c->stream));
    } else {
      p2p_put_kernel<<<grid, 512, 0, c->stream>>>(T);
    }
    c->lc.end("p2p_put");
    c->lc.n++;
    CB_CUDA(cudaGetLastError());
  }
  // barrier in stream order: when it completes here, every rank's put kernel has completed (its stores have landed)
  u64* bar = c->ex_scratch + c->ex_scratch_words - 1;
  CB_NCCL(nccl_api()->AllReduce(bar, bar, 1, ncclUint64, ncclSum, comm_of(c), c->stream));
  return true;
}

void launch_node_gather(cb_ctx* c) {
  if (!c->gather_deferred) return;
  c->gather_deferred = false;
  const size_t n_local = (size_t)(c->node_hi - c->node_lo);
  CB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  CB_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
  c->ex_stream = c->side_stream;
  struct Unset {
    cb_ctx* c;
    ~Unset() { c->ex_stream = nullptr; }
  } unset{c};
  gather_all(c, c->own_seq.p, (size_t)c->NW * 8, n_local, c->gather_counts, c->node_seq.p);
  gather_all(c, c->own_cov.p, 4, n_local, c->gather_counts, c->node_cov.p);
  CB_CUDA(cudaEventRecord(c->ev_seq, c->side_stream));
  gather_all(c, c->own_line_tmp.p, 4, n_local, c->gather_counts, c->node_line.p);
  gather_all(c, c->own_id_tmp.p, sizeof(IdKey), n_local, c->gather_counts, c->node_idkey.p);
  CB_CUDA(cudaEventRecord(c->ev_rest, c->side_stream));
}

void wait_node_table(cb_ctx* c, bool rest) {
  launch_node_gather(c);  // (nobody has started it yet: every rank reaches this point in the same call)
  if (c->gather_pending_seq) {
    CB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_seq, 0));
    c->gather_pending_seq = false;
  }
  if (rest && c->gather_pending_rest) {
    CB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rest, 0));
    c->gather_pending_rest = false;
  }
}

Gathered reduce_counters(cb_ctx* c, const u64* extra, int n_extra) {
  std::vector<u64> mine(DC_COUNT + n_extra);
  for (int i = 0; i < DC_COUNT; i++) mine[i] = c->hcount[i];
  for (int j = 0; j < n_extra; j++) mine[DC_COUNT + j] = extra[j];
  Gathered G = ex_gather_checked(c, mine.data(), DC_COUNT + n_extra);
  for (int i = 0; i < DC_COUNT; i++) c->gcount[i] = G.sum(i);
  c->gcount[DC_MAX_COV] = G.max(DC_MAX_COV);
  c->gcount[DC_MAX_LIST] = G.max(DC_MAX_LIST);
  return G;
}

__global__ void __launch_bounds__(256) dest_keys_kernel(const u32* __restrict__ dest, size_t n,
                                                        u64* __restrict__ keys, u32* __restrict__ vals) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = dest[i]; vals[i] = (u32)i; }
}

// counts[d] = number of keys whose digit (key >> shift) & mask equals d, for keys SORTED by that digit: one
// thread per bin finds its bounds by bisection.  (A histogram with a shared-memory atomic per key took 1 ms for
// the 69 M tuples of config 2: a handful of bins means every lane of a warp hits the same few addresses.)
__global__ void __launch_bounds__(256) sorted_digit_counts_kernel(const u64* __restrict__ keys, size_t n, int shift,
                                                                  u32 mask, int nbins, ull* __restrict__ counts) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nbins) return;
  size_t bound[2];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const u32 want = (u32)d + q;  // first key with digit >= want
    size_t a = 0, b = n;
    while (a < b) {
      const size_t m = a + ((b - a) >> 1);
      if (((u32)(keys[m] >> shift) & mask) < want) a = m + 1;
      else b = m;
    }
    bound[q] = a;
  }
  counts[d] = (ull)(bound[1] - bound[0]);
}

void sorted_digit_counts(const u64* keys, size_t n, int shift, u32 mask, int nbins, u64* counts, cudaStream_t s) {
  sorted_digit_counts_kernel<<<div_up((size_t)nbins, 256), 256, 0, s>>>(keys, n, shift, mask, nbins, (ull*)counts);
  CB_CUDA(cudaGetLastError());
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
                   size_t& n_out, u64* piggy, int n_piggy, bool big) {
  cudaStream_t s = c->stream;
  const int world = c->cfg.world;
  const int words = (int)(elem_bytes / 4);
  if (elem_bytes % 4) throw Error(CB_E_INVALID, "route_records: record size must be a multiple of 4");
  std::vector<u64> cnt(world + 1, 0), send_bytes(world), recv_bytes(world);
  DevBuf<char> send;
  guarded_defer(c, [&] {  // every rank reaches the exchange below, whatever happens locally
    if (!n || failed_locally(c)) { send.alloc(elem_bytes, s); return; }
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
    sorted_digit_counts(w ? kb.p : ka.p, n, 0, (1u << bits) - 1, world + 1, dcnt.p, s);
    c->lc.rb.get(cnt.data(), dcnt.p, (world + 1) * sizeof(u64), s);
    c->lc.rb.sync(s);
    size_t n_send = n - (size_t)cnt[world];
    send.alloc((n_send ? n_send : 1) * elem_bytes, s);
    if (n_send)
      gather_records_kernel<<<div_up(n_send * words, 256), 256, 0, s>>>((const u32*)recs, words, w ? vb.p : va.p,
                                                                       n_send, (u32*)send.p);
    c->lc.n += 3;
    CB_CUDA(cudaGetLastError());
  });
  // one gather: the count matrix (row = source rank), the caller's piggy-backed sums, the failure flags
  if (failed_locally(c)) std::fill(cnt.begin(), cnt.end(), 0);
  std::vector<u64> mine(cnt.begin(), cnt.begin() + world);
  for (int j = 0; j < n_piggy; j++) mine.push_back(piggy[j]);
  const Gathered G = ex_gather_checked(c, mine.data(), world + n_piggy);
  for (int j = 0; j < n_piggy; j++) piggy[j] = G.sum(world + j);
  n_out = 0;
  for (int r = 0; r < world; r++) {
    send_bytes[r] = cnt[r] * elem_bytes;
    recv_bytes[r] = G.at(r, c->cfg.rank) * elem_bytes;
    n_out += (size_t)G.at(r, c->cfg.rank);
  }
  if (big) {  // stores into the peers' mailboxes when the one-sided path is there
    const void* snd[1] = {send.p};
    const size_t el[1] = {elem_bytes};
    void* rcv[1] = {nullptr};
    if (p2p_exchange(c, G, 1, snd, el, rcv)) {
      out.borrow((char*)rcv[0], (n_out ? n_out : 1) * elem_bytes);
      // (the send buffer goes back to the arena here: the barrier above is stream-ordered after the put kernel)
      return;
    }
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
