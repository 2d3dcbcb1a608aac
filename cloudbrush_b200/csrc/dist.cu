// dist.cu -- multi-GPU plumbing: device-side partitioning of records by destination rank and the
// calls into the host program's transport (cb_exchange, include/cloudbrush_gpu.h).
//
// This replaces Hadoop's HashPartitioner + shuffle (every JobClient.runJob of the path, e.g.
// MatchPrefix.java:468) with: partition on the device -> exchange byte counts -> one all-to-all.
#include "ctx.cuh"

namespace cb {

static void ex_fail(const char* what, int rc) {
  throw Error(CB_E_CUDA, std::string("exchange callback ") + what + " failed with " + std::to_string(rc));
}

static void need_ex(cb_ctx* c) {
  if (!c->has_ex) throw Error(CB_E_INVALID, "cfg.world > 1 but no cb_exchange was set (cb_set_exchange)");
}

void ex_allreduce(cb_ctx* c, u64* vals, int n, int op) {
  if (!distributed(c)) return;
  need_ex(c);
  int rc = c->ex.allreduce_u64(c->ex.user, (uint64_t*)vals, n, op);
  if (rc) ex_fail("allreduce_u64", rc);
}

void ex_allreduce_max_u8(cb_ctx* c, uint8_t* dev, size_t n) {
  if (!distributed(c)) return;
  need_ex(c);
  if (!c->ex.allreduce_max_u8_dev) throw Error(CB_E_INVALID, "cb_exchange.allreduce_max_u8_dev is not set");
  CB_CUDA(cudaStreamSynchronize(c->stream));
  int rc = c->ex.allreduce_max_u8_dev(c->ex.user, dev, (uint64_t)n);
  if (rc) ex_fail("allreduce_max_u8_dev", rc);
}

void ex_counts(cb_ctx* c, const u64* send, u64* recv, int n_per_rank) {
  need_ex(c);
  int rc = c->ex.exchange_counts(c->ex.user, (const uint64_t*)send, (uint64_t*)recv, n_per_rank);
  if (rc) ex_fail("exchange_counts", rc);
}

void ex_alltoallv(cb_ctx* c, const void* send, const u64* send_bytes, void* recv, const u64* recv_bytes) {
  need_ex(c);
  CB_CUDA(cudaStreamSynchronize(c->stream));
  int rc = c->ex.alltoallv(c->ex.user, send, (const uint64_t*)send_bytes, recv, (const uint64_t*)recv_bytes);
  if (rc) ex_fail("alltoallv", rc);
}

void ex_allgatherv(cb_ctx* c, const void* send, u64 send_bytes, void* recv, const u64* recv_bytes) {
  need_ex(c);
  CB_CUDA(cudaStreamSynchronize(c->stream));
  int rc = c->ex.allgatherv(c->ex.user, send, send_bytes, recv, (const uint64_t*)recv_bytes);
  if (rc) ex_fail("allgatherv", rc);
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
  if (n) {
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
  } else {
    send.alloc(elem_bytes, s);
  }
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
