// ctx.cuh -- the context object behind the C ABI (include/cloudbrush_gpu.h) and the stage
// entry points implemented in the stage_*.cu files.
#pragma once
#include <deque>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "primitives.cuh"

namespace cb {

// device-side counters (one u64 each), downloaded after every stage group
enum DevCounter {
  DC_READS_GOOD = 0,
  DC_READS_GOODBP,
  DC_READS_SHORT,
  DC_READS_SKIPPED,
  DC_LINES_INVALID,
  DC_LEN_MISMATCH,     // good reads whose length != -readlen (unsupported)
  DC_LONG_ID,          // ids longer than 16 bytes (unsupported in this round)
  DC_SURVIVORS,
  DC_MAX_COV,
  DC_TUPLES_INVALID,   // window slots that emit nothing (poly-A interior windows, palindromes)
  DC_POLYA_WEIGHT,     // BuildHighKmerList weight of the all-A / all-T k-mer
  DC_HKMER,
  DC_EDGES_RAW,        // side-buffer records of the join (edges whose mirror may not be proposed)
  DC_EDGES_M,          // mirror records written straight into their adjacency list
  DC_EDGES_A,          // edges of checkpoint A after assembly
  DC_CAND_VERIFIED,    // verified, maximal candidates (|S|)
  DC_UPKMER_LISTS,     // k-mer groups larger than UP_KMER (order dependent in the reference)
  DC_EDGE_REMOVAL,
  DC_TRANS_EDGE,
  DC_CONTAINED_EDGE,
  DC_TR_OVERFLOW,      // kept-list capacity of the literal TR scan exceeded
  DC_SLOW_LISTS,       // (node, direction) lists that took the literal path
  DC_EDGES_C,          // edges after CutChimericLinks + EdgeRemoval
  DC_EDGES_B,          // edges of checkpoint B
  DC_MAX_LIST,         // longest adjacency list
  DC_BUCKET_OVER,      // hash buckets larger than the bucket kernel holds (flat pipeline takes over)
  DC_COUNT
};

// Bit layout of one directed adjacency record in a u64, most significant field first:
//   node | x | (L - ov) | y | nbr        x, y: 0 = 'f', 1 = 'r'
// so that an ascending sort groups by node, then direction x, then overlap DEscending -- the
// order CutChimericLinks / TransitiveReduction scan in (TransitiveReduction.java:175-221).
struct EdgeLayout {
  int nb = 0;  // bits of a node index
  int ob = 0;  // bits of (L - ov)
  int L = 0;
  __host__ __device__ int total_bits() const { return 2 * nb + ob + 2; }
  __host__ __device__ u64 encode(u32 node, u32 x, u32 ov, u32 y, u32 nbr) const {
    return ((u64)node << (nb + ob + 2)) | ((u64)x << (nb + ob + 1)) | ((u64)(L - ov) << (nb + 1)) |
           ((u64)y << nb) | (u64)nbr;
  }
  __host__ __device__ u32 node(u64 k) const { return (u32)(k >> (nb + ob + 2)); }
  __host__ __device__ u32 x(u64 k) const { return (u32)(k >> (nb + ob + 1)) & 1u; }
  __host__ __device__ u32 ov(u64 k) const { return (u32)L - ((u32)(k >> (nb + 1)) & ((1u << ob) - 1)); }
  __host__ __device__ u32 y(u64 k) const { return (u32)(k >> nb) & 1u; }
  __host__ __device__ u32 nbr(u64 k) const { return (u32)(k & ((1ull << nb) - 1)); }
  __host__ __device__ u32 list(u64 k) const { return (u32)(k >> (nb + ob + 1)); }  // node*2 + x
  __host__ __device__ u64 mirror(u64 k) const {  // Node.flip_link: xy -> flip(y) flip(x)
    return encode(nbr(k), y(k) ^ 1u, ov(k), x(k) ^ 1u, node(k));
  }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  bool done = false;
};

// One checkpoint of the graph: adjacency list `lid` = node*2 + x owns the slot range
// keys[adj[lid].x .. adj[lid].x + adj[lid].y); a slot holds a record or ~0 (empty), in no
// particular order (the ranges are handed out per k-mer bucket, not in list order).  All
// checkpoints share cb_ctx::adj and the slot positions, so removing an edge is writing ~0 into its slot.
struct EdgeSet {
  DevBuf<u64> keys;  // [slots], ~0 = empty slot; or, if compact, the n records back to back in no order
  size_t n = 0;      // records
  bool valid = false;
  bool compact = false;
};

}  // namespace cb

struct cb_ctx {
  cb::Arena arena;  // declared first: destroyed after every DevBuf below
  cb_config cfg;
  cudaStream_t stream = nullptr;
  std::string err;
  cb::LaunchCounter lc;
  std::map<std::string, int64_t> counters;
  std::map<std::string, cb::Timer> timers;

  // multi-GPU (cfg.world > 1): transport callbacks and the global numbering
  cb_exchange ex = {};
  bool has_ex = false;
  void* nccl_comm = nullptr;      // native transport (cb_init_nccl); takes precedence over the callbacks
  void* nccl_comm_side = nullptr; // a second communicator (ncclCommSplit) for the side stream: NCCL orders the operations
                                  //   of ONE communicator, so the node-table gather would hold back the main stream's
  cudaStream_t side_stream = nullptr;  // the node-table all-gather runs here, behind keygen + tuple exchange + partition
  cudaStream_t ex_stream = nullptr;    // stream the transport calls are enqueued on (nullptr: the context's stream)
  cudaEvent_t ev_fork = nullptr, ev_seq = nullptr, ev_rest = nullptr;
  bool gather_pending_seq = false, gather_pending_rest = false;
  bool gather_deferred = false;   // the gather has not been enqueued yet (launch_node_gather)
  std::vector<u64> gather_counts; // nodes built by every rank
  cb::DevBuf<u64> own_seq;        // packed reads / coverage of the nodes built from this shard (keygen reads these
  cb::DevBuf<u32> own_cov;        //   while the replicated table is still being gathered)
  cb::DevBuf<u32> own_line_tmp;   // send buffers of the pending gather
  cb::DevBuf<cb::IdKey> own_id_tmp;
  u64* ex_scratch = nullptr;      // small device scratch of the native transport
  size_t ex_scratch_words = 0;
  // One-sided exchange (native transport, every rank on one NVLink domain): each rank owns a mailbox that
  // every other rank maps over CUDA IPC; the big exchanges of the path are stores into the peers' mailboxes.
  bool p2p_tried = false, p2p_on = false;
  char* p2p_box = nullptr;        // this rank's mailbox (cudaMalloc; valid until the next big exchange)
  size_t p2p_cap = 0;
  std::vector<char*> p2p_peer;    // [world] every rank's mailbox as mapped here (own entry: p2p_box)
  int64_t n_host_coll = 0;        // host-synchronous collectives since cb_preprocess began (counter "host_collectives")
  int carried_code = 0;           // first rank-local failure since the last checked gather (defer_error)
  std::string carried_msg;
  int lg_world = 0;               // log2(world)
  int pbit = 0;                   // k-mer groups are owned by rank (key >> pbit) & (world - 1)
  u32 line_base = 0;              // global index of this shard's first SFA line
  std::vector<u64> line_bases;    // [world + 1]
  u32 node_lo = 0, node_hi = 0;   // nodes built from this shard (global node indices)
  std::vector<u64> node_bases;    // [world + 1] first node of every rank's shard
  uint64_t gcount[cb::DC_COUNT] = {0};  // counters reduced over the ranks

  // input text
  // cb_prefetch_sfa_buffer: the NEXT inputs (at most two), copied on copy_stream while this one is processed
  struct Prefetched {
    cb::DevBuf<char> buf;
    size_t n = 0;
    cudaEvent_t done = nullptr;  // recorded on copy_stream behind the copy
  };
  std::deque<Prefetched> next_q;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_main = nullptr;
  // cb_export_*_begin: the records are produced on the context's stream into these buffers and copied to the
  // caller's memory on d2h_stream (device-to-host, next to the prefetch's host-to-device copy)
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ev_d2h = nullptr;
  cb::DevBuf<cb_node> stage_nodes;
  cb::DevBuf<cb_edge> stage_edges;
  bool d2h_pending = false;
  cb::DevBuf<char> text;          // owned copy (cb_load_sfa / cb_load_sfa_buffer / cb_load_sfa_device)
  const char* text_ptr = nullptr; // what the kernels read: text.p or a borrowed device buffer
  size_t text_n = 0;
  bool loaded = false, preprocessed = false, overlapped = false, reduced = false;

  // parse
  u32 n_lines = 0;
  int L = 0, K = 0, NW = 0, W = 0;  // read length, k, words per read, windows per read
  cb::DevBuf<u32> line_start;       // [n_lines + 1]
  cb::DevBuf<uint8_t> status;       // [n_lines] 0 good, 1 invalid, 2 skipped, 3 short
  cb::DevBuf<u64> reads;            // [n_lines][NW]
  cb::DevBuf<cb::IdKey> idkey;      // [n_lines]
  cb::DevBuf<u64> dcount;           // [DC_COUNT]
  uint64_t hcount[cb::DC_COUNT] = {0};

  // nodes (00-preprocess)
  u32 n_nodes = 0;
  cb::DevBuf<u32> node_line;
  cb::DevBuf<u64> node_seq;         // [n_nodes][NW]
  cb::DevBuf<cb::IdKey> node_idkey;
  cb::DevBuf<u32> node_cov;

  // graph
  cb::EdgeLayout lay;
  cb::DevBuf<uint2> adj;            // [2*n_nodes] (first slot, slots) of list node*2 + x
  size_t slots = 0;
  cb::DevBuf<uint4> list_order;     // bucket path: (list, first slot, slots) of the lists that own slots, in slot order
  size_t n_list_order = 0;
  cb::DevBuf<uint8_t> list_flag;    // [2*n_nodes] 1: all overhangs nest (consistent), 0: not, 2: unknown
  cb::DevBuf<u32> ovh;              // [slots][NWo] overhang nbr^y[ov:] of the record in the slot (any checkpoint),
  int NWo = 0;                      //   2-bit packed, left aligned, 16 bases per word; NWo = ceil((L-K)/16)
  // 01-overlap lives in ckA.keys until CutChimericLinks removes records IN PLACE: from then on the
  // array is 01-overlap.chl2 and (slot, record) of everything removed is in cut_log, from which an
  // export of 01-overlap restores a private copy.  01-overlap.re (ckB) is a compact record list.
  cb::EdgeSet ckA, ckB;
  cb::DevBuf<u64> cut_log;          // [2 * n_cut_log]: slot, record
  size_t n_cut_log = 0;
  size_t n_edges_c = 0;             // records of 01-overlap.chl2
};

// kernels are templated on the words per packed read (NW = 2, 4, 6, 8 <=> reads of up to 64 .. 256 bases)
#define CB_DISPATCH_NW(nw, CALL)                                             \
  switch (nw) {                                                              \
    case 2: { constexpr int NW_ = 2; CALL; } break;                          \
    case 4: { constexpr int NW_ = 4; CALL; } break;                          \
    case 6: { constexpr int NW_ = 6; CALL; } break;                          \
    case 8: { constexpr int NW_ = 8; CALL; } break;                          \
    default: throw Error(CB_E_UNSUPPORTED, "read length above 256 bases");   \
  }

namespace cb {

inline bool distributed(const cb_ctx* c) { return c->cfg.world > 1; }
void timer_start(cb_ctx* c, const char* name);
void timer_stop(cb_ctx* c, const char* name);
void download_counters(cb_ctx* c);

// dist.cu -- routing helpers of the multi-GPU path
int nccl_unique_id(void* id128, std::string& err);
void nccl_init(cb_ctx* c, const void* id128);
void nccl_shutdown(cb_ctx* c);
void ex_allreduce(cb_ctx* c, u64* vals, int n, int op);
// Rank-local failures ahead of a collective: every rank calls this at the same point with its own
// code (0 = fine); if any rank failed, EVERY rank throws (a rank that returned early would leave
// its peers blocked inside the next NCCL call).  On a single GPU it simply throws.
void sync_error(cb_ctx* c, int local_code, const std::string& msg);
// runs f(); a cb::Error it throws is turned into a collective failure (sync_error)
template <class F>
inline void guarded(cb_ctx* c, F&& f) {
  int code = 0;
  std::string msg;
  try {
    f();
  } catch (const Error& e) {
    code = e.code;
    msg = e.what();
    if (!distributed(c)) throw;
  }
  sync_error(c, code, msg);
}
// The cheaper form of the same rule: the failure is remembered and travels as one extra word with the NEXT
// checked gather (ex_gather_checked / reduce_counters / route_records), where every rank throws.  Until then
// the failing rank goes on to that gather and must skip the device work that depended on what failed
// (failed_locally).  On a single GPU the failure is thrown where it happens.
void defer_error(cb_ctx* c, int code, const std::string& msg);
inline bool failed_locally(const cb_ctx* c) { return c->carried_code != 0; }
template <class F>
inline void guarded_defer(cb_ctx* c, F&& f) {
  try {
    f();
  } catch (const Error& e) {
    if (!distributed(c)) throw;
    defer_error(c, e.code, e.what());
  }
}
// n values of every rank on every rank, through ONE host-synchronous collective
struct Gathered {
  int world = 1, n = 0;
  std::vector<u64> v;  // [world][n]
  u64 at(int r, int j) const { return v[(size_t)r * n + j]; }
  u64 sum(int j) const { u64 t = 0; for (int r = 0; r < world; r++) t += at(r, j); return t; }
  u64 max(int j) const { u64 t = 0; for (int r = 0; r < world; r++) t = at(r, j) > t ? at(r, j) : t; return t; }
};
// gathers mine[0..n) and, with them, every rank's deferred failure: if any rank carries one, every rank throws here
Gathered ex_gather_checked(cb_ctx* c, const u64* mine, int n);
// in-place max over the ranks of a DEVICE byte array
void ex_allreduce_max_u8(cb_ctx* c, uint8_t* dev, size_t n);
void ex_alltoallv(cb_ctx* c, const void* send, const u64* send_bytes, void* recv, const u64* recv_bytes);
// two all-to-alls with the same element counts (keys and payloads of the k-mer tuples) as ONE transfer group
void ex_alltoallv2(cb_ctx* c, const void* send_a, void* recv_a, size_t elem_a, const void* send_b, void* recv_b,
                   size_t elem_b, const u64* send_cnt, const u64* recv_cnt);
// The big exchanges (k-mer tuples, duplicate-detection records) as ONE kernel of stores into the peers' mailboxes
// over NVLink, when the one-sided path is available: M is the gathered count matrix (M.at(src, dst) elements),
// send[q] holds array q grouped by destination rank.  On success recv[q] points INTO THIS RANK'S MAILBOX (valid
// until the next big exchange) and the stream has passed a barrier after which every peer's stores have landed.
// Returns false when the path is not available (callback transport, ranks on different hosts, no peer access):
// the caller then uses ex_alltoallv.  Collective: every rank takes the same branch.
bool p2p_exchange(cb_ctx* c, const Gathered& M, int narr, const void* const* send, const size_t* elem, void** recv);
void p2p_release(cb_ctx* c);
// Enqueues the node-table gather that cb_preprocess left pending, on the side stream behind whatever the main
// stream holds now.  The overlap stage group calls it right AFTER its tuple exchange: issued earlier (round 2,
// first half) the 1.4 GB gather shared the NVLink with the tuple stores and the exchange could not finish before
// it (a2a_tuples 2.4 ms for a 0.6 ms put kernel at N=4).  No-op when nothing is pending.
void launch_node_gather(cb_ctx* c);
// the main stream waits for the replicated node table (rest: id keys and SFA lines too)
void wait_node_table(cb_ctx* c, bool rest);
void ex_allgatherv(cb_ctx* c, const void* send, u64 send_bytes, void* recv, const u64* recv_bytes);
// counts[d] = keys with digit (key >> shift) & mask == d, for keys sorted by that digit (device arrays)
void sorted_digit_counts(const u64* keys, size_t n, int shift, u32 mask, int nbins, u64* counts, cudaStream_t s);
// routes n records of elem_bytes (multiple of 4) to rank dest[i] (dest[i] == world: dropped);
// out receives the records addressed to this rank, grouped by source rank
// (a checked gather: deferred failures surface here; piggy[0..n_piggy) are summed over the ranks on the way)
// big: one of the big exchanges -- `out` may then be a view of this rank's mailbox (see p2p_exchange)
void route_records(cb_ctx* c, const void* recs, size_t elem_bytes, const u32* dest, size_t n, DevBuf<char>& out,
                   size_t& n_out, u64* piggy = nullptr, int n_piggy = 0, bool big = false);
// all-gathers a per-rank array of `count` elements into `out` (rank order); counts[r] of every rank
void gather_all(cb_ctx* c, const void* local, size_t elem_bytes, size_t count, const std::vector<u64>& counts,
                void* out);
// global counters (sum over ranks; max for the two maxima); identity on a single GPU.  One checked gather, which
// also carries the caller's n_extra values: they come back per rank in columns DC_COUNT + j of the result.
Gathered reduce_counters(cb_ctx* c, const u64* extra = nullptr, int n_extra = 0);

// stage_preprocess.cu
void stage_parse_and_pack(cb_ctx* c);

// stage_overlap.cu
void stage_build_overlap(cb_ctx* c);
// stage_string.cu
void stage_string_graph(cb_ctx* c);
void restore_overlap_copy(cb_ctx* c, DevBuf<u64>& out);

}  // namespace cb
