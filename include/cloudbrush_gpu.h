/* cloudbrush_gpu.h -- C ABI of the B200 overlap-discovery + transitive-reduction path.
 *
 * This is the drop-in boundary for the Brush stages of CloudBrush's hot path
 * (reference: src/Brush/BrushAssembler.java:256-379):
 *
 *   cb_preprocess      replaces  GenNonContainedReads.run   (GenNonContainedReads.java:256-289)
 *                                RedundantRemoval.run       (RedundantRemoval.java:107-139)
 *                                BuildHighKmerList.run      (BuildHighKmerList.java:121-153)
 *   cb_build_overlap   replaces  MatchPrefix.run            (MatchPrefix.java:435-469)
 *                                VerifyOverlap.run          (VerifyOverlap.java:348-382)
 *                                GenReverseEdge.run         (GenReverseEdge.java:248-280)
 *   cb_string_graph    replaces  CutChimericLinks.run x<=2  (CutChimericLinks.java:415-450)
 *                                EdgeRemoval.run x<=3       (EdgeRemoval.java:207-242)
 *                                TransitiveReduction.run    (TransitiveReduction.java:549-583)
 *   cb_get_counter     replaces  RunningJob.getCounters().findCounter("Brush", name)
 *                                                           (BrushAssembler.java:124-127)
 *   cb_write_nodes     replaces  the TextOutputFormat part files of 00-preprocess, 01-overlap,
 *                                01-overlap.chl2, 01-overlap.re in Node.toNodeMsg format
 *                                                           (Node.java:1757-1982)
 *   cb_load_sfa*       replaces  TextInputFormat over the -reads directory (id \t SEQ lines,
 *                                data/preprocessor.pl:27-52)
 *
 * Conventions: plain C types only; every function returns 0 (CB_OK) or a negative CB_E* code
 * and never throws or exits; one context per thread; a context owns all device memory and one
 * CUDA stream; host buffers passed in are borrowed for the duration of the call only.  There is
 * NO CPU fallback: without a CUDA device cb_create fails with CB_E_NODEVICE.
 *
 * The Java side (java/Brush/GpuOverlap.java + java/jni/cloudbrush_jni.c)
 * maps a non-zero return to IOException so BrushAssembler.end(job) behaves as before.
 */
#ifndef CLOUDBRUSH_GPU_H
#define CLOUDBRUSH_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 1

enum {
  CB_OK = 0,
  CB_E_INVALID = -1,      /* bad argument / call order */
  CB_E_NODEVICE = -2,     /* no CUDA device / driver */
  CB_E_CUDA = -3,         /* CUDA runtime error (message in cb_last_error) */
  CB_E_NOMEM = -4,        /* device or host allocation failed */
  CB_E_IO = -5,           /* file could not be read / written */
  CB_E_NOREADS = -6,      /* reads_good == 0  (BrushAssembler.java:273-276 "No good reads") */
  CB_E_UNSUPPORTED = -7,  /* input outside the supported contract */
  CB_E_CAPACITY = -8      /* an internal fixed capacity was exceeded */
};

/* Checkpoints (directories under <asm>.tmp/ in the reference, BrushAssembler.java:75-89). */
enum {
  CB_CK_PREPROCESS = 0, /* 00-preprocess      : nodes, no edges            */
  CB_CK_OVERLAP = 1,    /* 01-overlap         : checkpoint A               */
  CB_CK_CHL2 = 2,       /* 01-overlap.chl2    : after CutChimericLinks     */
  CB_CK_RE = 3          /* 01-overlap.re      : checkpoint B               */
};

/* BrushConfig.java:55-84 -- the options the path reads (defaults in cb_default_config). */
typedef struct cb_config {
  int32_t abi_version; /* CB_ABI_VERSION */
  int32_t k;           /* -k        (BrushConfig.K)        odd, 1 <= k <= 31           */
  int32_t readlen;     /* -readlen  (BrushConfig.READLEN)                              */
  int64_t up_kmer;     /* -kmerup   (UP_KMER  = 2000)                                  */
  int64_t low_kmer;    /* -kmerlow  (LOW_KMER = 1)                                     */
  float majority;      /* -maj      (MAJORITY = 0.6f)                                  */
  float pwm_n;         /* -N        (PWM_N    = 0.1f)                                  */
  int32_t device;      /* CUDA device ordinal                                          */
  int32_t rank;        /* this process' shard index   (0 for a single GPU)             */
  int32_t world;       /* number of shards / GPUs     (1 for a single GPU)             */
  int32_t flags;       /* CB_FLAG_* (0 = defaults)                                     */
  int32_t reserved[4];
} cb_config;

/* cb_config.flags.  CB_FLAG_FLAT_PIPELINE: use the flat shuffle replacement (LSD radix sort of the
 * k-mer tuples + flat join) for every input instead of the hash-bucket path; the library falls back
 * to it on its own when a hash bucket does not fit the bucket kernel.  Results are identical. */
#define CB_FLAG_FLAT_PIPELINE 1

/* One directed adjacency entry: the last `ov` bases of read `node` in orientation x equal the
 * first `ov` bases of read `nbr` in orientation y; type = 0 ff, 1 fr, 2 rf, 3 rr (Node.java:78).
 * `node` / `nbr` are 0-based line indices of the SFA input. */
typedef struct cb_edge {
  uint32_t node;
  uint32_t nbr;
  uint16_t ov;
  uint8_t type;
  uint8_t pad;
} cb_edge;

/* One node of the graph (00-preprocess and later): SFA line, coverage (*v). */
typedef struct cb_node {
  uint32_t line;
  uint32_t cov;
} cb_node;

typedef struct cb_ctx cb_ctx;

void cb_default_config(cb_config* cfg);
int cb_create(const cb_config* cfg, cb_ctx** out);
void cb_destroy(cb_ctx* ctx);
const char* cb_last_error(const cb_ctx* ctx); /* library-owned, valid until the next call */
const char* cb_version(void);

/* Input: SFA text ("id\tSEQ[\tQ]\n" lines).  cb_load_sfa reads a file or every regular file of
 * a directory (sorted by name); cb_load_sfa_buffer copies a caller-owned host buffer to the
 * device; cb_load_sfa_device adopts nothing and copies from a caller-owned DEVICE buffer. */
int cb_load_sfa(cb_ctx* ctx, const char* path_or_dir);
int cb_load_sfa_buffer(cb_ctx* ctx, const char* buf, size_t n);
int cb_load_sfa_device(cb_ctx* ctx, const void* dev_buf, size_t n);
/* Zero-copy variant: the context reads the caller's DEVICE buffer in place; it must stay valid
 * and unchanged until the next cb_load_* / cb_borrow_* / cb_destroy.  The buffer must be 16-byte
 * aligned (CB_E_INVALID otherwise: the parser reads aligned 16-byte vectors).  For both *_device
 * calls the work that PRODUCES the buffer must be complete (synchronise the producing stream)
 * before the call; cb_load_sfa_device has finished its copy when it returns. */
int cb_borrow_sfa_device(cb_ctx* ctx, const void* dev_buf, size_t n);

/* Pipelined input for callers that process one batch after another: starts the host-to-device copy of
 * a LATER input on a copy stream of the context's own and returns at once; the copy runs while the stage
 * groups of the current input execute.  Up to two prefetched inputs wait in order (so that the copies run
 * back to back, also during cb_preprocess; a third is CB_E_INVALID); the oldest becomes the current input
 * at the next cb_preprocess (which waits for its copy).  `buf` must be page-locked host memory and must
 * stay valid and unchanged until that cb_preprocess has returned.  Results of the current input stay
 * available until then.  A blocking cb_load_* / cb_borrow_* discards what was prefetched. */
int cb_prefetch_sfa_buffer(cb_ctx* ctx, const char* buf, size_t n);

/* The three stage groups, run to completion on the calling thread (see top of file). */
int cb_preprocess(cb_ctx* ctx);
int cb_build_overlap(cb_ctx* ctx);
int cb_string_graph(cb_ctx* ctx);

/* Hadoop counters of group "Brush" with the reference's names: reads_good, reads_goodbp,
 * reads_short, reads_skipped, input_lines_invalid, contained, nodecount, redundant, nodes,
 * hkmer, edge_removal, edge_removal_1, edge_removal_2, trans_edge, contained_edge; plus the
 * library's own: tuples, candidates_verified, edges_overlap, edges_chl2, edges_re,
 * order_dependent, upkmer_truncated_lists.  Unknown name -> CB_E_INVALID. */
int cb_get_counter(cb_ctx* ctx, const char* name, int64_t* value);

/* Size query then fill (out == NULL or cap too small -> *n is set, nothing is written).
 * Edges come sorted by (node, type, nbr, ov); nodes by line. */
int cb_export_edges(cb_ctx* ctx, int checkpoint, cb_edge* out, size_t cap, size_t* n);
int cb_export_nodes(cb_ctx* ctx, cb_node* out, size_t cap, size_t* n);

/* The same exports for a caller that processes batch after batch (the device-to-host counterpart of
 * cb_prefetch_sfa_buffer): the records are produced on the context's stream and copied into `out` on a
 * separate copy stream while the context goes on with its next stage group; *n is set at once, `out`
 * (pinned host memory, or the copy is not asynchronous) holds the records after cb_export_wait.  One
 * transfer of each kind may be in flight; a second cb_export_nodes_begin / cb_export_edges_begin first
 * waits for the transfers before it.  Replaces the same reference steps as the blocking forms
 * (the part files of TextOutputFormat, BrushAssembler.java:75-89). */
int cb_export_nodes_begin(cb_ctx* ctx, cb_node* out, size_t cap, size_t* n);
int cb_export_edges_begin(cb_ctx* ctx, int checkpoint, cb_edge* out, size_t cap, size_t* n);
int cb_export_wait(cb_ctx* ctx);

/* Writes <out_dir>/part-000<rank> in Node.toNodeMsg format (out_dir is emptied first, like
 * FileSystem.delete(outputPath, true) in every stage's run()): part-00000 on a single GPU.  On a sharded
 * context the call is COLLECTIVE (every rank calls it for the same checkpoint and directory, which all
 * ranks must be able to reach): each rank writes the nodes built from its shard of the SFA lines, with ALL
 * their edges (records held by other ranks are routed to it), so the directory is a multi-part Hadoop
 * output that the unchanged compressChains reads (BrushAssembler.java:75-89, :381). */
int cb_write_nodes(cb_ctx* ctx, int checkpoint, const char* out_dir);

/* The nodes as FASTA, the format of Graph2Fasta.java:49-75 (">id\tlen=<L>\tcov=<cov>.0", sequence in lines
 * of 60): <out_dir>/part-000<rank>, emptied first; collective on a sharded context like cb_write_nodes. */
int cb_write_fasta(cb_ctx* ctx, const char* out_dir);

/* Device time (ms, CUDA events) of the last run of a stage group: "preprocess", "overlap",
 * "string", or one of the per-kernel-group names listed in DESIGN.md ("sort_tuples", ...). */
int cb_get_time_ms(cb_ctx* ctx, const char* what, double* ms);

/* Number of kernel launches issued by this context since creation. */
int cb_get_launch_count(cb_ctx* ctx, int64_t* n);

/* Per-kernel device timing.  cb_set_profile(ctx, 1) clears the statistics and makes every
 * instrumented kernel launch be bracketed by CUDA events on the context's stream;
 * cb_get_kernel_stat returns the summed duration and the launch count of one kernel
 * ("rs_scatter:<sort>" and "rs_upsweep:<sort>" with <sort> in tuples, dedupe, edges, export; "parse_lines",
 * "dedupe_runs", "keygen", "join_verify", "cut", "tr"). */
int cb_set_profile(cb_ctx* ctx, int on);
int cb_get_kernel_stat(cb_ctx* ctx, const char* name, double* total_ms, int64_t* launches);

/* Test hook (kernel-level parity of the shuffle replacement, tests/test_gpu_sort.py): sorts the
 * caller's HOST arrays in place on the device with the library's LSD radix sort -- stable, on key
 * bits [begin_bit, end_bit) -- and copies them back.  vals may be NULL (keys only).  n < 2^32. */
int cb_debug_sort(cb_ctx* ctx, uint64_t* keys, uint32_t* vals, size_t n, int begin_bit, int end_bit);

/* Test hooks of the hash partition that replaces the shuffle on the default path (tests/test_gpu_sort.py,
 * tests/test_capi_host.py).  cb_debug_bucket_plan: the two digit widths, the bucket count and whether two
 * passes suffice for n_tuples (host only, no device needed).  cb_debug_partition: partitions the caller's HOST
 * tuples (key = k-mer in the low kmer_bits bits, all-ones keys are dropped) in place on the device:
 * keys/vals come back bucketed (n_kept of them), bstart[n_buckets + 1] holds the bucket offsets; with
 * bstart == NULL or too small only *n_buckets is set. */
int cb_debug_bucket_plan(size_t n_tuples, int32_t* b1, int32_t* b2, uint32_t* n_buckets, int32_t* fits);
int cb_debug_partition(cb_ctx* ctx, uint64_t* keys, uint32_t* vals, size_t n, int kmer_bits, uint32_t* bstart,
                       size_t bstart_cap, uint32_t* n_buckets, size_t* n_kept);

/* ---- multi-GPU: one process (and one context) per GPU -------------------------------------
 * The MapReduce shuffle of the reference (HashPartitioner on the Text key inside every
 * JobClient.runJob, e.g. MatchPrefix.java:468) becomes an all-to-all of k-mer tuples between the
 * contexts: context `rank` of `world` (cb_config) loads ITS shard of the SFA lines (shards are
 * consecutive line ranges in rank order) and owns the k-mer groups whose key hashes to it.  The
 * library does the partitioning on the device and calls back into the host program for the
 * actual transfers, so the transport (torch.distributed over NCCL in cloudbrush_b200/dist.py, or
 * NCCL/MPI directly from a Java launcher) stays outside the ABI.  world must be a power of two.
 *
 * Every callback returns 0 on success.  Device buffers are complete when a callback is entered
 * (the library synchronises its stream first) and the callback must have finished the transfer
 * when it returns.  Counts are in BYTES. */
typedef struct cb_exchange {
  void* user;
  /* recv[i] = what rank i sent to this rank; arrays of world * n_per_rank values (host) */
  int (*exchange_counts)(void* user, const uint64_t* send, uint64_t* recv, int32_t n_per_rank);
  /* send_dev holds world consecutive blocks of send_bytes[i]; recv_dev receives world blocks */
  int (*alltoallv)(void* user, const void* send_dev, const uint64_t* send_bytes, void* recv_dev,
                   const uint64_t* recv_bytes);
  /* every rank contributes send_bytes; recv_dev receives the world blocks in rank order */
  int (*allgatherv)(void* user, const void* send_dev, uint64_t send_bytes, void* recv_dev,
                    const uint64_t* recv_bytes);
  /* in-place all-reduce of n host values; op 0 = sum, 1 = max */
  int (*allreduce_u64)(void* user, uint64_t* vals, int32_t n, int32_t op);
  /* in-place element-wise MAX over the ranks of a DEVICE array of n bytes */
  int (*allreduce_max_u8_dev)(void* user, void* dev, uint64_t n);
} cb_exchange;

/* Must be called before cb_preprocess when cfg.world > 1 (or cb_init_nccl, below).  On a sharded
 * context counters are global sums, cb_export_nodes returns the nodes built from this rank's shard
 * of the SFA lines (in rank order their union is the node table sorted by line) and
 * cb_export_edges the adjacency lists this rank owns (their union over the ranks is the checkpoint).
 * cb_preprocess, cb_build_overlap, cb_string_graph, cb_write_nodes and cb_write_fasta are COLLECTIVE on a
 * sharded context: every rank calls them, in the same order.  The exports (blocking and *_begin) are local:
 * a rank may call them on its own. */
int cb_set_exchange(cb_ctx* ctx, const cb_exchange* ex);

/* Native transport: NCCL over NVLink / NVSwitch, driven by the library on its own stream (no host
 * round trip per transfer; replaces the cb_exchange callbacks, which stay available for hosts
 * that bring their own transport).  Rank 0 obtains the 128-byte id with cb_nccl_get_unique_id and
 * the host program hands it to every rank by any channel (torch.distributed broadcast in
 * cloudbrush_b200/api.py; a socket or MPI_Bcast from a Java launcher); then every rank calls
 * cb_init_nccl before cb_preprocess.  libnccl.so.2 is located at run time (dlopen; the
 * environment variable CB_NCCL_LIB overrides the name); CB_E_UNSUPPORTED if it cannot be loaded. */
#define CB_NCCL_ID_BYTES 128
int cb_nccl_get_unique_id(void* id128);
const char* cb_nccl_last_error(void); /* why the last cb_nccl_get_unique_id of this thread failed (library-owned) */
int cb_init_nccl(cb_ctx* ctx, const void* id128);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDBRUSH_GPU_H */
