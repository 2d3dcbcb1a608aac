/*
 * GpuOverlap.java -- JNI shim that lets the unmodified CloudBrush driver run the
 * overlap-discovery + transitive-reduction path on a B200 through libcloudbrush_gpu.so.
 *
 * It mirrors the three driver methods it replaces (BrushAssembler.java:256-379) and returns an
 * org.apache.hadoop.mapred.RunningJob for each, so BrushAssembler.end(job) / counter(job, tag)
 * (BrushAssembler.java:104-127) keep working.  NOT compiled in this repository's image (no JDK,
 * no Hadoop jars): see INTEGRATION.md for the build line and the three-line driver change.
 */
package Brush;

import java.io.IOException;

import org.apache.hadoop.mapred.Counters;
import org.apache.hadoop.mapred.JobID;
import org.apache.hadoop.mapred.RunningJob;
import org.apache.hadoop.mapred.TaskAttemptID;
import org.apache.hadoop.mapred.TaskCompletionEvent;

public class GpuOverlap implements AutoCloseable {
    static { System.loadLibrary("cloudbrush_jni"); }

    // checkpoints of include/cloudbrush_gpu.h
    public static final int CK_PREPROCESS = 0, CK_OVERLAP = 1, CK_CHL2 = 2, CK_RE = 3;

    private long ctx;  // cb_ctx*

    // ---- native side: java/jni/cloudbrush_jni.c, one call per C-ABI entry point ----
    private static native long nCreate(int k, int readlen, long upKmer, long lowKmer, float majority, float pwmN,
                                       int device, int rank, int world);
    private static native void nDestroy(long ctx);
    private static native String nLastError(long ctx);
    private static native int nLoadSfa(long ctx, String pathOrDir);
    private static native int nPreprocess(long ctx);
    private static native int nBuildOverlap(long ctx);
    private static native int nStringGraph(long ctx);
    private static native int nWriteNodes(long ctx, int checkpoint, String outDir);
    private static native long nCounter(long ctx, String name);      // Long.MIN_VALUE if unknown
    private static native double nTimeMs(long ctx, String what);
    private static native byte[] nNcclUniqueId();                      // multi-GPU launchers: rank 0
    private static native int nInitNccl(long ctx, byte[] id128);      // every rank, before preprocess
    private static native String nNcclLastError();                    // why nNcclUniqueId returned null

    public GpuOverlap() throws IOException {
        this(0, 0, 1);
    }

    /** One JVM per GPU: shard `rank` of `world` on CUDA device `device` (the launcher then hands rank 0's
     *  ncclUniqueId() to every rank and each calls initNccl before preprocess; cb_write_nodes is collective
     *  and leaves one part-000<rank> per rank in every output directory). */
    public GpuOverlap(int device, int rank, int world) throws IOException {
        // the options the path reads, straight from BrushConfig (BrushConfig.java:55-84)
        ctx = nCreate((int) BrushConfig.K, (int) BrushConfig.READLEN, BrushConfig.UP_KMER, BrushConfig.LOW_KMER,
                      BrushConfig.MAJORITY, BrushConfig.PWM_N, device, rank, world);
        if (ctx == 0) throw new IOException("cb_create failed: no CUDA device (there is no CPU fallback)");
    }

    public static byte[] ncclUniqueId() throws IOException {
        byte[] id = nNcclUniqueId();
        if (id == null) throw new IOException("NCCL is not available: " + nNcclLastError());
        return id;
    }

    public void initNccl(byte[] id128) throws IOException {
        check(nInitNccl(ctx, id128));
    }

    private void check(int rc) throws IOException {
        // every non-zero return becomes the IOException JobClient.runJob would have thrown
        if (rc != 0) throw new IOException("cloudbrush_gpu error " + rc + ": " + nLastError(ctx));
    }

    /** BrushAssembler.preprocess (BrushAssembler.java:256-309): jobs GenNonContainedReads,
     *  RedundantRemoval, BuildHighKmerList.  Materialises basePath+preprocess and the (empty) stopword dir. */
    public RunningJob preprocess(String inputPath, String basePath, String preprocess, String hkmerlist)
            throws IOException {
        check(nLoadSfa(ctx, inputPath));
        check(nPreprocess(ctx));                      // CB_E_NOREADS == IOException("No good reads")
        check(nWriteNodes(ctx, CK_PREPROCESS, basePath + preprocess));
        return job("preprocess", "preprocess", "nodecount", "reads_good", "reads_goodbp", "reads_short",
                   "reads_skipped", "contained", "redundant", "nodes", "hkmer");
    }

    /** BrushAssembler.buildOverlap (:313-333): MatchPrefix, VerifyOverlap, GenReverseEdge -> 01-overlap. */
    public RunningJob buildOverlap(String basePath, String overlap) throws IOException {
        check(nBuildOverlap(ctx));
        check(nWriteNodes(ctx, CK_OVERLAP, basePath + overlap));
        return job("buildOverlap", "overlap", "hkmer");
    }

    /** Head of BrushAssembler.buildStringGraph (:337-379): CutChimericLinks/EdgeRemoval x<=2,
     *  TransitiveReduction, EdgeRemoval -> 01-overlap.chl2 and 01-overlap.re. */
    public RunningJob buildStringGraph(String basePath, String overlap) throws IOException {
        check(nStringGraph(ctx));
        check(nWriteNodes(ctx, CK_CHL2, basePath + overlap + ".chl2"));
        check(nWriteNodes(ctx, CK_RE, basePath + overlap + ".re"));
        return job("buildString", "string", "edge_removal", "trans_edge", "contained_edge");
    }

    @Override public void close() { if (ctx != 0) { nDestroy(ctx); ctx = 0; } }

    // ---- a finished RunningJob carrying the "Brush" counters ----
    private RunningJob job(final String name, String timer, String... counterNames) {
        final Counters counters = new Counters();
        for (String c : counterNames) {
            long v = nCounter(ctx, c);
            if (v != Long.MIN_VALUE) counters.findCounter("Brush", c).increment(v);
        }
        final String id = "job_gpu_" + name;
        return new RunningJob() {
            public JobID getID() { return JobID.forName("job_gpu_0001"); }
            public String getJobID() { return id; }
            public String getJobName() { return name; }
            public String getJobFile() { return ""; }
            public String getTrackingURL() { return ""; }
            public float mapProgress() { return 1f; }
            public float reduceProgress() { return 1f; }
            public float cleanupProgress() { return 1f; }
            public float setupProgress() { return 1f; }
            public boolean isComplete() { return true; }
            public boolean isSuccessful() { return true; }   // failures surfaced as IOException above
            public void waitForCompletion() { }
            public int getJobState() { return 2; /* JobStatus.SUCCEEDED */ }
            public void killJob() { }
            public void setJobPriority(String priority) { }
            public TaskCompletionEvent[] getTaskCompletionEvents(int startFrom) { return new TaskCompletionEvent[0]; }
            public void killTask(TaskAttemptID taskId, boolean shouldFail) { }
            public void killTask(String taskId, boolean shouldFail) { }
            public Counters getCounters() { return counters; }
        };
    }
}
