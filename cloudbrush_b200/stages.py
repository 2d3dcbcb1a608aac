"""Host-side mirror of the reference's stage interface for the hot path.

The reference's driver (BrushAssembler.java:256-379) calls one `stage.run(in, out)` per MapReduce
job, then `end(job)` and `counter(job, tag)`.  The GPU path fuses the twelve jobs of the path into
three stage groups with the driver's own names and directory layout:

    GpuOverlap.preprocess(reads, base, "00-preprocess", "stopwords")     BrushAssembler.java:256-309
    GpuOverlap.buildOverlap(reads, base, "00-preprocess", "01-overlap", "stopwords")     :313-333
    GpuOverlap.buildStringGraph(reads, base, "01-overlap", "02-string")  :337-379 (up to 01-overlap.re)

Each returns a RunningJob look-alike (getJobID / isSuccessful / getCounters().findCounter("Brush",
name).getValue()) and materialises the same directories of node-text part files, so `-start` /
`-stop` and the unchanged downstream stages (compressChains ...) keep working.  The Java shim in
java/Brush/GpuOverlap.java has the same three methods over JNI.
"""
import itertools
import os
import shutil

from . import _lib as L
from .api import BrushConfig, BrushGpu, BrushGpuError

_job_ids = itertools.count(1)


class _Counter:
    def __init__(self, v):
        self._v = v

    def getValue(self):
        return self._v


class _Counters:
    def __init__(self, values):
        self._values = dict(values)

    def findCounter(self, group, name):
        if group != "Brush":
            return _Counter(0)
        return _Counter(self._values.get(name, 0))  # Hadoop creates missing counters at 0


class RunningJob:
    """What BrushAssembler.end()/counter() need from org.apache.hadoop.mapred.RunningJob."""

    def __init__(self, name, counters, seconds, ok=True):
        self._id = "job_gpu_%04d" % next(_job_ids)
        self.name = name
        self._counters = _Counters(counters)
        self.seconds = seconds
        self._ok = ok

    def getJobID(self):
        return self._id

    def isSuccessful(self):
        return self._ok

    def getCounters(self):
        return self._counters


PRE_COUNTERS = ["nodecount", "reads_good", "reads_goodbp", "reads_short", "reads_skipped", "contained",
                "redundant", "nodes", "input_lines_invalid"]
OVL_COUNTERS = ["hkmer", "tuples", "candidates_verified", "edges_overlap", "upkmer_truncated_lists",
                "order_dependent"]
STR_COUNTERS = ["edge_removal", "edge_removal_1", "edge_removal_2", "trans_edge", "contained_edge", "edges_chl2",
                "edges_re"]


class GpuOverlap:
    """Drop-in for the bodies of BrushAssembler.preprocess / buildOverlap / buildStringGraph."""

    def __init__(self, cfg: BrushConfig, write_text=True):
        self.cfg = cfg
        self.gpu = BrushGpu(cfg)
        self.write_text = write_text

    def close(self):
        self.gpu.close()

    def _job(self, name, names, what):
        vals = {n: self.gpu.counter(n) for n in names}
        return RunningJob(name, vals, self.gpu.time_ms(what) / 1e3)

    def preprocess(self, inputPath, basePath, preprocess="00-preprocess", hkmerlist="stopwords"):
        self.gpu.load_sfa(inputPath)
        self.gpu.preprocess()   # raises BrushGpuError(CB_E_NOREADS) == IOException("No good reads")
        if self.write_text:
            self.gpu.write_nodes(L.CK_PREPROCESS, os.path.join(basePath, preprocess))
            sw = os.path.join(basePath, hkmerlist)  # BuildHighKmerList output: empty on this path
            shutil.rmtree(sw, ignore_errors=True)
            os.makedirs(sw)
            open(os.path.join(sw, "part-00000"), "w").close()
        return self._job("preprocess", PRE_COUNTERS, "preprocess")

    def buildOverlap(self, inputPath, basePath, preprocess="00-preprocess", overlap="01-overlap",
                     hkmerlist="stopwords"):
        self.gpu.build_overlap()
        if self.write_text:
            self.gpu.write_nodes(L.CK_OVERLAP, os.path.join(basePath, overlap))
        return self._job("buildOverlap", OVL_COUNTERS, "overlap")

    def buildStringGraph(self, inputPath, basePath, overlap="01-overlap", Sgraph="02-string"):
        self.gpu.string_graph()
        if self.write_text:
            self.gpu.write_nodes(L.CK_CHL2, os.path.join(basePath, overlap + ".chl2"))
            self.gpu.write_nodes(L.CK_RE, os.path.join(basePath, overlap + ".re"))
        return self._job("buildString", STR_COUNTERS, "string")


__all__ = ["GpuOverlap", "RunningJob", "BrushGpuError"]
