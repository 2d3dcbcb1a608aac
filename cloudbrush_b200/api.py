"""BrushGpu: one context of the C-ABI library (one GPU, one thread)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L

EDGE_TYPES = ("ff", "fr", "rf", "rr")  # Node.java:78

EDGE_DTYPE = np.dtype([("node", "<u4"), ("nbr", "<u4"), ("ov", "<u2"), ("type", "u1"), ("pad", "u1")])
NODE_DTYPE = np.dtype([("line", "<u4"), ("cov", "<u4")])


class BrushGpuError(IOError):
    """Raised for every non-zero return of the C ABI (the Java shim raises IOException)."""

    def __init__(self, code, msg):
        super().__init__(f"{L.ERRORS.get(code, code)}: {msg}")
        self.code = code


@dataclass
class BrushConfig:
    """The BrushConfig options the path reads (BrushConfig.java:55-84, :207-406)."""
    k: int = 21            # -k
    readlen: int = 36      # -readlen
    up_kmer: int = 2000    # -kmerup
    low_kmer: int = 1      # -kmerlow
    majority: float = 0.6  # -maj
    pwm_n: float = 0.1     # -N
    device: int = 0
    rank: int = 0
    world: int = 1
    flat_pipeline: bool = False  # CB_FLAG_FLAT_PIPELINE: LSD sort + flat join instead of the hash-bucket path


class BrushGpu:
    def __init__(self, cfg: BrushConfig):
        self._L = L.load()
        c = L.cb_config()
        self._L.cb_default_config(C.byref(c))
        c.k, c.readlen = cfg.k, cfg.readlen
        c.up_kmer, c.low_kmer = cfg.up_kmer, cfg.low_kmer
        c.majority, c.pwm_n = cfg.majority, cfg.pwm_n
        c.device, c.rank, c.world = cfg.device, cfg.rank, cfg.world
        c.flags = L.FLAG_FLAT_PIPELINE if cfg.flat_pipeline else 0
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = self._L.cb_create(C.byref(c), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise BrushGpuError(rc, "cb_create failed (a CUDA device is required; there is no CPU fallback)")

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.cb_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BrushGpuError(rc, self._L.cb_last_error(self._h).decode(errors="replace"))

    # -- input ----------------------------------------------------------------------------
    def load_sfa(self, path):
        self._check(self._L.cb_load_sfa(self._h, str(path).encode()))

    def load_sfa_buffer(self, buf):
        """buf: bytes / bytearray / numpy uint8 array / pinned torch uint8 CPU tensor."""
        if hasattr(buf, "data_ptr"):
            ptr, n = buf.data_ptr(), buf.numel()
        elif isinstance(buf, np.ndarray):
            ptr, n = buf.ctypes.data, buf.nbytes
        else:
            b = bytes(buf)
            self._keep = b
            ptr, n = C.cast(C.c_char_p(b), C.c_void_p).value, len(b)
        self._check(self._L.cb_load_sfa_buffer(self._h, ptr, n))

    def prefetch_sfa_buffer(self, buf):
        """Starts the host-to-device copy of the NEXT input (pinned torch uint8 CPU tensor or numpy array over
        page-locked memory) and returns at once; it becomes the current input at the next preprocess()."""
        if hasattr(buf, "data_ptr"):
            ptr, n = buf.data_ptr(), buf.numel()
        else:
            ptr, n = buf.ctypes.data, buf.nbytes
        self._prefetched = buf  # keep the host buffer alive until it is consumed
        self._check(self._L.cb_prefetch_sfa_buffer(self._h, ptr, n))

    def load_sfa_device(self, dev_ptr, n):
        self._check(self._L.cb_load_sfa_device(self._h, dev_ptr, n))

    def borrow_sfa_device(self, dev_ptr, n):
        """Zero-copy: the context reads the caller's device buffer in place."""
        self._check(self._L.cb_borrow_sfa_device(self._h, dev_ptr, n))

    def set_exchange(self, exchange):
        """exchange: cloudbrush_b200.dist.TorchExchange (required when cfg.world > 1)."""
        self._exchange = exchange  # keeps the ctypes callbacks alive
        st = exchange.as_struct()
        self._check(self._L.cb_set_exchange(self._h, C.cast(C.byref(st), C.c_void_p)))

    def init_nccl(self, group=None):
        """Native transport (cfg.world > 1): the library drives NCCL itself on its own stream.

        torch.distributed (any backend) is only used to hand rank 0's NCCL unique id to the ranks."""
        import torch
        import torch.distributed as dist
        buf = (C.c_uint8 * 128)()
        if dist.get_rank(group) == 0:
            rc = self._L.cb_nccl_get_unique_id(buf)
            if rc != 0:
                raise BrushGpuError(rc, "cb_nccl_get_unique_id failed: " + self._L.cb_nccl_last_error().decode(errors="replace"))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(self.cfg.device)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        self._check(self._L.cb_init_nccl(self._h, ident))

    def set_profile(self, on=True):
        self._check(self._L.cb_set_profile(self._h, 1 if on else 0))

    def kernel_stat(self, name):
        """(total ms, launches) of one instrumented kernel since set_profile(True); None if never launched."""
        ms, n = C.c_double(), C.c_int64()
        rc = self._L.cb_get_kernel_stat(self._h, name.encode(), C.byref(ms), C.byref(n))
        return (ms.value, n.value) if rc == 0 else None

    # -- the three stage groups (BrushAssembler.java:256-379) --------------------------------
    def preprocess(self):
        self._check(self._L.cb_preprocess(self._h))

    def build_overlap(self):
        self._check(self._L.cb_build_overlap(self._h))

    def string_graph(self):
        self._check(self._L.cb_string_graph(self._h))

    def run(self):
        self.preprocess()
        self.build_overlap()
        self.string_graph()

    # -- results --------------------------------------------------------------------------
    def counter(self, name):
        v = C.c_int64()
        self._check(self._L.cb_get_counter(self._h, name.encode(), C.byref(v)))
        return v.value

    def edges(self, checkpoint, out=None):
        """structured array (node, nbr, ov, type) sorted by (node, type, nbr, ov).

        `out`: optional preallocated EDGE_DTYPE array (e.g. a view of pinned host memory, which makes
        the device-to-host copy run at PCIe speed); a slice of it is returned."""
        n = C.c_size_t()
        self._check(self._L.cb_export_edges(self._h, checkpoint, None, 0, C.byref(n)))
        if out is None or len(out) < n.value:
            out = np.empty(n.value, EDGE_DTYPE)
        if n.value:
            self._check(self._L.cb_export_edges(self._h, checkpoint, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def edge_rows(self, checkpoint):
        """int64[n,4] rows (node line, type, nbr line, ov) -- the oracle's canonical form."""
        e = self.edges(checkpoint)
        return np.stack([e["node"].astype(np.int64), e["type"].astype(np.int64), e["nbr"].astype(np.int64),
                         e["ov"].astype(np.int64)], axis=1) if len(e) else np.zeros((0, 4), np.int64)

    def num_edges(self, checkpoint):
        n = C.c_size_t()
        self._check(self._L.cb_export_edges(self._h, checkpoint, None, 0, C.byref(n)))
        return n.value

    def nodes(self, out=None):
        n = C.c_size_t()
        self._check(self._L.cb_export_nodes(self._h, None, 0, C.byref(n)))
        if out is None or len(out) < n.value:
            out = np.empty(n.value, NODE_DTYPE)
        if n.value:
            self._check(self._L.cb_export_nodes(self._h, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value]

    def edges_begin(self, checkpoint, out):
        """Asynchronous edges(): `out` (EDGE_DTYPE view of PINNED host memory, large enough) holds the records
        after export_wait(); the slice that will be valid is returned at once."""
        n = C.c_size_t()
        self._check(self._L.cb_export_edges_begin(self._h, checkpoint, out.ctypes.data, len(out), C.byref(n)))
        if len(out) < n.value:
            raise BrushGpuError(-1, f"edges_begin: buffer of {len(out)} records for {n.value}")
        return out[:n.value]

    def nodes_begin(self, out):
        """Asynchronous nodes() (see edges_begin)."""
        n = C.c_size_t()
        self._check(self._L.cb_export_nodes_begin(self._h, out.ctypes.data, len(out), C.byref(n)))
        if len(out) < n.value:
            raise BrushGpuError(-1, f"nodes_begin: buffer of {len(out)} records for {n.value}")
        return out[:n.value]

    def export_wait(self):
        self._check(self._L.cb_export_wait(self._h))

    def write_nodes(self, checkpoint, out_dir):
        self._check(self._L.cb_write_nodes(self._h, checkpoint, str(out_dir).encode()))

    def debug_sort(self, keys, vals=None, begin_bit=0, end_bit=64):
        """Test hook: the library's stable LSD radix sort on key bits [begin_bit, end_bit), in place on numpy
        arrays (keys uint64, vals uint32 or None)."""
        assert keys.dtype == np.uint64 and keys.flags.c_contiguous
        assert vals is None or (vals.dtype == np.uint32 and vals.flags.c_contiguous and len(vals) == len(keys))
        self._check(self._L.cb_debug_sort(self._h, keys.ctypes.data, vals.ctypes.data if vals is not None else None,
                                          len(keys), begin_bit, end_bit))

    def debug_partition(self, keys, vals, kmer_bits):
        """Test hook: the hash partition of the default path on numpy arrays (keys uint64, vals uint32), in place.
        Returns (keys[:kept], vals[:kept], bstart) with bstart[b] .. bstart[b+1] the tuples of bucket b."""
        assert keys.dtype == np.uint64 and vals.dtype == np.uint32 and len(keys) == len(vals)
        nb, kept = C.c_uint32(), C.c_size_t()
        self._check(self._L.cb_debug_partition(self._h, keys.ctypes.data, vals.ctypes.data, len(keys), kmer_bits, None, 0,
                                               C.byref(nb), C.byref(kept)))
        bstart = np.zeros(nb.value + 1, np.uint32)
        self._check(self._L.cb_debug_partition(self._h, keys.ctypes.data, vals.ctypes.data, len(keys), kmer_bits,
                                               bstart.ctypes.data, len(bstart), C.byref(nb), C.byref(kept)))
        return keys[:kept.value], vals[:kept.value], bstart

    def write_fasta(self, out_dir):
        """The nodes in Graph2Fasta's format (Graph2Fasta.java:49-75), one part file per rank."""
        self._check(self._L.cb_write_fasta(self._h, str(out_dir).encode()))

    def time_ms(self, what):
        v = C.c_double()
        self._check(self._L.cb_get_time_ms(self._h, what.encode(), C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64()
        self._L.cb_get_launch_count(self._h, C.byref(v))
        return v.value
