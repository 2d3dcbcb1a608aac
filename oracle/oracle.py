"""ctypes wrapper of the CPU parity oracle (oracle/brush_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (cloudbrush_b200/) never imports it.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbrush_oracle.so")
_lib = None

# checkpoint codes of bo_* (brush_oracle.cpp: checkpoint())
CK_PREPROCESS = 0   # 00-preprocess
CK_A = 1            # 01-overlap
CK_CHL2 = 2         # 01-overlap.chl2
CK_B = 3            # 01-overlap.re
CK_PREFIX = 10      # 00-preprocess.prefix (candidate edges)
CK_VO = 11          # 00-preprocess.vo
CK_TR = 12          # 01-overlap.tr
CK_PRE0 = 20        # 00-preprocess.0

EDGE_TYPES = ("ff", "fr", "rf", "rr")


def build(force=False):
    src = os.path.join(_HERE, "brush_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.bo_run.restype = C.c_void_p
        L.bo_run.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_long, C.c_long, C.c_float,
                             C.c_float, C.c_int]
        L.bo_error.restype = C.c_char_p
        L.bo_error.argtypes = [C.c_void_p]
        L.bo_counter.restype = C.c_int64
        L.bo_counter.argtypes = [C.c_void_p, C.c_char_p]
        for f in (L.bo_num_nodes, L.bo_num_edges):
            f.restype = C.c_size_t
            f.argtypes = [C.c_void_p, C.c_int]
        L.bo_get_nodes.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bo_get_edges.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bo_num_tuples.restype = C.c_size_t
        L.bo_num_tuples.argtypes = [C.c_void_p]
        L.bo_get_tuples.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.bo_node_text.restype = C.c_size_t
        L.bo_node_text.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.bo_str2dna.restype = C.c_size_t
        L.bo_str2dna.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.bo_dna2str.restype = C.c_size_t
        L.bo_dna2str.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.bo_consensus.restype = C.c_int
        L.bo_consensus.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_float, C.c_float,
                                   C.c_char_p, C.c_size_t]
        L.bo_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleRun:
    """One run of the restated pipeline over an SFA byte buffer."""

    def __init__(self, sfa: bytes, k: int, readlen: int, up_kmer=2000, low_kmer=1, majority=0.6, pwm_n=0.1,
                 stop_after=3):
        self._L = lib()
        self._sfa = bytes(sfa)
        self._h = self._L.bo_run(self._sfa, len(self._sfa), k, readlen, up_kmer, low_kmer, majority, pwm_n,
                                 stop_after)
        self.error = self._L.bo_error(self._h).decode()

    def close(self):
        if self._h:
            self._L.bo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counter(self, name):
        return int(self._L.bo_counter(self._h, name.encode()))

    def nodes(self, ck=CK_PREPROCESS):
        """(line int64[n], cov float32[n], contained int8[n]) sorted by input line."""
        n = self._L.bo_num_nodes(self._h, ck)
        line = np.empty(n, np.int64)
        cov = np.empty(n, np.float32)
        cont = np.empty(n, np.int8)
        if n:
            self._L.bo_get_nodes(self._h, ck, line.ctypes.data, cov.ctypes.data, cont.ctypes.data)
        return line, cov, cont

    def edges(self, ck):
        """int64[n,4] rows (node line, type 0..3, nbr line, ov), lexicographically sorted."""
        n = self._L.bo_num_edges(self._h, ck)
        node = np.empty(n, np.int64)
        typ = np.empty(n, np.int8)
        nbr = np.empty(n, np.int64)
        ov = np.empty(n, np.int32)
        if n:
            self._L.bo_get_edges(self._h, ck, node.ctypes.data, typ.ctypes.data, nbr.ctypes.data, ov.ctypes.data)
        return np.stack([node, typ.astype(np.int64), nbr, ov.astype(np.int64)], axis=1) if n else np.zeros((0, 4), np.int64)

    def tuples(self):
        """MatchPrefix map output: dict of arrays key,line,kind,dir,rev_idx,ov."""
        n = self._L.bo_num_tuples(self._h)
        key = np.empty(n, np.uint64)
        line = np.empty(n, np.int64)
        kind = np.empty(n, np.int8)
        d = np.empty(n, np.int8)
        rev = np.empty(n, np.int8)
        ov = np.empty(n, np.int32)
        if n:
            self._L.bo_get_tuples(self._h, key.ctypes.data, line.ctypes.data, kind.ctypes.data, d.ctypes.data,
                                  rev.ctypes.data, ov.ctypes.data)
        return dict(key=key, line=line, kind=kind, dir=d, rev_idx=rev, ov=ov)

    def node_text(self, ck):
        n = self._L.bo_node_text(self._h, ck, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._L.bo_node_text(self._h, ck, buf, n)
        return buf.raw[:n].decode()


def str2dna(s: str) -> str:
    buf = C.create_string_buffer(len(s) + 8)
    lib().bo_str2dna(s.encode(), buf, len(buf))
    return buf.value.decode()


def dna2str(s: str) -> str:
    buf = C.create_string_buffer(2 * len(s) + 8)
    lib().bo_dna2str(s.encode(), buf, len(buf))
    return buf.value.decode()


def consensus(entries, majority=0.6, threshold=0.1):
    """entries: list of (overhang, cov). Returns the consensus string or None."""
    n = len(entries)
    strs = (C.c_char_p * n)(*[e[0].encode() for e in entries])
    covs = (C.c_float * n)(*[float(e[1]) for e in entries])
    buf = C.create_string_buffer(max(len(e[0]) for e in entries) + 8 if n else 8)
    ok = lib().bo_consensus(n, strs, covs, majority, threshold, buf, len(buf))
    return buf.value.decode() if ok else None


def edge_digest(ids, edges):
    """SURVEY.md section 6 digest: first 16 hex of SHA-256 over lines 'id\\ttype\\tnbr\\tov\\n',
    nodes sorted by id string, types in order ff,fr,rf,rr, entries sorted by (nbr string, ov).
    `ids[line]` is the id string of input line `line`; `edges` rows are (line, type, nbr line, ov)."""
    rows = [(ids[int(a)], int(t), ids[int(b)], int(o)) for a, t, b, o in edges]
    rows.sort()
    h = hashlib.sha256()
    for a, t, b, o in rows:
        h.update(f"{a}\t{EDGE_TYPES[t]}\t{b}\t{o}\n".encode())
    return h.hexdigest()[:16]
