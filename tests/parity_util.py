"""Shared helpers of the parity tests: run the oracle and the CUDA path on one SFA buffer and
compare counters, nodes and the three edge checkpoints as canonicalised sets."""
import numpy as np

from oracle import oracle as O

COUNTERS = ["reads_good", "reads_goodbp", "contained", "nodecount", "nodes", "hkmer", "edge_removal_1",
            "edge_removal_2", "trans_edge", "contained_edge"]
OPTIONAL_ZERO = ["reads_short", "reads_skipped", "input_lines_invalid"]


def run_oracle(sfa, k, readlen, **kw):
    r = O.OracleRun(sfa, k, readlen, **kw)
    assert r.error == "", r.error
    return r


def run_gpu(sfa, k, readlen, **kw):
    from cloudbrush_b200 import BrushConfig, BrushGpu
    g = BrushGpu(BrushConfig(k=k, readlen=readlen, **kw))
    g.load_sfa_buffer(sfa)
    g.run()
    return g


def first_diff(a, b):
    sa = set(map(tuple, a.tolist()))
    sb = set(map(tuple, b.tolist()))
    return sorted(sa - sb)[:5], sorted(sb - sa)[:5], len(sa - sb), len(sb - sa)


def compare(g, r, label=""):
    """Returns a list of human-readable differences (empty == parity)."""
    diffs = []
    for c in COUNTERS:
        ov, gv = max(r.counter(c), 0), g.counter(c)  # a Hadoop counter never incremented is absent (-1 here)
        if ov != gv:
            diffs.append(f"{label} counter {c}: oracle {ov} gpu {gv}")
    for c in OPTIONAL_ZERO:
        ov, gv = max(r.counter(c), 0), g.counter(c)
        if ov != gv:
            diffs.append(f"{label} counter {c}: oracle {ov} gpu {gv}")
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    gn = g.nodes()
    if len(gn) != len(line) or not np.array_equal(gn["line"].astype(np.int64), line):
        diffs.append(f"{label} node set differs: oracle {len(line)} gpu {len(gn)}")
    elif not np.array_equal(gn["cov"].astype(np.int64), cov.astype(np.int64)):
        diffs.append(f"{label} node coverage differs at {int(np.sum(gn['cov'] != cov.astype(np.int64)))} nodes")
    for name, ock, gck in (("A", O.CK_A, 1), ("chl2", O.CK_CHL2, 2), ("B", O.CK_B, 3)):
        oe, ge = r.edges(ock), g.edge_rows(gck)
        if oe.shape != ge.shape or not np.array_equal(oe, ge):
            only_o, only_g, no, ng = first_diff(oe, ge)
            diffs.append(f"{label} checkpoint {name}: oracle {len(oe)} gpu {len(ge)} edges; "
                         f"only-oracle {no} e.g. {only_o}; only-gpu {ng} e.g. {only_g}")
    return diffs
