"""The CPU oracle against the golden vectors of the INDEPENDENT closed-form restatement
(tests/golden/restate.py -> tests/golden/restate_goldens.json): counts, SHA-256 digests of the node table and
of the three checkpoints, and the cut / reduction counters, on the reference's bundled input, every micro
case and reduced genomes of configs 2, 3 and 5.  `-m gpu`: the CUDA path against the same vectors."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import restate  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "restate_goldens.json")))
INPUTS = restate.golden_inputs()


def _ids_and_seqs(buf):
    lines = buf.replace(b"\r\n", b"\n").replace(b"\r", b"\n").split(b"\n")
    return [l.split(b"\t")[0].decode() for l in lines], [(l.split(b"\t") + [b""])[1].decode().upper() for l in lines]


def _edge_digest(ids, rows):
    e = {(ids[a], restate.TYPES[t], ids[b], int(ov)) for a, t, b, ov in rows.tolist()}
    return len(e), restate.edge_digest(e)


def _node_digest(ids, seqs, line, cov):
    h = hashlib.sha256()
    for rid, s, c in sorted((ids[l], seqs[l], int(c)) for l, c in zip(line.tolist(), cov.tolist())):
        h.update(f"{rid}\t{s}\t{c}\n".encode())
    return h.hexdigest()[:16]


def test_goldens_cover_the_inputs():
    assert sorted(GOLD) == sorted(INPUTS) and len(GOLD) >= 20
    assert GOLD["ec10k"]["ckA_digest"] == "4cd804917c7896c3"  # == the survey-time pickles (tests/golden/ec10k_survey.json)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_independent_restatement(name):
    from oracle import oracle as O
    buf, k = INPUTS[name]
    want = GOLD[name]
    ids, seqs = _ids_and_seqs(buf)
    L = len(next(s for s in seqs if len(s) > k and set(s) <= set("ACGT")))
    if name == "case_line_formats":
        L = 40
    r = O.OracleRun(buf, k, L)
    assert r.error == "", r.error
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    assert len(line) == want["nodes"] and _node_digest(ids, seqs, line, cov) == want["nodes_digest"]
    for key, ck in (("ckA", O.CK_A), ("chl2", O.CK_CHL2), ("ckB", O.CK_B)):
        assert _edge_digest(ids, r.edges(ck)) == (want[key], want[key + "_digest"]), key
    for c in ("edge_removal_1", "edge_removal_2", "trans_edge", "contained_edge"):
        assert max(r.counter(c), 0) == want[c], c
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD))
def test_cuda_path_matches_independent_restatement(name):
    from cloudbrush_b200 import BrushConfig, BrushGpu
    buf, k = INPUTS[name]
    want = GOLD[name]
    ids, seqs = _ids_and_seqs(buf)
    L = 40 if name == "case_line_formats" else len(next(s for s in seqs if len(s) > k and set(s) <= set("ACGT")))
    g = BrushGpu(BrushConfig(k=k, readlen=L))
    g.load_sfa_buffer(buf)
    g.run()
    n = g.nodes()
    assert len(n) == want["nodes"]
    assert _node_digest(ids, seqs, n["line"].astype(np.int64), n["cov"].astype(np.int64)) == want["nodes_digest"]
    for key, ck in (("ckA", 1), ("chl2", 2), ("ckB", 3)):
        assert _edge_digest(ids, g.edge_rows(ck)) == (want[key], want[key + "_digest"]), key
    for c in ("edge_removal_1", "edge_removal_2", "trans_edge", "contained_edge"):
        assert g.counter(c) == want[c], c
    g.close()
