"""Generates tests/golden/ec10k_survey.json and the gzip'd config-1 input fixture.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Sources of the pins (the reference ships no expected outputs):
  * /root/reference/data/Ec10k.sim.sfa            -- the reference's bundled example input
    (BASELINE.json configs[0]); stored gzip'd as tests/golden/ec10k_reads.sfa.gz so the GPU box,
    which has no /root/reference, can run config 1.
  * /root/reference/data/Ec10k.sim.sfa.ck{A,B}.pkl -- checkpoints written by the survey-time,
    independent Python restatement of the 12 jobs (plain dict/list/tuple pickles; loaded with an
    Unpickler that refuses every global).  They are a SECOND OPINION, not the reference itself.
The JSON holds counts and SHA-256 digests (SURVEY.md section 6 definition) of those checkpoints,
plus the full checkpoint-B edge list (small) for diagnostics.
"""
import gzip
import hashlib
import json
import os
import pickle

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"
TYPES = ("ff", "fr", "rf", "rr")


class NoGlobals(pickle.Unpickler):
    def find_class(self, module, name):
        raise pickle.UnpicklingError("globals are not allowed in golden pickles")


def flat(adj):
    rows = []
    for a, d in adj.items():
        for t in TYPES:
            for b, ov in d[t]:
                rows.append((a, t, b, int(ov)))
    rows.sort(key=lambda r: (r[0], TYPES.index(r[1]), r[2], r[3]))
    return rows


def digest(rows):
    h = hashlib.sha256()
    for a, t, b, ov in rows:
        h.update(f"{a}\t{t}\t{b}\t{ov}\n".encode())
    return h.hexdigest()[:16]


def main():
    sfa = open(os.path.join(REF, "Ec10k.sim.sfa"), "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "ec10k_reads.sfa.gz"), "wb", mtime=0) as f:
        f.write(sfa)
    A = NoGlobals(open(os.path.join(REF, "Ec10k.sim.sfa.ckA.pkl"), "rb")).load()
    B = NoGlobals(open(os.path.join(REF, "Ec10k.sim.sfa.ckB.pkl"), "rb")).load()
    nodes = A[0]
    node_rows = sorted((k, v[0], int(v[1])) for k, v in nodes.items())
    hn = hashlib.sha256()
    for i, s, c in node_rows:
        hn.update(f"{i}\t{s}\t{c}\n".encode())
    ea, ec, eb = flat(A[1]), flat(B[1]), flat(B[2])
    out = {
        "source": "survey-time independent Python restatement (not the JVM reference)",
        "input": "data/Ec10k.sim.sfa", "input_sha256": hashlib.sha256(sfa).hexdigest(),
        "k": 21, "readlen": 36,
        "reads": sfa.count(b"\n"),
        "nodes": len(nodes), "nodes_digest": hn.hexdigest()[:16],
        "cov_sum": sum(c for _, _, c in node_rows),
        "ckA_edges": len(ea), "ckA_digest": digest(ea),
        "chl2_edges": len(ec), "chl2_digest": digest(ec),
        "ckB_edges": len(eb), "ckB_digest": digest(eb),
        # SURVEY.md section 6 table, row 1 (same restatement)
        "survey_counters": {"redundant": 3493, "tuples": 264112, "candidates": 287376,
                            "edge_removal_1": 42600, "edge_removal_2": 15, "trans_edge": 102813},
        "ckB_edge_list": eb,
    }
    with open(os.path.join(HERE, "ec10k_survey.json"), "w") as f:
        json.dump(out, f)
    print({k: v for k, v in out.items() if k != "ckB_edge_list"})


if __name__ == "__main__":
    main()
