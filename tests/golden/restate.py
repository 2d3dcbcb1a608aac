"""An INDEPENDENT restatement of the hot path in plain Python, used only to generate golden vectors.

It is deliberately not a port of oracle/brush_oracle.cpp: the oracle replays the twelve MapReduce jobs
message by message; this file computes the two checkpoints from their closed-form definitions
(SURVEY.md section 8a') and restates the two per-list scans (CutChimericLinks + Node.Consensus,
TransitiveReduction) directly on adjacency sets.  Agreement of the two on the reference's bundled input
and on every micro case is the pin the reference itself cannot give (it ships no expected outputs and no
JVM exists in this image or on the GPU box: profiles/r02_java_check.txt).

Reference lines restated here (relative to /root/reference/src/Brush):
  nodes           GenNonContainedReads.java:58-130,142-251; RedundantRemoval.java:51-103
  checkpoint A    MatchPrefix.java:107-193,241-431; VerifyOverlap.java:203-344; GenReverseEdge.java:149-244
  cut             CutChimericLinks.java:210-256,258-409; Node.java:1293-1377; driver BrushAssembler.java:347-367
  removal         EdgeRemoval.java:121-202
  reduction       TransitiveReduction.java:175-221,300-409

    python tests/golden/restate.py            # regenerates tests/golden/restate_goldens.json

Supported contract (same as the library): equal-length reads, odd k, empty stopword list, no k-mer group
above -kmerup.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TYPES = ("ff", "fr", "rf", "rr")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def flip(d):
    return "r" if d == "f" else "f"


def flip_link(t):  # Node.java:2071-2078
    return flip(t[1]) + flip(t[0])


def parse_sfa(buf, K):
    """Hadoop LineReader + GenNonContainedReadsMapper filters -> [(line index, id, seq)] of the good reads."""
    text = buf.decode("latin-1")
    lines, cur, i = [], [], 0
    while i < len(text):  # a line ends at \n, \r\n or a lone \r
        c = text[i]
        if c == "\n" or c == "\r":
            lines.append("".join(cur))
            cur = []
            if c == "\r" and i + 1 < len(text) and text[i + 1] == "\n":
                i += 1
        else:
            cur.append(c)
        i += 1
    if cur:
        lines.append("".join(cur))
    good = []
    for n, line in enumerate(lines):
        f = line.split("\t")
        while f and f[-1] == "":  # Java String.split drops trailing empty strings
            f.pop()
        if len(f) not in (2, 3):
            continue
        seq = f[1].upper()
        if any(c not in "ACGT" for c in seq) or len(seq) <= K:
            continue
        good.append((n, f[0], seq))
    return good


def build_nodes(good):
    """Among reads identical up to reverse complement the smallest id (String.compareTo) survives in its own
    orientation; cov = class size (a read equal to its own reverse complement counts every duplicate twice)."""
    classes = {}
    for n, rid, seq in good:
        classes.setdefault(min(seq, rc(seq)), []).append((rid, n, seq))
    nodes = {}
    for members in classes.values():
        rid, n, seq = min(members)  # Java compares the id strings; ids are unique
        dup = len(members) - 1
        nodes[rid] = (seq, 1 + dup * (2 if seq == rc(seq) else 1), n)
    return nodes


def checkpoint_a(nodes, K):
    """S = {(a, xy, b, max ov proposed)}; A = S u mirror(S)   (SURVEY.md 8a' item 2)"""
    L = len(next(iter(nodes.values()))[0])
    orient = {}
    starts = {}
    for rid, (seq, _, _) in nodes.items():
        for d, s in (("f", seq), ("r", rc(seq))):
            orient[(rid, d)] = s
            starts.setdefault(s[:K], []).append((rid, d))
    S = {}
    for (a, x), s in orient.items():
        for ov in range(K, L):
            X = s[L - ov:L - ov + K]
            if ov > K:
                if X == "A" * K or X == "T" * K:
                    continue
            elif not X > rc(X):
                continue
            for b, y in starts.get(X, ()):
                if a != b and s[L - ov:] == orient[(b, y)][:ov]:
                    key = (a, x + y, b)
                    if S.get(key, 0) < ov:
                        S[key] = ov
    A = set()
    for (a, t, b), ov in S.items():
        A.add((a, t, b, ov))
        A.add((b, flip_link(t), a, ov))
    return A


def direction_lists(edges):
    out = {}
    for a, t, b, ov in edges:
        out.setdefault((a, t[0]), []).append((b, t, ov))
    return out


def sort_entries(entries, d):
    """OvelapSizeComparator_f / _r: overlap descending, then id ascending (f) or descending (r)."""
    entries = sorted(entries, key=lambda e: e[0], reverse=(d == "r"))
    return sorted(entries, key=lambda e: -e[2])  # stable


def overhang(nodes, b, t, ov):
    s = nodes[b][0]
    return (rc(s) if t[1] == "r" else s)[ov:]


def consensus(items, majority, pwm_n):
    """Node.Consensus (Node.java:1293-1377) on [(overhang, weight)] already in the caller's order."""
    items = sorted(items, key=lambda it: -len(it[0]))  # stable: legacy merge sort with the tie-blind comparator
    if len(items) == 2 or items[0][1] + items[1][1] > 2:
        clen = len(items[1][0])
    else:
        clen = len(items[2][0])
    cons, n_count = [], 0
    for j in range(clen):
        cnt = {"A": 0, "T": 0, "C": 0, "G": 0}
        tot = 0
        for h, w in items:
            if j < len(h):
                tot += w
                cnt[h[j]] += w
        for base in "ATCG":  # the reference tests A, T, C, G in this order, in float arithmetic
            if tot and np.float32(cnt[base]) / np.float32(tot) > np.float32(majority):
                cons.append(base)
                break
        else:
            cons.append("N")
            n_count += 1
    if clen and np.float32(n_count) / np.float32(clen) > np.float32(pwm_n):
        return None
    return "".join(cons)


def cut_round(nodes, edges, majority, pwm_n):
    """One CutChimericLinks job: the removal messages and the edge_removal counter."""
    dead, count = set(), 0
    for (a, d), entries in direction_lists(edges).items():
        if len(entries) < 2:
            continue
        entries = sort_entries(entries, d)
        hs = [(overhang(nodes, b, t, ov), int(nodes[b][1])) for b, t, ov in entries]
        cons = consensus(hs, majority, pwm_n)
        if cons is None:
            continue
        for (b, t, ov), (h, _) in zip(entries, hs):
            if any(h[j] != cons[j] and cons[j] != "N" for j in range(min(len(h), len(cons)))):
                dead.add((a, t, b, ov))
                dead.add((b, flip_link(t), a, ov))
                count += 1
    return dead, count


def transitive_reduction(nodes, edges):
    """Greedy scan per (node, direction); an edge survives iff both endpoints keep it."""
    L = len(next(iter(nodes.values()))[0])
    kept, trans, contained = set(), 0, 0
    for (a, d), entries in direction_lists(edges).items():
        prefixes, ids = [], {}
        for b, t, ov in sort_entries(entries, d):
            if ov == L:
                contained += 1
            if b in ids.get(t, ()):
                continue  # an edge of this type to b is already kept (maximal-overlap filter)
            mine = overhang(nodes, b, t, ov)
            if any(not (pov == ov) and mine.startswith(ph) for pov, ph in prefixes):  # equal-length reads
                trans += 1
                continue
            prefixes.append((ov, mine))
            ids.setdefault(t, set()).add(b)
            kept.add((a, t, b, ov))
    return {e for e in kept if (e[2], flip_link(e[1]), e[0], e[3]) in kept}, trans, contained


def run(buf, K, majority=0.6, pwm_n=0.1):
    good = parse_sfa(buf, K)
    nodes = build_nodes(good)
    A = checkpoint_a(nodes, K)
    cur, removal = set(A), []
    for rnd in (1, 2):  # BrushAssembler.java:347-367
        dead, n = cut_round(nodes, cur, majority, pwm_n)
        removal.append(n)
        if rnd > 1 and n == 0:
            break
        cur -= dead
        if n == 0:
            break
    B, trans, contained = transitive_reduction(nodes, cur)
    return dict(good=good, nodes=nodes, A=A, chl2=cur, B=B, edge_removal=removal + [0] * (2 - len(removal)),
                trans_edge=trans, contained_edge=contained)


def edge_digest(edges):
    """sha256 over 'node\\ttype\\tnbr\\tov\\n' sorted by (node id, type, nbr id, ov)."""
    h = hashlib.sha256()
    for a, t, b, ov in sorted(edges, key=lambda e: (e[0], TYPES.index(e[1]), e[2], e[3])):
        h.update(f"{a}\t{t}\t{b}\t{ov}\n".encode())
    return h.hexdigest()[:16]


def summary(r):
    hn = hashlib.sha256()
    for rid, (seq, cov, _) in sorted(r["nodes"].items()):
        hn.update(f"{rid}\t{seq}\t{cov}\n".encode())
    return {"reads_good": len(r["good"]), "nodes": len(r["nodes"]), "nodes_digest": hn.hexdigest()[:16],
            "ckA": len(r["A"]), "ckA_digest": edge_digest(r["A"]), "chl2": len(r["chl2"]),
            "chl2_digest": edge_digest(r["chl2"]), "ckB": len(r["B"]), "ckB_digest": edge_digest(r["B"]),
            "edge_removal_1": r["edge_removal"][0], "edge_removal_2": r["edge_removal"][1],
            "trans_edge": r["trans_edge"], "contained_edge": r["contained_edge"]}


def golden_inputs():
    """name -> (sfa bytes, k): the reference's bundled input, every micro case, one config-3-shaped genome"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cases import all_cases
    from cloudbrush_b200 import synth
    out = {"ec10k": (gzip.open(os.path.join(HERE, "ec10k_reads.sfa.gz")).read(), 21)}
    for name, (buf, k, _) in sorted(all_cases().items()):
        out["case_" + name] = (buf, k)
    for cfg, genome in (("cfg2_50k", 5_000), ("cfg3_50k", 20_000), ("cfg5_50k", 8_000)):
        buf, m = synth.make_config(cfg, scale_genome=genome)
        out[f"{cfg}_{genome}"] = (bytes(buf), m["k"])
    return out


def main():
    res = {}
    for name, (buf, k) in golden_inputs().items():
        res[name] = summary(run(buf, k))
        print(name, res[name], flush=True)
    with open(os.path.join(HERE, "restate_goldens.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
