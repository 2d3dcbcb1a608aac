"""CPU tests of the parity oracle: golden pins, codec, Consensus, and one test per reference rule."""
import numpy as np
import pytest

from cases import all_cases, rc, rand_dna, sfa
from oracle import oracle as O

CASES = all_cases()


def run(name, **kw):
    buf, k, L = CASES[name]
    r = O.OracleRun(buf, k, L, **kw)
    assert r.error == "", r.error
    return r


def ids_of(buf):
    out = []
    for line in buf.replace(b"\r\n", b"\n").split(b"\n"):
        out.append(line.split(b"\t")[0].decode() if line else "")
    return out


def named_edges(name, ck):
    buf = CASES[name][0]
    ids = ids_of(buf)
    return {(ids[a], O.EDGE_TYPES[t], ids[b], ov) for a, t, b, ov in run(name).edges(ck).tolist()}


# ---------------------------------------------------------------------------------------------
def test_golden_ec10k(ec10k_sfa, ec10k_golden):
    """Pins the oracle against the survey-time independent restatement on BASELINE config 1."""
    r = O.OracleRun(ec10k_sfa, 21, 36)
    assert r.error == ""
    g = ec10k_golden
    ids = [l.split(b"\t")[0].decode() for l in ec10k_sfa.split(b"\n") if l]
    assert r.counter("reads_good") == g["reads"]
    assert r.counter("nodes") == g["nodes"]
    assert r.counter("RedundantRemoval:redundant") == g["survey_counters"]["redundant"]
    assert len(r.tuples()["key"]) == g["survey_counters"]["tuples"]
    assert len(r.edges(O.CK_PREFIX)) == g["survey_counters"]["candidates"]
    assert r.counter("edge_removal_1") == g["survey_counters"]["edge_removal_1"]
    assert r.counter("edge_removal_2") == g["survey_counters"]["edge_removal_2"]
    assert r.counter("trans_edge") == g["survey_counters"]["trans_edge"]
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    assert int(cov.sum()) == g["cov_sum"]
    for ck, key in ((O.CK_A, "ckA"), (O.CK_CHL2, "chl2"), (O.CK_B, "ckB")):
        e = r.edges(ck)
        assert len(e) == g[key + "_edges"]
        assert O.edge_digest(ids, e) == g[key + "_digest"]
    got = sorted((ids[a], O.EDGE_TYPES[t], ids[b], ov) for a, t, b, ov in r.edges(O.CK_B).tolist())
    assert got == sorted(tuple(x) for x in g["ckB_edge_list"])


def test_codec_known_values_and_roundtrip():
    # Node.java:90-133: A=A, B..E=AA..AT, F=C, G..J=CA..CT, K=G, L..O=GA..GT, P=T, Q..T=TA..TT
    assert O.str2dna("A") == "A" and O.str2dna("C") == "F" and O.str2dna("G") == "K" and O.str2dna("T") == "P"
    assert O.str2dna("AA") == "B" and O.str2dna("AT") == "E" and O.str2dna("CA") == "G" and O.str2dna("TT") == "T"
    assert O.str2dna("ACGT") == "CO"
    assert O.str2dna("ACGTA") == "COA" and O.str2dna("ACGTAC") == "COC" and O.str2dna("ACGTACG") == "COCK"
    for n in range(1, 70):
        s = rand_dna(n, n)
        assert O.dna2str(O.str2dna(s)) == s
        assert len(O.str2dna(s)) == (n + 1) // 2


def test_consensus_matches_integer_rule():
    """Node.Consensus' float tests (cnt/sum > 0.6f, N/len > 0.1f) equal 5*cnt > 3*sum and 10*N > len."""
    f = np.float32
    for s in range(1, 400):
        c = np.arange(0, s + 1)
        lhs = (c.astype(f) / f(s)) > f(0.6)
        assert np.array_equal(lhs, 5 * c > 3 * s)
        assert np.array_equal((c.astype(f) / f(s)) > f(0.1), 10 * c > s)
    # majority column, N column, null result
    assert O.consensus([("ACG", 1), ("ACT", 1), ("ACG", 1)]) == "ACG"       # 2/3 > 0.6 at the last column
    assert O.consensus([("AC", 1), ("AG", 1)]) is None                      # 1 N of 2 > 0.1
    assert O.consensus([("AAAAAAAAAAAC", 1), ("AAAAAAAAAAAG", 1)]) == "AAAAAAAAAAAN"  # 1/12 <= 0.1
    # consensus length: 2nd longest, or the 3rd when the two longest have coverage 1 each
    assert O.consensus([("ACGTAC", 1), ("ACGT", 1), ("AC", 1)]) == "AC"
    assert O.consensus([("ACGTAC", 2), ("ACGT", 1), ("AC", 1)]) == "ACGT"


def test_rc_duplicates_string_order_survivor():
    r = run("rc_duplicates")
    buf = CASES["rc_duplicates"][0]
    ids = ids_of(buf)
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    got = {ids[l]: int(c) for l, c in zip(line, cov)}
    assert got == {"10": 3, "7": 1}  # "10" < "8" < "9" as strings; three copies of one read
    assert r.counter("contained") == 2
    text = r.node_text(O.CK_PREPROCESS)
    s = buf.split(b"\n")[1].split(b"\t")[1].decode()  # "10" keeps ITS orientation (the rc)
    assert f"10\tN\t*s\t{O.str2dna(s)}\t*v\t3.00" in text


def test_self_reverse_complement_counts_twice():
    r = run("self_rc")
    ids = ids_of(CASES["self_rc"][0])
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    assert {ids[l]: int(c) for l, c in zip(line, cov)} == {"p1": 3, "q": 1}


def test_chain_transitive_reduction():
    a = named_edges("chain3", O.CK_A)
    b = named_edges("chain3", O.CK_B)
    assert ("r0", "ff", "r1", 35) in a and ("r1", "ff", "r2", 35) in a and ("r0", "ff", "r2", 30) in a
    assert ("r2", "rr", "r0", 30) in a  # mirror
    assert b == {("r0", "ff", "r1", 35), ("r1", "rr", "r0", 35), ("r1", "ff", "r2", 35), ("r2", "rr", "r1", 35)}
    assert run("chain3").counter("trans_edge") == 2


@pytest.mark.parametrize("seed", range(4))
def test_exact_k_overlap_found_from_one_side(seed):
    name = f"exact_k_{seed}"
    buf = CASES[name][0]
    a_seq, b_seq = [l.split(b"\t")[1].decode() for l in buf.split(b"\n")[:2]]
    X = b_seq[:21]
    vo = named_edges(name, O.CK_VO)
    # proposed a->b iff the shared k-mer is the NON-canonical orientation
    assert (("a", "ff", "b", 21) in vo) == (X > rc(X))
    assert (("b", "rr", "a", 21) in vo) == (rc(X) > X)
    assert named_edges(name, O.CK_A) == {("a", "ff", "b", 21), ("b", "rr", "a", 21)}


def test_tuple_kinds_and_poly_a():
    r = run("poly_a")
    t = r.tuples()
    # per read: one prefix NODEMSG, one suffix NODEMSG, SUFFIXMSG for windows 1..L-K-1 minus poly-A
    L, K = 40, 21
    per_read = {}
    for line, kind in zip(t["line"], t["kind"]):
        per_read.setdefault(int(line), []).append(int(kind))
    for line, kinds in per_read.items():
        assert kinds.count(1) == 1 and kinds.count(2) == 1
        assert kinds.count(0) <= L - K - 1
    seqs = [l.split(b"\t")[1].decode() for l in CASES["poly_a"][0].split(b"\n") if l]
    for i, s in enumerate(seqs):
        interior_poly = sum(1 for p in range(1, L - K) if s[p:p + K] in ("A" * K, "T" * K))
        assert per_read[i].count(0) == L - K - 1 - interior_poly
    assert any(s[p:p + K] == "A" * K for s in seqs for p in range(1, L - K))  # the case is not vacuous
    assert not np.any((t["key"] == 0) & (t["kind"] == 0))


def test_low_kmer_groups_give_no_edges():
    r = run("no_overlap")
    assert len(r.edges(O.CK_A)) == 0 and r.counter("nodes") == 5


def test_maximal_overlap_filter():
    vo = named_edges("periodic", O.CK_VO)
    ovs = sorted(ov for a, t, b, ov in vo if a == "m" and b == "n" and t == "ff")
    assert ovs == [32]  # 24 would also match (period 8) but the larger overlap wins
    cand = named_edges("periodic", O.CK_PREFIX)
    assert {ov for a, t, b, ov in cand if a == "n" and t == "rr" and b == "m"} >= {32, 24}


def test_branch_is_cut_by_consensus():
    r = run("branch")
    assert r.counter("edge_removal_1") > 0
    a = named_edges("branch", O.CK_A)
    c = named_edges("branch", O.CK_CHL2)
    removed = a - c
    assert removed and all("odd" in (x[0], x[2]) for x in removed)
    assert {(x[2], x[0]) for x in removed} == {(x[0], x[2]) for x in removed}  # both ends lose the edge


def test_prefix_equals_rc_of_suffix_merges_records():
    r = run("prefix_rc_suffix")
    assert r.error == ""
    text = r.node_text(O.CK_PREFIX)
    assert sum(1 for l in text.split("\n") if l.startswith("h\t")) == 1  # one record: both entries share a group


def test_line_format_rules():
    r = run("line_formats")
    assert r.counter("GenNonContainedReads:input_lines_invalid") == 3  # one field, empty line, four fields
    assert r.counter("GenNonContainedReads:reads_skipped") == 1        # N
    assert r.counter("GenNonContainedReads:reads_short") == 1
    assert r.counter("reads_good") == 7                                # ok1 lower crlf qual tabs "" last
    ids = ids_of(CASES["line_formats"][0])
    line, cov, _ = r.nodes(O.CK_PREPROCESS)
    assert {ids[l] for l in line} == {"ok1", "lower", "crlf", "qual", "tabs", "", "last"}


def test_symmetry_and_nesting_of_checkpoints():
    for name in ("tiling_both_strands", "branch", "chain3"):
        r = run(name)
        a = {tuple(x) for x in r.edges(O.CK_A).tolist()}
        c = {tuple(x) for x in r.edges(O.CK_CHL2).tolist()}
        b = {tuple(x) for x in r.edges(O.CK_B).tolist()}
        flip = {0: 3, 1: 1, 2: 2, 3: 0}  # Node.flip_link
        for s in (a, c, b):
            assert {(y, flip[t], x, ov) for x, t, y, ov in s} == s
        assert b <= c <= a
