"""Hand-built micro inputs, each exercising one rule of the reference."""
import random

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def rand_dna(n, seed):
    r = random.Random(seed)
    return "".join(r.choice("ACGT") for _ in range(n))


def sfa(reads):
    """reads: list of (id, seq) -> SFA bytes"""
    return "".join(f"{i}\t{s}\n" for i, s in reads).encode()


def tiled_reads(genome, L, step, ids=None, flip=()):
    out = []
    for n, p in enumerate(range(0, len(genome) - L + 1, step)):
        s = genome[p:p + L]
        if n in flip:
            s = rc(s)
        out.append(((ids[n] if ids else f"r{n}"), s))
    return out


def all_cases():
    """name -> (sfa bytes, k, readlen)"""
    cases = {}
    g = rand_dna(400, 11)
    # 1. three-read chain a -> b -> c: TransitiveReduction removes a -> c and its mirror
    cases["chain3"] = (sfa(tiled_reads(g, 40, 5)[:3]), 21, 40)
    # 2. a longer tiling with both strands and the "10" < "9" id trap
    ids = [format(i + 1, "X") for i in range(200)]
    cases["tiling_both_strands"] = (sfa(tiled_reads(g, 40, 3, ids=ids, flip={1, 4, 5, 9, 12, 20, 33})), 21, 40)
    # 3. duplicates incl. reverse complements: ids "10" and "9" (string order: "10" survives)
    s = g[10:50]
    cases["rc_duplicates"] = (sfa([("9", s), ("10", rc(s)), ("7", g[13:53]), ("8", s)]), 21, 40)
    # 4. a read equal to its own reverse complement, twice (cov = 1 + 2*dups)
    half = rand_dna(20, 5)
    pal = half + rc(half)
    cases["self_rc"] = (sfa([("p1", pal), ("p2", pal), ("q", rand_dna(40, 6))]), 21, 40)
    # 5. exact-K overlaps (discovered from one side only, mirrored by GenReverseEdge)
    for seed in range(4):
        gg = rand_dna(59, 100 + seed)
        cases[f"exact_k_{seed}"] = (sfa([("a", gg[0:40]), ("b", gg[19:59])]), 21, 40)
    # 6. internal poly-A k-mer: the window is not emitted, the overlap through it is not proposed
    pa = rand_dna(12, 7) + "A" * 23 + rand_dna(25, 8)
    cases["poly_a"] = (sfa([("x", pa[0:40]), ("y", pa[5:45]), ("z", pa[12:52]), ("w", pa[20:60])]), 21, 40)
    # 7. unrelated reads: every k-mer group has one tuple (LOW_KMER), no edges
    cases["no_overlap"] = (sfa([(f"u{i}", rand_dna(40, 50 + i)) for i in range(5)]), 21, 40)
    # 8. two valid overlaps between one pair (period-4 sequence): maximal filter keeps the larger
    unit = "ACGTTGCA"
    per = (unit * 12)
    cases["periodic"] = (sfa([("m", rand_dna(8, 9) + per[:32]), ("n", per[:32] + rand_dna(8, 10))]), 21, 40)
    # 9. repeat boundary: three reads continue one way after a shared segment, one read another way
    shared = rand_dna(30, 21)
    left = rand_dna(30, 22)
    right1 = rand_dna(30, 23)
    right2 = rand_dna(30, 24)
    right2 = ("A" if right1[0] != "A" else "C") + right2[1:]  # the two continuations differ at once
    ga, gb = left + shared + right1, left + shared + right2
    # s0 ends inside the shared segment; t2, t3 and "odd" reach the first differing base
    reads = [("s0", ga[15:55])] + [(f"t{i}", ga[18 + 2 * i:58 + 2 * i]) for i in range(4)] + [("odd", gb[21:61])]
    cases["branch"] = (sfa(reads), 21, 40)
    # 10. read whose prefix k-mer is the reverse complement of its suffix k-mer (one group)
    pre = rand_dna(21, 31)
    cases["prefix_rc_suffix"] = (sfa([("h", pre + rand_dna(4, 32) + rc(pre)), ("i", rand_dna(6, 33) + pre + rand_dna(19, 34)),
                                      ("j", rand_dna(19, 35) + pre + rand_dna(6, 36))]), 21, 46)
    # 11. line-format edge cases: invalid field counts, lowercase, non-ACGT, short reads, CRLF,
    #     a third (quality) field, trailing tabs, empty lines, a last line without terminator
    body = (b"ok1\t" + g[0:40].encode() + b"\n" + b"bad_one_field\n" + b"\n" + b"lower\t" + g[3:43].lower().encode() + b"\n"
            + b"hasN\t" + (g[6:30] + "N" + g[31:46]).encode() + b"\n" + b"short\tACGTACGT\n"
            + b"crlf\t" + g[9:49].encode() + b"\r\n" + b"qual\t" + g[12:52].encode() + b"\tIIIIIIII\n"
            + b"tabs\t" + g[15:55].encode() + b"\t\t\n" + b"four\t" + g[18:58].encode() + b"\tx\ty\n"
            + b"\t" + g[21:61].encode() + b"\n" + b"last\t" + g[24:64].encode())
    cases["line_formats"] = (body, 21, 40)
    # 12. a repeated k-mer in the INTERIOR of 700 reads (one k-mer group of 700+ tuples spanning several
    #     256-tuple blocks of the group walk) plus reads that start / end with it (node entries of that group)
    rep = rand_dna(21, 41)
    reads = [(f"i{n}", rand_dna(7, 1000 + n) + rep + rand_dna(8, 3000 + n)) for n in range(700)]
    reads += [(f"p{n}", rep + rand_dna(15, 5000 + n)) for n in range(3)]
    reads += [(f"s{n}", rand_dna(15, 6000 + n) + rep) for n in range(3)]
    cases["long_group"] = (sfa(reads), 21, 36)
    # 13. a repeat boundary with many branches: 40 reads end with the k-mer, 40 start with it (exact-K overlaps
    #     between all of them: 40 node entries in one group, adjacency lists of 40 inconsistent overhangs)
    rep2 = rand_dna(21, 42)
    reads = [(f"e{n}", rand_dna(15, 7000 + n) + rep2) for n in range(40)]
    reads += [(f"b{n}", rep2 + rand_dna(15, 8000 + n)) for n in range(40)]
    cases["many_branches"] = (sfa(reads), 21, 36)
    return cases


def huge_group_case():
    """One k-mer in the interior of 2500 reads (a group larger than a hash bucket of the bucket kernel holds:
    the library must fall back to the flat pipeline on its own) plus node entries of that group.  -kmerup is
    raised so that the input stays inside the contract.  Returns (sfa, k, readlen, up_kmer)."""
    rep = rand_dna(21, 44)
    reads = [(f"i{n}", rand_dna(7, 11000 + n) + rep + rand_dna(8, 14000 + n)) for n in range(2500)]
    reads += [(f"p{n}", rep + rand_dna(15, 17000 + n)) for n in range(3)]
    reads += [(f"s{n}", rand_dna(15, 18000 + n) + rep) for n in range(3)]
    return sfa(reads), 21, 36, 100000


def upkmer_case():
    """A k-mer group that has node entries and MORE tuples than a (small) -kmerup, while BuildHighKmerList's
    weighted count over windows [0, L-K) stays below it (the last-window tuples are not counted there,
    BuildHighKmerList.java:59-60): the reference truncates the candidate list in Hadoop's value order
    (MatchPrefix.java:366-368).  Returns (sfa, k, readlen, up_kmer)."""
    rep = rand_dna(21, 43)
    reads = [(f"s{n}", rand_dna(15, 9000 + n) + rep) for n in range(30)]        # last window: not counted
    reads += [(f"p{n}", rep + rand_dna(15, 9100 + n)) for n in range(3)]        # window 0
    reads += [(f"i{n}", rand_dna(7, 9200 + n) + rep + rand_dna(8, 9300 + n)) for n in range(10)]
    return sfa(reads), 21, 36, 20
