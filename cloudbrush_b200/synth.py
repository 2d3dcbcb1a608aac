"""Synthetic SFA inputs of BASELINE.md section 3 (deterministic, numpy Philox).

genome: i.i.d. uniform ACGT of length G, seed = config number; N = round(G*cov/L) error-free reads
at uniform start positions; optional reverse complement with probability 1/2; ids are '%X' of the
1-based read index (data/preprocessor.pl:42 of the reference, -renum).
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_HEX = np.frombuffer(b"0123456789ABCDEF", dtype=np.uint8)

CONFIGS = {
    # name: (genome length, read length, coverage, both strands, k, seed)
    "cfg2": (4_600_000, 36, 100, False, 21, 2),
    "cfg3": (4_600_000, 100, 50, True, 31, 3),
    "cfg4": (100_000_000, 100, 50, True, 31, 4),
    "cfg5": (500_000_000, 150, 30, True, 31, 5),
    # reduced genomes of the same shape, for parity tests against the CPU oracle
    "cfg2_50k": (50_000, 36, 100, False, 21, 2),
    "cfg3_50k": (50_000, 100, 50, True, 31, 3),
    "cfg5_50k": (50_000, 150, 30, True, 31, 5),
}


N_PARTS = 8  # the read set is always drawn as 8 independent streams, so it does not depend on the GPU count


def part_sizes(n_reads):
    base, rem = divmod(n_reads, N_PARTS)
    return [base + (1 if p < rem else 0) for p in range(N_PARTS)]


def synth_reads(genome_len, read_len, coverage, both_strands, seed, n_reads=None, parts=None):
    """codes uint8[n, L] in 0..3 of the requested parts (default: all 8, i.e. the whole read set).

    Part p uses its own Philox stream (jumped p+1 times), so rank r of W ranks can generate parts
    [r*8/W, (r+1)*8/W) on its own and the union over the ranks is the single-GPU read set."""
    genome = np.random.Generator(np.random.Philox(seed)).integers(0, 4, size=genome_len, dtype=np.uint8)
    n = int(round(genome_len * coverage / read_len)) if n_reads is None else n_reads
    sizes = part_sizes(n)
    out = []
    for p in (range(N_PARTS) if parts is None else parts):
        rng = np.random.Generator(np.random.Philox(seed).jumped(p + 1))
        starts = rng.integers(0, genome_len - read_len + 1, size=sizes[p], dtype=np.int64)
        idx = starts[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
        reads = genome[idx]
        if both_strands:
            flip = rng.integers(0, 2, size=sizes[p], dtype=np.uint8).astype(bool)
            reads = np.where(flip[:, None], 3 - reads[:, ::-1], reads)
        out.append(np.ascontiguousarray(reads, dtype=np.uint8))
    return np.concatenate(out, axis=0) if out else np.zeros((0, read_len), np.uint8)


def sfa_bytes(codes, first_id=1):
    """SFA text 'HEXID\\tSEQ\\n' for every row of `codes` (uint8 array of the whole buffer)."""
    n, L = codes.shape
    ids = np.arange(first_id, first_id + n, dtype=np.int64)
    ndig = np.ones(n, dtype=np.int64)
    v = ids >> 4
    while np.any(v > 0):
        ndig += (v > 0)
        v >>= 4
    line_len = ndig + 1 + L + 1
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(line_len, out=offs[1:])
    out = np.empty(int(offs[-1]), dtype=np.uint8)
    seq = _ACGT[codes]
    for d in range(1, int(ndig.max()) + 1):
        sel = np.nonzero(ndig == d)[0]
        if len(sel) == 0:
            continue
        base = offs[sel]
        for j in range(d):  # hex digits, most significant first
            out[base + j] = _HEX[(ids[sel] >> (4 * (d - 1 - j))) & 15]
        out[base + d] = 9  # '\t'
        pos = (base + d + 1)[:, None] + np.arange(L, dtype=np.int64)[None, :]
        out[pos] = seq[sel]
        out[base + d + 1 + L] = 10  # '\n'
    return out


def make_config(name, scale_genome=None, rank=0, world=1):
    """SFA bytes of rank's shard (consecutive lines of the whole read set) and the workload's description."""
    G, L, cov, both, k, seed = CONFIGS[name]
    if scale_genome is not None:
        G = scale_genome
    n_total = int(round(G * cov / L))
    per = N_PARTS // world
    if per * world != N_PARTS:
        raise ValueError("world must divide 8")
    parts = list(range(rank * per, (rank + 1) * per))
    codes = synth_reads(G, L, cov, both, seed, parts=parts)
    first_id = 1 + sum(part_sizes(n_total)[:rank * per])
    return sfa_bytes(codes, first_id=first_id), dict(genome=G, readlen=L, coverage=cov, both_strands=both, k=k,
                                                     seed=seed, n_reads=codes.shape[0], n_reads_total=n_total)
