"""-m gpu: kernel-level parity of the radix sort that replaces the Hadoop shuffle/sort
(MatchPrefix.java:468) against numpy's stable sort, through the cb_debug_sort test hook."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ctx():
    from cloudbrush_b200 import BrushConfig, BrushGpu
    return BrushGpu(BrushConfig())


def _check(g, keys, vals, b0, b1):
    k, v = keys.copy(), (vals.copy() if vals is not None else None)
    g.debug_sort(k, v, b0, b1)
    mask = np.uint64(((1 << (b1 - b0)) - 1) << b0) if b1 - b0 < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(k, keys[order])
    if vals is not None:
        assert np.array_equal(v, vals[order])  # stable: equal keys keep their input order


@pytest.mark.parametrize("n", [0, 1, 2, 4097, 1_000_003])
@pytest.mark.parametrize("pairs", [True, False])
def test_sort_random_equal_sorted(n, pairs):
    rng = np.random.default_rng(n + 7)
    g = _ctx()
    idx = np.arange(n, dtype=np.uint32)
    for name, keys in (("random", rng.integers(0, 1 << 62, size=n, dtype=np.uint64)),
                       ("all_equal", np.full(n, 0x123456789AB, dtype=np.uint64)),
                       ("sorted", np.sort(rng.integers(0, 1 << 42, size=n, dtype=np.uint64))),
                       ("few_distinct", rng.integers(0, 5, size=n, dtype=np.uint64) << np.uint64(17)),
                       ("all_ones_tail", np.where(rng.random(n) < 0.1, np.uint64(0xFFFFFFFFFFFFFFFF),
                                                  rng.integers(0, 1 << 42, size=n, dtype=np.uint64)))):
        for b0, b1 in ((0, 42), (0, 62), (0, 64), (21, 24), (0, 36)):
            _check(g, keys, idx if pairs else None, b0, b1)
    g.close()


def test_sort_full_tuple_count():
    """69 M pairs -- the tuple count of BASELINE config 2 (16 850 tiles of the decoupled look-back)."""
    n = 69_014_800
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 42, size=n, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32)
    g = _ctx()
    k, v = keys.copy(), vals.copy()
    g.debug_sort(k, v, 0, 42)
    g.close()
    assert np.all(k[1:] >= k[:-1])
    assert np.array_equal(keys[v], k)                 # a permutation that carries its payload
    same = k[1:] == k[:-1]
    assert np.all(v[1:][same] > v[:-1][same])         # stable


def _kmer_hash(x):
    x = (x * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(32)
    return (x * np.uint64(0xD6E8FEB86659FD93)) & np.uint64(0xFFFFFFFFFFFFFFFF)


@pytest.mark.parametrize("n,kind", [(0, "random"), (1, "random"), (4097, "random"), (300_000, "random"),
                                    (300_000, "few_kmers"), (300_000, "invalid_tail"), (5_000_000, "random")])
def test_hash_partition_buckets(n, kind):
    """The shuffle replacement of the default path (primitives.cu: bucket_partition): the tuples come back as a
    permutation of the valid input (all-ones keys dropped), every bucket holds exactly the tuples whose k-mer hash
    has the bucket's top bits, and payload bits above the k-mer in the key do not influence the bucket."""
    rng = np.random.default_rng(n + len(kind))
    kbits = 42
    kmer = rng.integers(0, 1 << kbits, size=n, dtype=np.uint64)
    if kind == "few_kmers":
        kmer = rng.integers(0, 50, size=n, dtype=np.uint64) * np.uint64(0x9E3779B1)
    high = rng.integers(0, 1 << 20, size=n, dtype=np.uint64) << np.uint64(kbits)  # payload bits above the k-mer
    keys = kmer | high
    if kind == "invalid_tail":
        keys[rng.random(n) < 0.2] = np.uint64(0xFFFFFFFFFFFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    g = _ctx()
    with np.errstate(over="ignore"):
        k, v, bstart = g.debug_partition(keys.copy(), vals.copy(), kbits)
        g.close()
        valid = keys != np.uint64(0xFFFFFFFFFFFFFFFF)
        assert len(k) == int(valid.sum()) == int(bstart[-1]) and bstart[0] == 0
        assert np.all(np.diff(bstart.astype(np.int64)) >= 0)
        assert np.array_equal(np.sort(v), vals[valid])          # a permutation of the valid tuples ...
        assert np.array_equal(keys[v], k)                       # ... that carries its key
        nb = len(bstart) - 1
        bits = nb.bit_length() - 1
        want = (_kmer_hash(k & np.uint64((1 << kbits) - 1)) >> np.uint64(64 - bits)).astype(np.int64)
        got = np.searchsorted(bstart.astype(np.int64), np.arange(len(k)), side="right") - 1
        assert np.array_equal(got, want)
