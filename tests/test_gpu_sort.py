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
