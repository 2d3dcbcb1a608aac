"""Parity of the CUDA path (through the C ABI) against the CPU oracle -- the -m gpu tests proper."""
import numpy as np
import pytest

from parity_util import compare, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def _synth(name, genome):
    from cloudbrush_b200 import synth
    buf, meta = synth.make_config(name, scale_genome=genome)
    return bytes(buf), meta


def test_ec10k_config1_parity_and_golden(ec10k_sfa, ec10k_golden):
    """BASELINE config 1: bit-exact against the oracle and against the survey-time digests."""
    from oracle import oracle as O
    r = run_oracle(ec10k_sfa, 21, 36)
    g = run_gpu(ec10k_sfa, 21, 36)
    assert compare(g, r, "ec10k") == []
    ids = [l.split(b"\t")[0].decode() for l in ec10k_sfa.split(b"\n") if l]
    assert O.edge_digest(ids, g.edge_rows(1)) == ec10k_golden["ckA_digest"]
    assert O.edge_digest(ids, g.edge_rows(2)) == ec10k_golden["chl2_digest"]
    assert O.edge_digest(ids, g.edge_rows(3)) == ec10k_golden["ckB_digest"]
    assert g.counter("nodes") == ec10k_golden["nodes"]
    g.close()


@pytest.mark.parametrize("name,genome", [("cfg2_50k", 20_000), ("cfg3_50k", 20_000), ("cfg5_50k", 20_000)])
def test_synthetic_configs_parity(name, genome):
    """Reduced-genome versions of BASELINE configs 2, 3 and 5 (same read length, coverage, k); the default
    hash-bucket path and the flat pipeline (CB_FLAG_FLAT_PIPELINE)."""
    sfa, meta = _synth(name, genome)
    r = run_oracle(sfa, meta["k"], meta["readlen"])
    for flat in (False, True):
        g = run_gpu(sfa, meta["k"], meta["readlen"], flat_pipeline=flat)
        assert compare(g, r, f"{name} flat={flat}") == []
        g.close()
