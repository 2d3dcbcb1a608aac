"""-m gpu: the drop-in itself -- the Python twin of java/Brush/GpuOverlap.java driven the way
BrushAssembler.preprocess / buildOverlap / buildStringGraph drive their stages (directories of part files in,
directories of node text out, RunningJob counters), on the reference's bundled input."""
import os

import numpy as np
import pytest

from parity_util import run_oracle
from test_gpu_cases import _canon_text

pytestmark = pytest.mark.gpu


def _reads_dir(tmp_path, sfa):
    """a Hadoop-style input directory: two part files (the second without a final newline), _SUCCESS, a hidden file"""
    d = tmp_path / "Ec10k"
    d.mkdir()
    lines = sfa.split(b"\n")
    if lines[-1] == b"":
        lines.pop()
    half = len(lines) // 2
    (d / "part-00000").write_bytes(b"\n".join(lines[:half]) + b"\n")
    (d / "part-00001").write_bytes(b"\n".join(lines[half:]))
    (d / "_SUCCESS").write_bytes(b"")
    (d / ".part-00000.crc").write_bytes(b"garbage\tACGT\n")
    return d


def test_gpuoverlap_directories_match_the_oracle_node_text(tmp_path, ec10k_sfa):
    from cloudbrush_b200 import BrushConfig
    from cloudbrush_b200.stages import GpuOverlap
    from oracle import oracle as O
    reads = _reads_dir(tmp_path, ec10k_sfa)
    base = tmp_path / "Ec10k_Brush.tmp"
    base.mkdir()
    r = run_oracle(ec10k_sfa, 21, 36)
    drv = GpuOverlap(BrushConfig(k=21, readlen=36))
    job = drv.preprocess(str(reads), str(base))
    assert job.isSuccessful() and job.getJobID().startswith("job_gpu_")
    c = job.getCounters()
    for name in ("reads_good", "reads_goodbp", "nodes", "contained", "redundant", "nodecount"):
        assert c.findCounter("Brush", name).getValue() == max(r.counter(name), 0), name
    assert os.listdir(base / "stopwords") == ["part-00000"] and os.path.getsize(base / "stopwords" / "part-00000") == 0
    job = drv.buildOverlap(str(reads), str(base))
    assert job.getCounters().findCounter("Brush", "hkmer").getValue() == 0
    job = drv.buildStringGraph(str(reads), str(base))
    c = job.getCounters()
    assert c.findCounter("Brush", "trans_edge").getValue() == r.counter("trans_edge") == 102813
    assert c.findCounter("Brush", "edge_removal").getValue() == max(r.counter("edge_removal_2"), 0)
    for name, ock in (("00-preprocess", O.CK_PREPROCESS), ("01-overlap", O.CK_A), ("01-overlap.chl2", O.CK_CHL2),
                      ("01-overlap.re", O.CK_B)):
        assert sorted(os.listdir(base / name)) == ["part-00000"]
        assert _canon_text(open(base / name / "part-00000").read()) == _canon_text(r.node_text(ock)), name
    drv.close()


def test_load_paths_agree(tmp_path, ec10k_sfa):
    """cb_load_sfa (file and directory), cb_load_sfa_buffer, cb_load_sfa_device and cb_borrow_sfa_device give the
    same graph; an unaligned borrowed device pointer is refused (the parser reads aligned 16-byte vectors)."""
    import torch
    from cloudbrush_b200 import BrushConfig, BrushGpu, BrushGpuError
    sfa = b"\n".join(ec10k_sfa.split(b"\n")[:3000]) + b"\n"
    reads = _reads_dir(tmp_path, sfa)
    one = tmp_path / "one.sfa"
    one.write_bytes(sfa)
    g = BrushGpu(BrushConfig(k=21, readlen=36))
    g.load_sfa_buffer(sfa)
    g.run()
    want = [g.edge_rows(ck) for ck in (1, 3)]
    dev = torch.frombuffer(bytearray(b"\0" * 16 + sfa), dtype=torch.uint8).cuda()
    dev13 = torch.frombuffer(bytearray(b"\0" * 13 + sfa), dtype=torch.uint8).cuda()
    loaders = [lambda: g.load_sfa(one), lambda: g.load_sfa(reads),
               lambda: g.load_sfa_device(dev.data_ptr() + 16, len(sfa)),
               lambda: g.load_sfa_device(dev13.data_ptr() + 13, len(sfa)),  # unaligned source: it is copied, fine
               lambda: g.borrow_sfa_device(dev.data_ptr() + 16, len(sfa))]
    for load in loaders:
        load()
        g.run()
        for w, ck in zip(want, (1, 3)):
            assert np.array_equal(w, g.edge_rows(ck))
    with pytest.raises(BrushGpuError) as e:
        g.borrow_sfa_device(dev.data_ptr() + 17, 100)
    assert e.value.code == -1
    g.close()


def test_fasta_of_the_nodes(tmp_path):
    """cb_write_fasta: Graph2Fasta.java:49-75 on the node set (id, len, cov, sequence in lines of 60)."""
    from cases import all_cases
    buf, k, L = all_cases()["rc_duplicates"]   # ids 9 / 10 / 7 / 8: "10" survives its class with cov 3
    from parity_util import run_gpu
    g = run_gpu(buf, k, L)
    g.write_fasta(tmp_path / "fa")
    txt = open(tmp_path / "fa" / "part-00000").read().split("\n")
    reads = dict(l.split("\t") for l in buf.decode().strip().split("\n"))
    assert txt[0] == ">10\tlen=40\tcov=3.0" and txt[1] == reads["10"]
    assert txt[2] == ">7\tlen=40\tcov=1.0" and txt[3] == reads["7"] and txt[4] == ""
    g.close()
    buf, meta = __import__("cloudbrush_b200").synth.make_config("cfg3_50k", scale_genome=3000)
    g = run_gpu(bytes(buf), meta["k"], meta["readlen"])
    g.write_fasta(tmp_path / "fb")
    lines = open(tmp_path / "fb" / "part-00000").read().split("\n")
    assert lines[0].startswith(">") and len(lines[1]) == 60 and len(lines[2]) == 40   # 100 bp in lines of 60
    assert sum(l.startswith(">") for l in lines) == g.counter("nodes")
    g.close()
