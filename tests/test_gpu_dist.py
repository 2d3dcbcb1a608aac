"""-m gpu: the sharded (multi-rank) pipeline gives exactly the single-GPU checkpoints.

Two ranks share cuda:0 here (gloo transport staged through the host, because NCCL refuses two ranks
on one device); on a multi-GPU box bench.py runs the same code over NCCL."""
import gzip
import os

import numpy as np
import pytest

from dist_util import ROOT, run_ranks

pytestmark = pytest.mark.gpu


def _sharded_worker(rank, world, sfa_path, k, L, outdir):
    import torch
    from cloudbrush_b200 import BrushConfig, BrushGpu
    from cloudbrush_b200.dist import TorchExchange, shard_lines
    sfa = open(sfa_path, "rb").read()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = BrushGpu(BrushConfig(k=k, readlen=L, device=0, rank=rank, world=world))
    g.set_exchange(TorchExchange(device=dev))
    g.load_sfa_buffer(shard_lines(sfa, rank, world))
    g.run()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), a=g.edge_rows(1), c=g.edge_rows(2), b=g.edge_rows(3),
             nodes=g.nodes(), counters=np.array([g.counter(x) for x in COUNTERS]))
    for ck in (0, 1, 2, 3):  # collective: every rank writes part-000<rank> of the checkpoint's directory
        g.write_nodes(ck, os.path.join(outdir, f"text{ck}"))
    g.close()


COUNTERS = ["reads_good", "nodes", "contained", "tuples", "edges_overlap", "edges_chl2", "edges_re", "trans_edge",
            "edge_removal_1", "edge_removal_2"]


def _check(sfa, k, L, tmp_path, world=2):
    from cloudbrush_b200 import BrushConfig, BrushGpu
    p = tmp_path / "in.sfa"
    p.write_bytes(sfa)
    run_ranks(_sharded_worker, world, args=(str(p), k, L, str(tmp_path)), tmpdir=tmp_path)
    g = BrushGpu(BrushConfig(k=k, readlen=L))
    g.load_sfa_buffer(sfa)
    g.run()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for key, ck in (("a", 1), ("c", 2), ("b", 3)):
        got = np.concatenate([q[key] for q in parts])
        got = got[np.lexsort((got[:, 3], got[:, 2], got[:, 1], got[:, 0]))]
        assert np.array_equal(got, g.edge_rows(ck)), f"checkpoint {key} differs between 1 and {world} ranks"
    # every rank exports the nodes built from its shard of the lines: in rank order, the node table
    assert np.array_equal(np.concatenate([q["nodes"] for q in parts]), g.nodes())
    for q in parts:  # the counters are global
        assert q["counters"].tolist() == [g.counter(x) for x in COUNTERS]
    # node text of a sharded run: one part file per rank; their union is the single-context directory
    from test_gpu_cases import _canon_text
    for ck in (0, 1, 2, 3):
        d1 = tmp_path / f"single{ck}"
        g.write_nodes(ck, d1)
        names = sorted(os.listdir(tmp_path / f"text{ck}"))
        assert names == [f"part-{r:05d}" for r in range(world)]
        sharded = "".join(open(tmp_path / f"text{ck}" / nm).read() for nm in names)
        assert _canon_text(sharded) == _canon_text(open(d1 / "part-00000").read()), f"node text of checkpoint {ck}"
    g.close()


def test_two_ranks_ec10k(tmp_path, ec10k_sfa):
    _check(ec10k_sfa, 21, 36, tmp_path)


def test_two_ranks_synthetic_both_strands(tmp_path):
    from cloudbrush_b200 import synth
    buf, m = synth.make_config("cfg3_50k", scale_genome=20_000)
    _check(bytes(buf), m["k"], m["readlen"], tmp_path)


def test_four_ranks_micro_cases(tmp_path):
    from cases import all_cases
    for name in ("tiling_both_strands", "branch", "rc_duplicates"):
        buf, k, L = all_cases()[name]
        d = tmp_path / name
        d.mkdir()
        _check(buf, k, L, d, world=4)


def test_native_nccl_two_gpus(tmp_path):
    """Native NCCL transport (cb_init_nccl) on a box with two GPUs: tools/dist_check.py under torchrun."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the same check runs through `gpurun --gpus 2 tools/dist_check.py`)")
    from dist_util import free_port
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
                        os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("PARITY OK") == 4
