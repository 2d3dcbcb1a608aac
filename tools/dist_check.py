"""Multi-GPU parity check (run under torchrun on a box with >= world GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [--transport nccl|torch]

Every rank runs the sharded pipeline on its shard; rank 0 also runs the whole input on one GPU
and compares the union of the ranks' checkpoints, the node table and the counters with it.
"""
import argparse
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COUNTERS = ["reads_good", "nodes", "contained", "tuples", "edges_overlap", "edges_chl2", "edges_re", "trans_edge",
            "edge_removal_1", "edge_removal_2"]


def main():
    import torch
    import torch.distributed as dist
    from cloudbrush_b200 import BrushConfig, BrushGpu, synth
    from cloudbrush_b200.dist import TorchExchange, shard_lines
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default="nccl")
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    inputs = [("ec10k", gzip.open(os.path.join(ROOT, "tests/golden/ec10k_reads.sfa.gz")).read(), 21, 36)]
    for name, scale in (("cfg2_50k", 200_000), ("cfg3_50k", 50_000), ("cfg5_50k", 30_000)):
        buf, m = synth.make_config(name, scale_genome=scale)
        inputs.append((name, bytes(buf), m["k"], m["readlen"]))
    ok = True
    for name, sfa, k, L in inputs:
        g = BrushGpu(BrushConfig(k=k, readlen=L, device=lr, rank=rank, world=world))
        if args.transport == "nccl":
            g.init_nccl()
        else:
            g.set_exchange(TorchExchange(device=torch.device("cuda", lr)))
        g.load_sfa_buffer(shard_lines(sfa, rank, world))
        g.run()
        mine = dict(a=g.edge_rows(1), c=g.edge_rows(2), b=g.edge_rows(3), nodes=g.nodes(),
                    counters=[g.counter(x) for x in COUNTERS])
        g.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            s = BrushGpu(BrushConfig(k=k, readlen=L, device=lr))
            s.load_sfa_buffer(sfa)
            s.run()
            bad = []
            for key, ck in (("a", 1), ("c", 2), ("b", 3)):
                got = np.concatenate([q[key] for q in parts])
                got = got[np.lexsort((got[:, 3], got[:, 2], got[:, 1], got[:, 0]))]
                if not np.array_equal(got, s.edge_rows(ck)):
                    bad.append(f"checkpoint {key}: {len(got)} vs {len(s.edge_rows(ck))}")
            if not np.array_equal(np.concatenate([q["nodes"] for q in parts]), s.nodes()):
                bad.append("node table")
            for q in parts:
                if q["counters"] != [s.counter(x) for x in COUNTERS]:
                    bad.append(f"counters {q['counters']} vs {[s.counter(x) for x in COUNTERS]}")
            print(f"{name}: world={world} transport={args.transport} nodes={s.counter('nodes')} "
                  f"A={s.counter('edges_overlap')} B={s.counter('edges_re')} -> " + ("PARITY OK" if not bad else "FAIL " + "; ".join(bad)),
                  flush=True)
            ok = ok and not bad
            s.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
