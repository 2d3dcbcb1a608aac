"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (cloudbrush_b200/dist.py)."""
import numpy as np
import torch

from dist_util import run_ranks


def _exchange_worker(rank, world):
    from cloudbrush_b200.dist import TorchExchange
    ex = TorchExchange(device=None)
    assert (ex.rank, ex.world, ex.backend) == (rank, world, "gloo")
    # counts: rank r tells rank d the value 10*r + d
    got = ex.exchange_counts([10 * rank + d for d in range(world)])
    assert got.tolist() == [10 * s + rank for s in range(world)]
    # two values per destination
    got = ex.exchange_counts([100 * rank + 2 * d + j for d in range(world) for j in range(2)])
    assert got.tolist() == [100 * s + 2 * rank + j for s in range(world) for j in range(2)]
    # the gather the library builds on this call (dist.cu ex_gather_checked): every rank sends the SAME m words
    # (n values + its failure flag) to every rank and receives every rank's words, here m = 4 with rank 1 "failing"
    mine = [7 * rank + 1, 1000 + rank, 5, 1 if rank == 1 else 0]
    got = ex.exchange_counts(mine * world)
    assert got.tolist() == [v for s in range(world) for v in (7 * s + 1, 1000 + s, 5, 1 if s == 1 else 0)]
    # ... also for more words than a count matrix has (the counters travel this way: 31 words per rank)
    got = ex.exchange_counts([rank * 100 + j for j in range(31)] * world)
    assert got.tolist() == [s * 100 + j for s in range(world) for j in range(31)]
    # alltoallv with ragged (and one empty) blocks: block r->d has (r + 2*d) bytes of value 16*r + d
    send_bytes = [rank + 2 * d for d in range(world)]
    recv_bytes = [s + 2 * rank for s in range(world)]
    send = torch.cat([torch.full((send_bytes[d],), 16 * rank + d, dtype=torch.uint8) for d in range(world)])
    recv = torch.zeros(sum(recv_bytes), dtype=torch.uint8)
    ex.alltoallv(send, send_bytes, recv, recv_bytes)
    want = torch.cat([torch.full((recv_bytes[s],), 16 * s + rank, dtype=torch.uint8) for s in range(world)])
    assert torch.equal(recv, want)
    # allgatherv: rank r contributes 3 + r bytes
    sizes = [3 + r for r in range(world)]
    mine = torch.arange(sizes[rank], dtype=torch.uint8) + 50 * rank
    out = torch.zeros(sum(sizes), dtype=torch.uint8)
    ex.allgatherv(mine, out, sizes)
    want = torch.cat([torch.arange(sizes[r], dtype=torch.uint8) + 50 * r for r in range(world)])
    assert torch.equal(out, want)
    # allreduce sum / max
    assert ex.allreduce([rank + 1, 7], 0).tolist() == [sum(range(1, world + 1)), 7 * world]
    assert ex.allreduce([rank + 1, 7 - rank], 1).tolist() == [world, 7]
    # element-wise max of a byte array (list summaries of the TransitiveReduction final pass)
    b = torch.zeros(8, dtype=torch.uint8)
    b[rank] = 10 + rank
    b[7] = 3 * rank
    ex.allreduce_max_u8(b)
    assert b.tolist() == [10 + r if r < world else 0 for r in range(7)] + [3 * (world - 1)]
    # the C callback structure can be built without a device
    st = ex.as_struct()
    assert st.exchange_counts and st.alltoallv and st.allgatherv and st.allreduce_u64 and st.allreduce_max_u8_dev


def test_torch_exchange_world2_gloo(tmp_path):
    run_ranks(_exchange_worker, 2, tmpdir=tmp_path)


def test_shards_are_consecutive_line_ranges():
    from cloudbrush_b200.dist import shard_lines
    sfa = b"".join(b"%d\tACGT\n" % i for i in range(11))
    parts = [shard_lines(sfa, r, 4) for r in range(4)]
    assert b"".join(parts) == sfa
    assert [p.count(b"\n") for p in parts] == [3, 3, 3, 2]
    assert shard_lines(sfa[:-1], 0, 1) == sfa  # a missing last terminator is normalised


def test_synthetic_shards_concatenate_to_the_single_gpu_read_set():
    from cloudbrush_b200 import synth
    whole, m = synth.make_config("cfg3_50k", scale_genome=4000)
    for world in (2, 4, 8):
        parts = [synth.make_config("cfg3_50k", scale_genome=4000, rank=r, world=world)[0] for r in range(world)]
        assert np.array_equal(np.concatenate(parts), whole)


def test_owner_of_kmer_group_is_a_pure_function_of_the_key():
    """Partitioning rule of the tuple all-to-all (stage_overlap.cu): owner = bits [K, K+lg W) of the
    canonical key.  Every rank must agree, and a k-mer and its reverse complement share an owner."""
    K, W = 21, 8
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 4, size=(1000, K))
    def pack(c):
        v = np.zeros(len(c), dtype=np.uint64)
        for j in range(K):
            v = (v << np.uint64(2)) | c[:, j].astype(np.uint64)
        return v
    fwd = pack(codes)
    rc = pack(3 - codes[:, ::-1])
    canon = np.minimum(fwd, rc)
    owner = (canon >> np.uint64(K)) & np.uint64(W - 1)
    owner_rc = (np.minimum(rc, fwd) >> np.uint64(K)) & np.uint64(W - 1)
    assert np.array_equal(owner, owner_rc)
    assert len(set(owner.tolist())) == W  # all ranks get work on random keys
