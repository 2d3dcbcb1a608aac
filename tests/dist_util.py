"""Helpers for the world_size-2 tests: spawn ranks with a 127.0.0.1 rendezvous."""
import os
import socket
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(rank, world, port, fn, args, errfile):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        try:
            fn(rank, world, *args)
        finally:
            dist.destroy_process_group()
    except Exception:
        with open(f"{errfile}.{rank}", "w") as f:
            f.write(traceback.format_exc())
        raise


def run_ranks(fn, world, args=(), tmpdir="/tmp"):
    import torch.multiprocessing as mp
    port = free_port()
    errfile = os.path.join(str(tmpdir), f"dist_err_{port}")
    try:
        mp.spawn(_entry, args=(world, port, fn, args, errfile), nprocs=world, join=True)
    except Exception as e:
        msgs = []
        for r in range(world):
            if os.path.exists(f"{errfile}.{r}"):
                msgs.append(f"--- rank {r} ---\n" + open(f"{errfile}.{r}").read())
        raise AssertionError("a rank failed:\n" + "\n".join(msgs)) from e
