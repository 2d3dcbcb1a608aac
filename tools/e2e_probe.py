"""Where the steady-state e2e time goes: host wall time of every C-ABI call in the pipelined loop of bench.py,
next to the device time of the stage groups.  Modes (argv[1]): blocking | async | async_nonodes | noexport"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402,F401
import torch  # noqa: E402

from bench import make_workload  # noqa: E402
from cloudbrush_b200 import BrushConfig, BrushGpu  # noqa: E402
from cloudbrush_b200.api import EDGE_DTYPE, NODE_DTYPE  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "async"
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sfa_np, meta = make_workload("cfg2")
host = torch.from_numpy(sfa_np).pin_memory()
g = BrushGpu(BrushConfig(k=21, readlen=36))
g.load_sfa_buffer(host)
g.run()
pins = [torch.empty((g.num_edges(3) + 1024) * EDGE_DTYPE.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
pin_n = torch.empty((g.counter("nodes") + 1024) * NODE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
out_e = [p.numpy().view(EDGE_DTYPE) for p in pins]
out_n = pin_n.numpy().view(NODE_DTYPE)
names = ["prefetch", "preprocess", "wait", "nodes", "overlap", "string", "edges"]
g.prefetch_sfa_buffer(host)
if depth == 1:
    pass
rows = []
torch.cuda.synchronize()
t_all = time.perf_counter()
N = 8
for i in range(N):
    t = [time.perf_counter()]
    if depth == 2:
        g.prefetch_sfa_buffer(host)
    t.append(time.perf_counter())
    g.preprocess()
    t.append(time.perf_counter())
    if mode.startswith("async"):
        g.export_wait()
    t.append(time.perf_counter())
    if depth == 1:
        g.prefetch_sfa_buffer(host)
    if mode == "async":
        g.nodes_begin(out_n)
    t.append(time.perf_counter())
    g.build_overlap()
    t.append(time.perf_counter())
    g.string_graph()
    t.append(time.perf_counter())
    if mode == "blocking":
        g.edges(3, out=out_e[0])
        g.nodes(out=out_n)
    elif mode.startswith("async"):
        g.edges_begin(3, out_e[i & 1])
        if mode == "async_nonodes":
            pass
    t.append(time.perf_counter())
    rows.append([(t[j + 1] - t[j]) * 1e3 for j in range(7)] + [g.time_ms("preprocess"), g.time_ms("overlap"),
                                                                 g.time_ms("string")])
if mode.startswith("async"):
    g.export_wait()
torch.cuda.synchronize()
total = (time.perf_counter() - t_all) * 1e3 / N
print(f"mode {mode} depth {depth}: {total:.2f} ms per batch")
print(" ".join(f"{n:>10}" for n in names + ["dev_pre", "dev_ov", "dev_str"]))
for r in rows:
    print(" ".join(f"{v:10.2f}" for v in r))
g.close()
