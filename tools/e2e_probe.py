import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import make_workload
from cloudbrush_b200 import BrushConfig, BrushGpu
from cloudbrush_b200 import _lib as L_
from cloudbrush_b200.api import EDGE_DTYPE, NODE_DTYPE
sfa_np, meta = make_workload("cfg2")
host = torch.from_numpy(sfa_np).pin_memory()
g = BrushGpu(BrushConfig(k=21, readlen=36))
g.load_sfa_buffer(host); g.run()
pin_e = torch.empty((g.num_edges(3) + 1024) * EDGE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
pin_n = torch.empty((g.counter("nodes") + 1024) * NODE_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
out_e = pin_e.numpy().view(EDGE_DTYPE); out_n = pin_n.numpy().view(NODE_DTYPE)
g.prefetch_sfa_buffer(host)
for i in range(5):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    g.preprocess(); t.append(time.perf_counter())
    g.prefetch_sfa_buffer(host); t.append(time.perf_counter())
    g.build_overlap(); t.append(time.perf_counter())
    g.string_graph(); t.append(time.perf_counter())
    g.edges(3, out=out_e); g.nodes(out=out_n); t.append(time.perf_counter())
    print("pre %.2f prefetch %.2f overlap %.2f string %.2f export %.2f total %.2f | dev pre %.2f ov %.2f st %.2f" % tuple(
        [(t[j+1]-t[j])*1e3 for j in range(5)] + [(t[5]-t[0])*1e3, g.time_ms("preprocess"), g.time_ms("overlap"), g.time_ms("string")]))
