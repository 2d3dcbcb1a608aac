"""Diagnostic: host wall time and pool statistics per stage over repeated resident runs."""
import sys, time
sys.path.insert(0, ".")
import torch
from cloudbrush_b200 import BrushConfig, BrushGpu, synth

genome = int(sys.argv[1]) if len(sys.argv) > 1 else 4_600_000
buf, m = synth.make_config("cfg2", scale_genome=genome)
host = torch.from_numpy(buf).pin_memory()
dev = host.cuda()
g = BrushGpu(BrushConfig(k=m["k"], readlen=m["readlen"]))
for it in range(5):
    g.borrow_sfa_device(dev.data_ptr(), dev.numel())
    row = []
    for st in ("preprocess", "build_overlap", "string_graph"):
        torch.cuda.synchronize()
        t = time.perf_counter()
        getattr(g, st)()
        torch.cuda.synchronize()
        row.append(f"{st}={1e3*(time.perf_counter()-t):.1f}ms")
    row.append(f"reserved={g.counter('pool_reserved_bytes')/1e9:.2f}GB used={g.counter('pool_used_bytes')/1e9:.2f}GB")
    row.append(f"torch_alloc={torch.cuda.memory_allocated()/1e9:.2f}GB free={torch.cuda.mem_get_info()[0]/1e9:.1f}GB")
    print(it, " ".join(row), flush=True)
    print("   ", {t: round(g.time_ms(t), 2) for t in ["parse", "dedupe", "keygen", "sort_tuples", "join_verify", "cut", "tr"]}, flush=True)
