"""Debug driver: run oracle + CUDA path on several inputs and print every difference."""
import gzip
import sys
import time
import traceback

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from parity_util import compare, run_oracle  # noqa: E402
from cloudbrush_b200 import BrushConfig, BrushGpu, synth  # noqa: E402


def one(label, sfa, k, L):
    print(f"=== {label}: {len(sfa)} bytes k={k} L={L}", flush=True)
    t = time.time()
    r = run_oracle(sfa, k, L)
    print(f"oracle {time.time()-t:.2f}s nodes={r.counter('nodes')} A={len(r.edges(1))} B={len(r.edges(3))}", flush=True)
    g = BrushGpu(BrushConfig(k=k, readlen=L))
    try:
        g.load_sfa_buffer(sfa)
        for st in ("preprocess", "build_overlap", "string_graph"):
            t = time.time()
            getattr(g, st)()
            print(f"gpu {st}: {time.time()-t:.3f}s", flush=True)
        for c in ["nodes", "tuples", "candidates_verified", "edges_overlap", "edges_chl2", "edges_re", "edge_removal_1",
                  "edge_removal_2", "trans_edge", "slow_lists", "hkmer"]:
            print(" ", c, g.counter(c))
        for tname in ["parse", "dedupe", "keygen", "sort_tuples", "join_verify", "cut", "tr"]:
            print(f"  t[{tname}] = {g.time_ms(tname):.3f} ms")
        d = compare(g, r, label)
        print("PARITY OK" if not d else "PARITY FAIL:\n  " + "\n  ".join(d), flush=True)
    except Exception:
        traceback.print_exc()
    finally:
        g.close()


if __name__ == "__main__":
    tiny = b"r1\tACGTTGCATGCCGATTAGCTAGGCTAACGTAGCTAGCTTGACC\nr2\tGCCGATTAGCTAGGCTAACGTAGCTAGCTTGACCTTTGACCAGT\n"
    one("tiny", tiny, 21, 44)
    ec = gzip.open("tests/golden/ec10k_reads.sfa.gz").read()
    one("ec10k", ec, 21, 36)
    for name in ("cfg2_50k", "cfg3_50k", "cfg5_50k"):
        buf, m = synth.make_config(name, scale_genome=20000)
        one(name, bytes(buf), m["k"], m["readlen"])
