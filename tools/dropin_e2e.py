"""e2e_dropin: the whole call sequence the Java driver would make through GpuOverlap:
SFA directory on disk -> preprocess -> buildOverlap -> buildStringGraph, each writing its node-text
directories (00-preprocess, 01-overlap, 01-overlap.chl2, 01-overlap.re) through cb_write_nodes.

    python tools/dropin_e2e.py [--config cfg2] [--genome G] [--out gpurun_out/dropin.json]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import make_workload
    from cloudbrush_b200 import BrushConfig
    from cloudbrush_b200.stages import GpuOverlap
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--genome", type=int, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    sfa, meta = make_workload(args.config, genome=args.genome)
    tmp = tempfile.mkdtemp(prefix="cb_dropin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    reads = os.path.join(tmp, "reads")
    base = os.path.join(tmp, "asm.tmp")
    os.makedirs(reads)
    os.makedirs(base)
    sfa.tofile(os.path.join(reads, "part-00000"))
    res = {"workload": f"{args.config}: {meta['n_reads']} reads of {meta['readlen']} bp, k={meta['k']}", "runs": []}
    for write_text in (True, False, True):
        drv = GpuOverlap(BrushConfig(k=meta["k"], readlen=meta["readlen"]), write_text=write_text)
        t = [time.perf_counter()]
        drv.preprocess(reads, base); t.append(time.perf_counter())
        drv.buildOverlap(reads, base); t.append(time.perf_counter())
        drv.buildStringGraph(reads, base); t.append(time.perf_counter())
        sizes = {d: sum(os.path.getsize(os.path.join(base, d, f)) for f in os.listdir(os.path.join(base, d)))
                 for d in sorted(os.listdir(base))} if write_text else {}
        res["runs"].append({"write_text": write_text, "s_preprocess": t[1] - t[0], "s_buildOverlap": t[2] - t[1],
                            "s_buildStringGraph": t[3] - t[2], "s_total": t[3] - t[0],
                            "reads_per_sec": meta["n_reads"] / (t[3] - t[0]), "text_bytes": sizes,
                            "device_ms": {k: drv.gpu.time_ms(k) for k in ("preprocess", "overlap", "string")}})
        drv.close()
        print(json.dumps(res["runs"][-1]), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
