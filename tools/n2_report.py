"""Prints the multi-GPU facts of a bench.py JSON line (file argument): step, e2e, exchange timers, put kernel."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "parity", (d.get("parity") or {}).get("ok"))
print(" stage", {k: round(v, 2) for k, v in d["stage_ms"].items()})
print(" counters", {k: v for k, v in d["counters"].items() if k in ("p2p_exchange", "host_collectives")})
print(" kernels", {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels"].items()
                   if k in ("p2p_put", "rs_scatter:partition", "rs_scatter:route", "bucket_join")})
