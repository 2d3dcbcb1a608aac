#!/bin/bash
# quick GPU iteration: parity on small inputs + one bench line summary.  usage: tools/quick.sh <tag> [bench args]
tag=$1; shift
python tools/gpu_check.py > gpurun_out/check_$tag.log 2>&1; grep -c "PARITY OK" gpurun_out/check_$tag.log; grep -B3 -A12 "PARITY FAIL\|Error\|error" gpurun_out/check_$tag.log | head -40
python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "roof", round(d["roofline"]["frac"],3))
print({k:round(v,2) for k,v in d["stage_ms"].items()})
print({k:round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
PY
