set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r02_bench_n1.json
python tools/n2_report.py gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-other-configs --no-cpu-baseline --no-parity > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
