CB_OS_EC=1 python tools/gpu_check.py 2>&1 | grep -c "PARITY OK"
for ec in 1 0; do CB_OS_EC=$ec python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err || tail -3 gpurun_out/bench_ab.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));k=d['kernels'];print('ec $ec',round(d['ms_per_step'],3),'tuples',round(k['rs_scatter:tuples']['ms_per_step'],3),k['rs_scatter:tuples']['launches_per_step'],'dedupe',round(k['rs_scatter:dedupe']['ms_per_step'],3),'frac',round(d['roofline']['frac'],3), d['counters']['edges_re'])"; done
