n=$1
for env in "X=1" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=64 NCCL_MAX_P2P_NCHANNELS=64 NCCL_MAX_NCHANNELS=64" "NCCL_P2P_USE_CUDA_MEMCPY=1"; do
env $env python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err || tail -5 gpurun_out/bench_ab.err; python -c "
import json;d=json.load(open('gpurun_out/bench_ab.json'));s=d['stage_ms'];print('$env', round(d['ms_per_step'],2), 'a2a',round(s['a2a_tuples'],2),'dd_route',round(s['dd_route'],2),'node_gather',round(s['node_gather'],2))"; done
