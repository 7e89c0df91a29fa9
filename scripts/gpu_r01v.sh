set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v2.json')); print('N=2', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; tail -3 gpurun_out/bench_v2.err
