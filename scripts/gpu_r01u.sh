set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_u$N.json 2> gpurun_out/bench_u$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_u$N.json')); print('N=$N', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; tail -3 gpurun_out/bench_u$N.err
done
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "4" 2>&1 | tail -5
