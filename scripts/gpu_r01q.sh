set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity3d.py -q -m gpu -x 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'])"; tail -3 gpurun_out/bench_q.err
