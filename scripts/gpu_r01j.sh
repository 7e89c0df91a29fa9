set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity3d.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; python -c "
import json; d=json.load(open('gpurun_out/bench_j.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_j.err
timeout 1200 python -m pytest tests/test_gpu_parity_variants.py -q -m gpu 2>&1 | tail -60
