set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for G in 8; do
WM_FUSED_G=$G timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g$G.json')); print('G=$G', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_g.err
done
