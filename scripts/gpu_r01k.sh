set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity3d.py tests/test_gpu_parity_variants.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; python -c "
import json; d=json.load(open('gpurun_out/bench_k.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_k.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 2 -c 1 -o gpurun_out/prof_fused_k python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 2 -c 1 -o gpurun_out/prof_gather_k python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out/
