set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_f.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_f.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gather" -s 1 -c 1 -o gpurun_out/prof_sort_r01f python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_sort.log 2>&1
tail -3 gpurun_out/ncu_sort.log | cut -c1-200
