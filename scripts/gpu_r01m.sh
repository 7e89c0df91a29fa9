set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity3d.py tests/test_gpu_parity_variants.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_m.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 2 -c 1 -o gpurun_out/prof_fused_m python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_m.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/
