set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity3d.py tests/test_gpu_parity_variants.py -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; python -c "
import json; d=json.load(open('gpurun_out/bench_l.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'], d['gpu_launches'])"; tail -5 gpurun_out/bench_l.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_l.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu && ./tools/fp64_peak | tee gpurun_out/fp64_peak.json
ls -la gpurun_out/
