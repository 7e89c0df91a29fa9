set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv | head -3
( time timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) 2>&1
timeout 900 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -2 gpurun_out/bench_w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_w.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'], d['e2e'], d['cpu_baseline'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_w_ref.json 2>> gpurun_out/bench_w.err; cat gpurun_out/bench_w_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_w_ncu.log 2>&1; wc -l gpurun_out/launches_w.csv
