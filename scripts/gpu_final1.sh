# final single-GPU artefacts of the round: tests, default bench line, reference arm, launch list with DRAM bytes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv | head -3
( time timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final_1gpu.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'], d['e2e'], d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_final_ncu.log 2>&1; wc -l gpurun_out/launches_final.csv
python __graft_entry__.py smoke 2>&1 | tail -2
