# round 2, first GPU call: all GPU tests, the default (strong, 256x256x128) bench line without and with e2e/cpu, the five-call path
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm,memory.total --format=csv | head -3
free -g | head -2; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) 2>&1 | tail -22
timeout 900 python bench.py --no-e2e --no-cpu > gpurun_out/r02_bench_strong1.json 2> gpurun_out/r02_bench_strong1.err; tail -3 gpurun_out/r02_bench_strong1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_strong1.json')); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'], d['clocks'], d['gpu_launches'])
PY
timeout 900 python bench.py --no-e2e --no-cpu --five-calls > gpurun_out/r02_bench_five.json 2> gpurun_out/r02_bench_five.err; tail -3 gpurun_out/r02_bench_five.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_five.json')); print('five calls:', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['checks']['parity'], d['gpu_launches'])
PY
timeout 900 python bench.py --weak --no-e2e --no-cpu > gpurun_out/r02_bench_weak1.json 2> gpurun_out/r02_bench_weak1.err; tail -3 gpurun_out/r02_bench_weak1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_weak1.json')); print('weak:', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'])
PY
( time timeout 1200 python bench.py > gpurun_out/r02_bench_full1.json 2> gpurun_out/r02_bench_full1.err ) 2>&1 | tail -4; tail -3 gpurun_out/r02_bench_full1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_full1.json')); print('full:', d['value']/1e9,'G/s', d['e2e'], d['cpu_baseline'])
PY
