set -x
mkdir -p gpurun_out
timeout 150 python bench.py --strong-nz 16 --steps 3 --warmup 3 > gpurun_out/r02_t7.json 2> gpurun_out/r02_t7.err; tail -3 gpurun_out/r02_t7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_t7.json')); print(d['value']/1e9, d['ms_per_step'], d['e2e'], d['cpu_baseline'], d['checks']['parity']['pass'], d['gpu_launches'], d['clocks'])
PY
