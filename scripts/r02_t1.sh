# 1 GPU: all GPU tests, then strong + weak bench lines (no e2e / cpu legs)
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -8
for m in "" "--weak"; do
timeout 900 python bench.py --no-e2e --no-cpu $m $BENCH_EXTRA > gpurun_out/r02_t1.json 2> gpurun_out/r02_t1.err; tail -2 gpurun_out/r02_t1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_t1.json')); print(d['scaling'], d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks']['parity']['pass'], d['checks']['gauss_residual'], d['gpu_launches'])
PY
done
