# quick A/B of the 3-D fused kernel: parity tests, then the weak C2 bench line (536.9 M particles) old vs new
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x -k "parity3d or occupancy or five_calls" 2>&1 | tail -5 ) 2>&1 | tail -8
for v in new old; do
  if [ $v = old ]; then export WM_FUSED3_OLD=1; else unset WM_FUSED3_OLD; fi
  timeout 600 python bench.py --weak --no-e2e --no-cpu > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.err; tail -2 gpurun_out/r02_ab_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_ab_$v.json')); print('$v', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks']['parity']['pass'], d['checks']['gauss_residual'], d['clocks'])
PY
done
