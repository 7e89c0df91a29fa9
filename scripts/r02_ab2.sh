# A/B of an alternative build of the library (WM_B200_LIB): parity tests on it, then the weak C2 bench line for both
set -x
mkdir -p gpurun_out
ALT=$PWD/wumingpic_b200/lib/libwuming_b200_alt.so
( WM_B200_LIB=$ALT timeout 900 python -m pytest tests -q -m gpu -x -k "parity3d or occupancy or five_calls or variants" 2>&1 | tail -4 ) 2>&1 | tail -6
for v in alt main; do
  if [ $v = alt ]; then export WM_B200_LIB=$ALT; else unset WM_B200_LIB; fi
  timeout 600 python bench.py --weak --no-e2e --no-cpu > gpurun_out/r02_ab2_$v.json 2> gpurun_out/r02_ab2_$v.err; tail -2 gpurun_out/r02_ab2_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_ab2_$v.json')); print('$v', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks']['parity']['pass'], d['checks']['gauss_residual'])
PY
done
