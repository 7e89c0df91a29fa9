set -x
mkdir -p gpurun_out
( timeout 22 python -m pytest tests/test_zzz_gpu_ref_golden.py -x -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r02_t8_pytest.log
timeout 14 python bench.py --strong-nz 16 --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r02_t8.json 2> gpurun_out/r02_t8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_t8.json')); print(d['value']/1e9, d['ms_per_step'], d['clocks'])
PY
