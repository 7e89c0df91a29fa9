set -x
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/r02_setup_$name.json 2> gpurun_out/r02_setup_$name.err || tail -5 gpurun_out/r02_setup_$name.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_setup_$name.json')); print('$name', round(d['value']/1e9,3),'G/s', round(d['ms_per_step'],3),'ms', d['config']['particles'], {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()}, d['checks'])
except Exception as e: print('$name failed', e)
PY
}

run shock2 --setup shock --dim 2 --nx 1000 --ny 1024 --ppc 16 --steps 10 --warmup 3

run shock3 --setup shock --dim 3 --nx 1000 --ny 64 --nz 16 --ppc 16 --steps 10 --warmup 3
