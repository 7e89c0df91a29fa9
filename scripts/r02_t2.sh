# 2 GPUs: slab parity tests at 2 ranks (all variants) and the strong bench line at N = 2
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "not 4 and not 8" 2>&1 | tail -6 ) 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --no-e2e --no-cpu $BENCH_EXTRA > gpurun_out/r02_t2.json 2> gpurun_out/r02_t2.err; tail -2 gpurun_out/r02_t2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_t2.json')); print(d['scaling'], d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks']['parity']['pass'], d['checks']['gauss_residual'], d['gpu_launches'])
PY
