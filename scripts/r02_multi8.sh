# 8-GPU box: multi-GPU parity tests (2, 4, 8 ranks) with logs, then the strong-scaling bench line at N = 2, 4, 8 (and weak at 8)
set -x
mkdir -p gpurun_out; rm -f gpurun_out/multigpu_parity.log
nvidia-smi --query-gpu=index,name --format=csv | head -10
( timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | tail -8
cat gpurun_out/multigpu_parity.log
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --no-e2e --no-cpu $BENCH_EXTRA > gpurun_out/r02_${TAG:-base}_strong$n.json 2> gpurun_out/r02_${TAG:-base}_strong$n.err; tail -2 gpurun_out/r02_${TAG:-base}_strong$n.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_${TAG:-base}_strong$n.json')); print('strong N=$n', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks']['parity'], d['clocks']['sm_mhz'])
PY
done
WM_FIELD_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --no-e2e --no-cpu --no-parity 2>&1 >/dev/null | grep "field__fdtd_i stages" | head -3
