# multi-GPU parity + bench, peer-memory cgm vs NCCL-sequenced cgm:  N=2 bash scripts/gpu_multi_ab.sh
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -15
for mode in ${MODES:-peer nccl}; do
[ $mode = nccl ] && export WM_CG_NCCL=1 || unset WM_CG_NCCL
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_m${N}_$mode.json 2> gpurun_out/bench_m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_m${N}_$mode.json')); print('N=$N $mode', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; grep -i "wuming\|error" gpurun_out/bench_m.err | head -5
done
