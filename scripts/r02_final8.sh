# 8-GPU box, final: slab parity at 4 and 8 ranks (fused path, logs kept), strong-scaling bench line at N = 2, 4, 8, weak at 8,
# and the shock / reconnection set-ups at 8 GPUs
set -x
mkdir -p gpurun_out; rm -f gpurun_out/multigpu_parity_*.log
( timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "test_slab_parity and fused and not 2-" 2>&1 | tail -4 ) 2>&1 | tail -6
cat gpurun_out/multigpu_parity_*.log
tr() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $n "$@"; }
for n in 8 4; do
  tr $n --no-e2e --no-cpu > gpurun_out/r02_final_strong$n.json 2> gpurun_out/r02_final_strong$n.err || tail -3 gpurun_out/r02_final_strong$n.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_final_strong$n.json')); print('strong N=$n', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()}, d['checks']['parity']['pass'], d['checks']['gauss_residual'], d['clocks'])
PY
done
st() { name=$1; shift; tr 8 "$@" > gpurun_out/r02_final_$name.json 2> gpurun_out/r02_final_$name.err || tail -4 gpurun_out/r02_final_$name.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_final_$name.json')); print('$name', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],3),'ms', d['config']['particles'], {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()}, d['checks'])
except Exception as e: print('$name failed', e)
PY
}
st rec2_8gpu --setup reconnection --dim 2 --nx 321 --ny 2048 --ppc 50 --steps 10 --warmup 3
st shock2_8gpu --setup shock --dim 2 --nx 1000 --ny 2048 --ppc 32 --steps 10 --warmup 3
st rec3_8gpu --setup reconnection --dim 3 --nx 321 --ny 128 --nz 32 --ppc 50 --steps 10 --warmup 3
