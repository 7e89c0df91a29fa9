# 8-GPU box: strong-scaling bench line (N = 8, and N = 4) under different overlap / cgm-grid settings
set -x
mkdir -p gpurun_out
run() {  # name n env...
  name=$1; n=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $n --no-e2e --no-cpu --no-parity --steps 8 > gpurun_out/r02_var_$name.json 2> gpurun_out/r02_var_$name.err || tail -5 gpurun_out/r02_var_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_var_$name.json')); print('$name N=$n', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()}, d['checks']['gauss_residual'])
except Exception as e: print('$name failed', e)
PY
}
run base8 8 WM_OVERLAP_SORT=0
run ov8_bps2 8 WM_OVERLAP_SORT=1 WM_CG_BPS=2
run ov8_bps4 8 WM_OVERLAP_SORT=1 WM_CG_BPS=4
run noov8_bps2 8 WM_OVERLAP_SORT=0 WM_CG_BPS=2
run noov8_bps1 8 WM_OVERLAP_SORT=0 WM_CG_BPS=1
run ov4_bps4 4 WM_OVERLAP_SORT=1 WM_CG_BPS=4
