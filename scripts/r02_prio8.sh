set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus 8 --no-e2e --no-cpu --no-parity --steps 10 > gpurun_out/r02_prio_$name.json 2> gpurun_out/r02_prio_$name.err || tail -5 gpurun_out/r02_prio_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_prio_$name.json')); print('$name', round(d['value']/1e9,2),'G/s', round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()})
except Exception as e: print('$name failed', e)
PY
}
run equal WM_STREAM_PRIO=0
run sorthigh WM_STREAM_PRIO=2
run equal_bps3 WM_STREAM_PRIO=0 WM_CG_BPS=3
WM_OVERLAP_SORT=0 WM_SORT_TIMING=1 WM_FIELD_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29950 bench.py --gpus 8 --no-e2e --no-cpu --no-parity --steps 10 2>&1 >/dev/null | grep "wuming_b200\] rank [03]" | cut -c1-400
