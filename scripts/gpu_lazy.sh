set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
for mode in lazy eager; do
[ $mode = eager ] && export WM_NO_LAZY_SORT=1 || unset WM_NO_LAZY_SORT
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$mode.json 2> gpurun_out/bench_lazy.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$mode.json')); print('$mode', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; tail -2 gpurun_out/bench_lazy.err
done
