# A/B of kernel variants: VARIANTS="main old ..." bash scripts/gpu_ab.sh ; main = the shipped library
set -x
mkdir -p gpurun_out
[ -n "$SKIPTESTS" ] || timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
for v in $VARIANTS; do
[ "$v" = main ] && L=$PWD/wumingpic_b200/lib/libwuming_b200.so || L=$PWD/wumingpic_b200/lib/libwuming_b200_$v.so
WM_B200_LIB=$L timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ab_$v.json')); print('$v', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; tail -3 gpurun_out/bench_ab.err
done
