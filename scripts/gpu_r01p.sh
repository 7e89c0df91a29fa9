set -x
mkdir -p gpurun_out
for v in "" _mb3 _mb2; do
WM_B200_LIB=$PWD/wumingpic_b200/lib/libwuming_b200$v.so timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_p$v.json 2> gpurun_out/bench_p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p$v.json')); print('$v', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'])"; tail -3 gpurun_out/bench_p.err
done
timeout 600 python bench.py --dim 2 --nx 4096 --ny 4096 --ppc 16 --steps 3 --warmup 2 > gpurun_out/bench_p2d.json 2> gpurun_out/bench_p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p2d.json')); print('2D', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'], d['config']['workload'])"; tail -3 gpurun_out/bench_p.err
timeout 600 python bench.py --dim 2 --nx 4096 --ny 4096 --ppc 16 --steps 3 --warmup 2 --unfused > gpurun_out/bench_p2du.json 2> gpurun_out/bench_p.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p2du.json')); print('2D unfused', d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['checks'])"; tail -3 gpurun_out/bench_p.err
