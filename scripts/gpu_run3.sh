mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err; tail -c 3000 gpurun_out/bench_v1.json; tail -5 gpurun_out/bench_v1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_v1.json 2>> gpurun_out/bench_v1.err; cat gpurun_out/bench_ref_v1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_v1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
