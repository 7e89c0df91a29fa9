set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 100 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
