set -x
mkdir -p gpurun_out
./tools/fp64_peak | tee gpurun_out/fp64_peak.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1500 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
