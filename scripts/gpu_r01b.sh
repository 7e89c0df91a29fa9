set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 2500 gpurun_out/bench_fused.json; tail -5 gpurun_out/bench_fused.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 1 -c 1 -o gpurun_out/prof_fused_r01b python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scatter3 -s 1 -c 1 -o gpurun_out/prof_scatter_r01b python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_scatter.log 2>&1
ls -la gpurun_out/
