set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -c 3000 gpurun_out/bench_i.json; tail -5 gpurun_out/bench_i.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_i_ref.json 2>> gpurun_out/bench_i.err; tail -c 800 gpurun_out/bench_i_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_i.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 2 -c 1 -o gpurun_out/prof_fused_i python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 2 -c 1 -o gpurun_out/prof_gather_i python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_gather.log 2>&1
ls -la gpurun_out/
