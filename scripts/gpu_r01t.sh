set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 2 -c 1 -o gpurun_out/prof_fused_t python bench.py --nz 8 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
