python -m pytest tests -q -m gpu 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 2 -c 1 -o gpurun_out/prof_fused_r01a python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter3 -s 2 -c 1 -o gpurun_out/prof_scatter_r01a python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_scatter.log 2>&1
ls -la gpurun_out/*.ncu-rep
