# one ncu --set full capture of the fused kernel at a reduced z extent (same per-cell load): NAME=xxx bash scripts/gpu_prof.sh
set -x
mkdir -p gpurun_out
NAME=${NAME:-prof}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fused3 -s 2 -c 1 -f -o gpurun_out/$NAME python bench.py --weak --nz 8 --steps 2 --warmup 2 --no-e2e --no-cpu --no-parity > gpurun_out/$NAME.log 2>&1
ls -la gpurun_out/$NAME.ncu-rep
