nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread|Core" 
python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -3
