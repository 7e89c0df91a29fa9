set -x
mkdir -p gpurun_out; rm -f gpurun_out/multigpu_parity_*.log
( timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "yslab or uneven" 2>&1 | tail -6 ) 2>&1 | tail -8
cat gpurun_out/multigpu_parity_*.log
