set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_variants.py tests/test_abi.py -q -m gpu -x 2>&1 | tail -25
