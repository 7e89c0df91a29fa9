set -x
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -30
