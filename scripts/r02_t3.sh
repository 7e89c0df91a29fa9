set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x -k "setups or restart or occupancy or five_calls" 2>&1 | tail -12 ) 2>&1 | tail -14
