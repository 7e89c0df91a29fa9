set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x 2>&1 | tail -25
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | tail -4 | cut -c1-1500
