nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_worker.py 2>&1 | grep -v "Warning\|warn" | tail -25
