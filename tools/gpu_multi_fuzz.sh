ZZB_MULTI_FUZZ=100 ZZB_MULTI_FUZZ_SEED=${1:-9} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py 2>&1 | tail -12
