# GPU check of the logistic target (config 3): parity tests, then timing of 1 / 8 / 64 replicas
set -x
timeout 60 python -m pytest tests/test_gpu_zz_logistic.py -x -q -s 2>&1 | tail -12
timeout 45 python tools/logit_bench.py 10 1 8 64 2>&1 | tail -6
