# one-shot GPU check of the logistic target (config 3) + regression run of the whole GPU suite
set -x
timeout 150 python -m pytest tests/test_gpu_zz_logistic.py -x -q -s --runxfail 2>&1 | tail -30
timeout 100 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zz_logistic.py 2>&1 | tail -5
