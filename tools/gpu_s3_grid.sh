python -m pytest tests/test_gpu_grid.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_grid.py --deselect tests/test_gpu_golden.py 2>&1 | tail -3
python tools/quick_bench.py 1000 2.0 loose 0.15 2>&1 | grep "trace=False" | cut -c1-220
