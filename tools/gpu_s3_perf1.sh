python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 1000 2.0 loose 0.15,0.12,0.2 2>&1 | grep "trace=False" | cut -c1-700
python tools/quick_bench.py 1000 2.0 tight 0.15 2>&1 | grep "trace=False" | cut -c1-300
