python tools/sticky_bench.py 316 20 2>&1 | tail -1
python tools/sticky_bench.py 1000 10 2>&1 | tail -1
