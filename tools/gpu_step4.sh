python tools/quick_bench.py 1000 2.0 loose 0.1 2>&1 | grep "trace=False"
python tools/quick_bench.py 100 20.0 loose 0.4 2>&1 | grep "trace=False"
python tools/quick_bench.py 8 200.0 loose 0.4 2>&1 | grep "trace=False"
