timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py 1000 2.0 loose 0.15 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'rebases.*//"
python tools/quick_bench.py 1000 2.0 tight 0.15 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'rebases.*//"
