timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 1000 2.0 loose 0.1,0.15,0.2 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'dbg0.*//" -e "s/1382882 switches 7754030 proposals -> //"
python tools/quick_bench.py 8 200.0 loose 0.4 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'dbg0.*//"
