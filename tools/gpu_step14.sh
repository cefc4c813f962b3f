python tools/quick_bench.py 1000 2.0 loose 0.02,0.03,0.04 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'rebases.*//"
python tools/quick_bench.py 1000 2.0 tight 0.02,0.03,0.05 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'rebases.*//"
python tools/quick_bench.py 100 20.0 loose 0.03,0.1 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'rebases.*//"
