python tools/quick_bench.py 1000 2.0 loose 0.1,0.15 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'dbg0.*//" -e "s/1382882 switches 7754030 proposals -> //"
