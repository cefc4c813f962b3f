python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 1000 2.0 loose 0.2,0.25,0.3 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //" -e "s/'node_evals.*'ns_scan'/'ns_scan'/" -e "s/'dbg0.*//"
python tools/quick_bench.py 1000 2.0 tight 0.25 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //" -e "s/'node_evals.*'ns_scan'/'ns_scan'/" -e "s/'dbg0.*//"
python tools/quick_bench.py 100 20.0 loose 0.25 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //" -e "s/'node_evals.*'ns_scan'/'ns_scan'/" -e "s/'dbg0.*//"
