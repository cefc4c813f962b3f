timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py 1000 2.0 loose 0.05,0.1,0.2 2>&1 | grep "trace=False"
for mb in 2 3; do ZZB200_CUBIN=$PWD/zigzagboomerang.jl_b200/zzb200_kernels_mb$mb.cubin python tools/quick_bench.py 1000 2.0 loose 0.05,0.1,0.2 2>&1 | grep "trace=False"; done
python tools/quick_bench.py 1000 2.0 tight 0.05 2>&1 | grep "trace=False"
python tools/quick_bench.py 100 20.0 loose 0.4 2>&1 | grep "trace=False"
python tools/quick_bench.py 8 200.0 loose 0.4 2>&1 | grep "trace=False"
