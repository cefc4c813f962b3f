python tools/quick_bench.py 1000 2.0 loose 0.1 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //'
for mb in 2 3 4; do ZZB200_CUBIN=$PWD/zigzagboomerang.jl_b200/zzb200_kernels_mb$mb.cubin python tools/quick_bench.py 1000 2.0 loose 0.1 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' ; done
