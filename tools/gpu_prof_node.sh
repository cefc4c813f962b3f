ZZB200_CUBIN=$PWD/zigzagboomerang.jl_b200/zzb200_kernels_prof.cubin python tools/quick_bench.py 1000 2.0 loose 0.25 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //"
