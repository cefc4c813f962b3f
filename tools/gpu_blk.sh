python tools/quick_bench.py 1000 2.0 loose 0.15 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'dbg0.*//" -e 's/^/B256 /'
for b in 320 384 512; do ZZB200_CUBIN=$PWD/zigzagboomerang.jl_b200/zzb200_kernels_b$b.cubin python tools/quick_bench.py 1000 2.0 loose 0.15 2>&1 | grep "trace=False" | sed -e 's/upload.*ms; //' -e "s/'dbg0.*//" -e "s/^/B$b /"; done
