# A/B of an alternative kernel image (ZZB200_CUBIN) against the built one
python tools/quick_bench.py 1000 2.0 loose 0.25 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //" -e "s/'node_evals.*'ns_scan'/'ns_scan'/" -e "s/'dbg0.*//" -e 's/^/base /'
for f in zigzagboomerang.jl_b200/zzb200_kernels_*.cubin; do
ZZB200_CUBIN=$PWD/$f python tools/quick_bench.py 1000 2.0 loose 0.25 2>&1 | grep "trace=False" | sed -e "s/upload.*proposals -> //" -e "s/'node_evals.*'ns_scan'/'ns_scan'/" -e "s/'dbg0.*//" -e "s|^|$f |"
done
