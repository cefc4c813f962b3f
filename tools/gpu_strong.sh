# EXPERIMENTAL: build the image with the strong-bound sparse sticky kernel (zz_strong.h) and run its parity tests on a B200.
# Not run in round 1 (no GPU budget left when it was written); first thing to do for SURVEY 8(f) rank 2 in round 2.
set -x
make -s -C zigzagboomerang.jl_b200/csrc strong
E=$PWD/zigzagboomerang.jl_b200/experimental
ZZB200_EXPERIMENTAL=1 ZZB200_LIB=$E/libzzb200.so ZZB200_CUBIN=$E/zzb200_kernels.cubin timeout 150 python -m pytest tests/test_gpu_zzz_strong.py -x -q -s 2>&1 | tail -15
