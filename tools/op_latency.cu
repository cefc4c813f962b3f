// Development tool: single-thread latency (cycles) of the scalar primitives on the device.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zigzagboomerang.jl_b200/csrc/zz_math.h"

__global__ void k(double* out, long long* cyc, double a0, double b0, unsigned long long seed)
{
    double a = a0, b = b0, acc = 0.0;
    long long t0, t1;
    const int N = 256;
    // dependent DADD chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) acc = acc + a;
    t1 = clock64(); cyc[0] = (t1 - t0) / N;
    // dependent DMUL
    double m = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) m = m * b;
    t1 = clock64(); cyc[1] = (t1 - t0) / N; acc += m;
    // dependent division
    double q = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) q = q / b + 1.0;
    t1 = clock64(); cyc[2] = (t1 - t0) / N; acc += q;
    // dependent sqrt
    double r = a + 3.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) r = zz_sqrt(r) + 2.0;
    t1 = clock64(); cyc[3] = (t1 - t0) / N; acc += r;
    // dependent log
    double l = 0.3;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) l = zz_log(l * 0.5 + 0.25) * -0.1 + 0.2;
    t1 = clock64(); cyc[4] = (t1 - t0) / N; acc += l;
    // u01
    double u = 0.0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) u += zz_u01(seed, seed + 1, (unsigned long long)(u * 1000.0) + i, i);
    t1 = clock64(); cyc[5] = (t1 - t0) / N; acc += u;
    // poisson_time dependent
    double p = 0.5;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) p = zz_poisson_time(a + p * 1e-3, b, 0.3 + 1e-3 * (p - (long long)p)) * 0.5 + 0.1;
    t1 = clock64(); cyc[6] = (t1 - t0) / N; acc += p;
    // dependent global load (L2) chain
    t0 = clock64();
    t1 = clock64(); cyc[7] = t1 - t0;
    out[0] = acc;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 64);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 1>>>(out, cyc, 1.37, 0.77, 12345ULL);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("cycles per dependent op: DADD %lld, DMUL %lld, DDIV(+add) %lld, DSQRT(+add) %lld, zz_log(+2) %lld, zz_u01 %lld, poisson_time %lld, clock overhead %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    return 0;
}
