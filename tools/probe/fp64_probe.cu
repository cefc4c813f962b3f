// Development probe: latency / issue cost of the FP64 operations the event-loop kernels are made of, one warp on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void probe(double* out, long long* cyc, double seed, int nact)
{
    const bool on = (int)threadIdx.x < nact;
    double a = seed + threadIdx.x, b = seed * 0.5, c = seed * 0.25, d = seed * 0.125;
    long long t0, t1;
    // 0: dependent DADD chain
    t0 = clock64();
    if (on) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + b;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 1: four independent DADD chains
    double a1 = a, a2 = a + 1, a3 = a + 2, a4 = a + 3;
    t0 = clock64();
    if (on) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) { a1 = a1 + b; a2 = a2 + c; a3 = a3 + d; a4 = a4 + b; }
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    a = a1 + a2 + a3 + a4;
    // 2: dependent DMUL
    t0 = clock64();
    if (on) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 3: dependent DFMA
    t0 = clock64();
    if (on) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 4: dependent division
    t0 = clock64();
    if (on) {
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) a = c / (a + b);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) * 8;
    // 5: shuffle + add chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + __shfl_sync(0xffffffffu, b, i & 31);
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 6: dependent FADD chain (fp32) for comparison
    float f = (float)seed, g = (float)b;
    t0 = clock64();
    if (on) {
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = f + g;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 7: sqrt
    t0 = clock64();
    if (on) {
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) a = sqrt(a + b);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0) * 8;
    out[threadIdx.x] = a + f;
}
int main()
{
    double* out; long long* cyc; long long h[8];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8);
    const char* names[8] = { "DADD dependent", "DADD 4 chains (per iteration of 4)", "DMUL dependent", "DFMA dependent", "DDIV dependent (+1 DADD)", "SHFL64 + DADD chain", "FADD dependent", "DSQRT dependent (+1 DADD)" };
    for (int nact : { 32, 1 }) {
        probe<<<1, 32>>>(out, cyc, 1.0000001, nact);
        probe<<<1, 32>>>(out, cyc, 1.0000001, nact);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        printf("active lanes %d\n", nact);
        for (int k = 0; k < 8; ++k) printf("  %-40s %.1f cycles/op\n", names[k], (double)h[k] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
