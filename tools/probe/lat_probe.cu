// Development probe: dependent-load latency (pointer chase) for the load flavours the kernels use, one thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define STEPS 4096
template <int MODE> __global__ void chase(const unsigned int* __restrict__ a, unsigned int start, long long* cyc, unsigned int* sink)
{
    unsigned int p = start;
    for (int i = 0; i < 64; ++i) p = MODE == 0 ? __ldg(a + p) : (MODE == 1 ? __ldcg(a + p) : __ldca(a + p));
    const long long t0 = clock64();
    for (int i = 0; i < STEPS; ++i) p = MODE == 0 ? __ldg(a + p) : (MODE == 1 ? __ldcg(a + p) : __ldca(a + p));
    const long long t1 = clock64();
    *cyc = t1 - t0; *sink = p;
}
__global__ void u01chain(unsigned long long* out, long long* cyc)
{
    unsigned long long z = out[0];
    const long long t0 = clock64();
    for (int i = 0; i < 1024; ++i) {
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
    }
    const long long t1 = clock64();
    *cyc = t1 - t0; out[1] = z;
}
int main()
{
    long long* cyc; unsigned int* sink; cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    for (size_t bytes : { (size_t)16 << 10, (size_t)2 << 20, (size_t)32 << 20, (size_t)1 << 30 }) {
        const size_t n = bytes / 4, stride = 64;   // one hop per 256 bytes, random cycle over the slots
        const size_t slots = n / stride;
        std::vector<unsigned int> perm(slots), h(n, 0);
        for (size_t i = 0; i < slots; ++i) perm[i] = (unsigned int)i;
        unsigned long long s = 12345;
        for (size_t i = slots - 1; i > 0; --i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; size_t j = (s >> 33) % (i + 1); std::swap(perm[i], perm[j]); }
        for (size_t i = 0; i < slots; ++i) h[(size_t)perm[i] * stride] = (unsigned int)((size_t)perm[(i + 1) % slots] * stride);
        unsigned int* d; cudaMalloc(&d, bytes); cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
        long long c[3];
        chase<0><<<1, 1>>>(d, 0, cyc, sink); chase<0><<<1, 1>>>(d, 0, cyc, sink); cudaMemcpy(&c[0], cyc, 8, cudaMemcpyDeviceToHost);
        chase<1><<<1, 1>>>(d, 0, cyc, sink); chase<1><<<1, 1>>>(d, 0, cyc, sink); cudaMemcpy(&c[1], cyc, 8, cudaMemcpyDeviceToHost);
        chase<2><<<1, 1>>>(d, 0, cyc, sink); chase<2><<<1, 1>>>(d, 0, cyc, sink); cudaMemcpy(&c[2], cyc, 8, cudaMemcpyDeviceToHost);
        printf("%8zu KB: ldg %.0f  ldcg %.0f  ldca %.0f cycles per dependent load (%zu slots%s)\n", bytes >> 10, (double)c[0] / STEPS, (double)c[1] / STEPS,
               (double)c[2] / STEPS, slots, slots < STEPS ? ", revisited" : "");
        cudaFree(d);
    }
    unsigned long long* o; cudaMalloc(&o, 16); cudaMemset(o, 1, 16);
    u01chain<<<1, 1>>>(o, cyc); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mix64 round: %.1f cycles\n", (double)c / 1024);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
