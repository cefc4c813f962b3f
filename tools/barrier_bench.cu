// Micro-benchmark: latency of a software grid barrier on a cooperative launch (development tool).
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long* p)
{ unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p)
{ return *(volatile const unsigned long long*)p; }

template <int MODE>
__global__ void bar_kernel(unsigned long long* ctr, int iters, unsigned long long* sink)
{
    unsigned long long epoch = 0;
    cg::grid_group grid = cg::this_grid();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 3) { grid.sync(); continue; }
        __syncthreads();
        if (threadIdx.x == 0) {
            epoch += gridDim.x;
            if (MODE == 0) {
                __threadfence(); atomicAdd(ctr, 1ULL);
                while (ld_acq(ctr) < epoch) { }
                __threadfence();
            } else if (MODE == 1) {
                __threadfence(); atomicAdd(ctr, 1ULL);
                while (ld_rlx(ctr) < epoch) { }
                __threadfence();
            } else {
                asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(ctr) : "memory");
                while (ld_rlx(ctr) < epoch) { __nanosleep(20); }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = epoch;
}

template <int MODE> float run(int grid, int block, int iters)
{
    unsigned long long *ctr, *sink;
    cudaMalloc(&ctr, 8); cudaMalloc(&sink, 8); cudaMemset(ctr, 0, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    void* args[] = { &ctr, &iters, &sink };
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)bar_kernel<MODE>, dim3(grid), dim3(block), args, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    cudaFree(ctr); cudaFree(sink);
    return ms;
}

int main()
{
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep)
        for (int block : { 256, 1024 })
            for (int grid : { 148, 296 }) {
                if (block == 1024 && grid == 296) continue;
                float a = run<0>(grid, block, iters), b = run<1>(grid, block, iters), c = run<2>(grid, block, iters), d = run<3>(grid, block, iters);
                printf("rep %d grid %d block %d: acquire-poll %.2f us, relaxed-poll %.2f us, red.release+nanosleep %.2f us, cg grid.sync %.2f us per barrier\n",
                       rep, grid, block, 1e3f * a / iters, 1e3f * b / iters, 1e3f * c / iters, 1e3f * d / iters);
            }
    return 0;
}
