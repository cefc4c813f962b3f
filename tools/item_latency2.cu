// Development tool: single-thread latency of ONE timeline item, variants of the draw / logarithm scheduling.
#include <cstdio>
#include <cuda_runtime.h>
#include "../zigzagboomerang.jl_b200/csrc/zz_core.h"
#include "logbf.h"

template <int VARIANT>
__global__ void k(double* out, long long* cyc, double x0, unsigned long long seed)
{
    ZzHood<5> hd;
    hd.n = 5; hd.self = 2;
    for (int m = 0; m < 5; ++m) {
        hd.wb[m] = (m == 2) ? 4.01 : -1.0; hd.wt[m] = 0.0; hd.fl[m] = (m == 2) ? 3u : 7u;
        hd.th[m] = (m & 1) ? 1.0 : -1.0; hd.tf[m] = 0.0; hd.xf[m] = x0 + 0.1 * m;
    }
    double th = 1.0, tf = 0.0, xf = x0, a = 5.0, b = 4.0, told = 0.0, c = 4.5, c100 = c / 100, tau = 0.01;
    uint32_t kc = 1, nflip = 0;
    const int N = 256;
    double La = 0, Lb = 0, Ua = 0;
    if (VARIANT == 3) { Ua = zz_u01(seed, seed + 1, 7, kc); Lb = zz_log_bf(zz_u01(seed, seed + 1, 7, kc + 1)); La = zz_log_bf(Ua); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N; ++it) {
        const double s = tau;
        double L2, u1;
        if (VARIANT == 0) { L2 = zz_log(zz_u01(seed, seed + 1, 7, kc + 1)); u1 = zz_u01(seed, seed + 1, 7, kc); }
        if (VARIANT == 2) { L2 = zz_log_bf(zz_u01(seed, seed + 1, 7, kc + 1)); u1 = zz_u01(seed, seed + 1, 7, kc); }
        double nUa = 0, nLa = 0, nLb = 0;
        if (VARIANT == 3) {
            L2 = Lb; u1 = Ua;
            nUa = zz_u01(seed, seed + 1, 7, kc + 2); nLb = zz_log_bf(zz_u01(seed, seed + 1, 7, kc + 3)); nLa = zz_log_bf(nUa);
        }
        const double xs = xf + th * (s - tf);
        double gt, gx, gp, gm;
        zz_eval_hood<5>(hd, true, s, xs, th, gt, gx, gp, gm);
        kc += 2;
        double gth = gp;
        const double l = zz_pos(gt * th), lb = zz_pos(a + b * (s - told));
        if (u1 * lb < l) { nflip++; xf = xs; tf = s; th = -th; gth = gm; }
        a = c + (gx - 0.0) * th;
        b = c100 + th * gth;
        told = s;
        double dt = zz_poisson_time_L(a, b, L2);
        if (!(dt < 1e300)) dt = 0.01;
        tau = s + dt * 1e-3;
        if (VARIANT == 3) { Ua = nUa; La = nLa; Lb = nLb; }
    }
    long long t1 = clock64();
    cyc[VARIANT] = (t1 - t0) / N;
    out[VARIANT] = tau + nflip + a + b + La;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
    for (int rep = 0; rep < 2; ++rep) { k<0><<<1, 1>>>(out, cyc, 0.3, 99ULL); k<2><<<1, 1>>>(out, cyc, 0.3, 99ULL); k<3><<<1, 1>>>(out, cyc, 0.3, 99ULL); }
    long long h[8]; double o[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 64, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("cycles per item: current %lld, branch-free log %lld, + logs one item ahead %lld (%s)  [%.17g %.17g]\n", h[0], h[2], h[3], cudaGetErrorString(e), o[0], o[2]);
    return 0;
}
