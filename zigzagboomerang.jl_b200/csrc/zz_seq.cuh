// zz_seq.cuh -- sequential chains: ONE WARP runs the exact event loop of one connected component of the dependency graph.
//
// The windowed relaxation of zz_kernels.cu needs many coordinates whose timelines are independent inside a window.  A problem
// whose dependency graph is (nearly) complete has none: config 3 of BASELINE.json (scripts/logistic.jl: 442 coordinates, two
// dense regressors that read every other coordinate) made every flip re-evaluate the dense columns, one thread at a time.  For
// such problems -- and for many small independent chains side by side (replicas) -- the event loop of the reference is run
// as it is written, spdmp_inner! (src/sfact.jl:73-145), by one warp per component:
//   * the priority queue (src/priorityqueue.jl) is the array of proposal times in SHARED memory, `peek` is a warp arg-min
//     (lane-strided scan + five shuffle steps, ties to the smaller coordinate: the strict order of the `ctr` contract);
//   * the whole state of the chain (anchors x, t, theta; queue times; bounds a, b, t_old, c; draw counters) lives in shared
//     memory for the duration of the launch: 80 bytes per coordinate, nothing but the design/precision entries comes from L2;
//   * the subsampled logistic gradient (scripts/logistic.jl:78-107) is evaluated by k lanes side by side -- lane r draws row
//     r, gathers its coordinates, evaluates the four sigmoid terms -- and summed in the reference's order through shuffles;
//   * a rejected proposal recomputes its own bound cooperatively (one column entry per lane, ordered sum through shuffles),
//     an accepted flip reschedules G1[i] with one neighbour per lane (fact_samplers.jl:50-54, sfact.jl:131-135).
// Arithmetic is the contract's (`ctr|lazy`): positions are flip-anchored, every sum runs in storage order,
// every draw comes from the drawing coordinate's own counter stream -- so the output is bit-identical to the oracle and to the
// windowed kernels.
//
// Stopping rule (sfact.jl:199: the loop ends after the first ACCEPTED flip at or after T): with several components the end
// time is the earliest such flip over all of them.  Phase 0 processes every item before T; phase 1 lets every chain look
// ahead (no writes) for its first accepted flip and min-reduces the time; phase 2 processes every item up to that time.
//
// Trace: a chain reserves ZZ_SEQ_RES records at a time (one atomic per 32 events); unused records become markers (i = 0)
// that the ordering step drops.  When the buffer is full the chain saves its state and the host drains and relaunches.
#ifndef ZZ_SEQ_CUH
#define ZZ_SEQ_CUH

#define ZZ_SEQ_RES 32u

struct ZzSeqSh {
    double *xf, *tf, *th, *tau, *a, *b, *told, *c;
    uint32_t* kc;
    int32_t *a0, *al;
    int32_t lo, nc;
};

__device__ __forceinline__ double zz_seq_pos(const ZzSeqSh& S, int32_t k, double s)
{
    return S.xf[k] + S.th[k] * (s - S.tf[k]);
}

// Ordered sums over column jg of a CSC matrix: sx = sum val * x_k(s), st = sum val * theta_k, storage order.
// Cooperative: one entry per lane, accumulated in order through shuffles (every lane ends with the sums).
__device__ __forceinline__ void zz_seq_col_coop(const ZzSeqSh& S, const int32_t* __restrict__ row, const double* __restrict__ val,
                                                int32_t e0, int32_t e1, double s, int lane, double& sx, double& st)
{
    double ax = 0.0, at = 0.0;
    for (int32_t base = e0; base < e1; base += 32) {
        const int32_t e = base + lane;
        double px = 0.0, pt = 0.0;
        if (e < e1) {
            const int32_t k = __ldg(row + e) - S.lo;
            const double w = __ldg(val + e);
            const double thk = S.th[k];
            px = w * (S.xf[k] + thk * (s - S.tf[k]));
            pt = w * thk;
        }
        const int cnt = min(32, e1 - base);
        for (int q = 0; q < cnt; ++q) {
            ax += __shfl_sync(0xffffffffu, px, q);
            at += __shfl_sync(0xffffffffu, pt, q);
        }
    }
    sx = ax; st = at;
}

// The same sums by one lane on its own.
__device__ __forceinline__ void zz_seq_col_serial(const ZzSeqSh& S, const int32_t* __restrict__ row, const double* __restrict__ val,
                                                  int32_t e0, int32_t e1, double s, double& sx, double& st)
{
    double ax = 0.0, at = 0.0;
    for (int32_t e = e0; e < e1; ++e) {
        const int32_t k = __ldg(row + e) - S.lo;
        const double w = __ldg(val + e);
        const double thk = S.th[k];
        ax += w * (S.xf[k] + thk * (s - S.tf[k]));
        at += w * thk;
    }
    sx = ax; st = at;
}

// gamma0 x_j - fdot_moving(...) at time s (scripts/logistic.jl:78-95,107): draws kc .. kc + L.k - 1 of coordinate jg's stream,
// lane r evaluates sampled row r; the 4 k terms are added in the reference's order.
__device__ __forceinline__ double zz_seq_logit_grad(const ZzLogit& L, const ZzView& v, const ZzSeqSh& S, int32_t li, int32_t jg,
                                                    double s, double xown, uint32_t kc, int lane)
{
    const int32_t e0 = S.a0[li], l = S.al[li];
    const double lk = (double)l / (double)L.k;
    double sacc = 0.0;
    for (int32_t r0 = 0; r0 < L.k; r0 += 32) {
        const int32_t r = r0 + lane;
        double t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0;
        if (r < L.k) {
            const double ur = zz_u01(v.seed0, v.seed1, (uint64_t)jg, (uint64_t)(kc + (uint32_t)r));
            int32_t i = (int32_t)(ur * (double)l);                  // uniform index into nzrange(A, j) (:83,86)
            if (i >= l) i = l - 1;
            const int32_t row = __ldg(L.arow + e0 + i);
            const double w = lk * __ldg(L.aval + e0 + i);
            const int32_t q0 = __ldg(L.rp + row), q1 = __ldg(L.rp + row + 1);
            const double yr = __ldg(L.y + row), nyr = __ldg(L.ny + row), u0 = __ldg(L.u0 + row);
            double u = 0.0;                                         // idot_moving!(At, row, ...), src/common.jl:33-42
            for (int32_t q = q0; q < q1; ++q) {
                const int32_t m = __ldg(L.rcol + q) - S.lo;
                u += __ldg(L.rval + q) * zz_seq_pos(S, m, s);
            }
            t1 = w * yr * zz_sigmoidn(u);                           // :87-88
            t2 = w * nyr * zz_nsigmoid(u);
            t3 = w * yr * zz_sigmoidn(u0);                          // :90-91 (control variate at the mode)
            t4 = w * nyr * zz_nsigmoid(u0);
        }
        const int cnt = min(32, L.k - r0);
        for (int q = 0; q < cnt; ++q) {
            sacc += __shfl_sync(0xffffffffu, t1, q);
            sacc += __shfl_sync(0xffffffffu, t2, q);
            sacc -= __shfl_sync(0xffffffffu, t3, q);
            sacc -= __shfl_sync(0xffffffffu, t4, q);
        }
    }
    return L.gamma0 * xown - sacc;                                  // :107
}

template <bool LOGIT>
__device__ __forceinline__ void zz_seq_body(const ZzParams& P, const ZzSeq& Q)
{
    extern __shared__ unsigned int zz_dyn[];
    const int lane = (int)threadIdx.x;
    ZzDevCtl* C = P.ctl;
    const int comp = (int)blockIdx.x;
    if (comp >= Q.ncomp) return;
    ZzSeqSh S;
    S.lo = __ldg(Q.comp + comp); S.nc = __ldg(Q.comp + comp + 1) - S.lo;
    {
        const size_t n = (size_t)Q.ncmax;   // (even: every array stays 8-byte aligned)
        double* base = reinterpret_cast<double*>(zz_dyn);
        S.xf = base; S.tf = base + n; S.th = base + 2 * n; S.tau = base + 3 * n;
        S.a = base + 4 * n; S.b = base + 5 * n; S.told = base + 6 * n; S.c = base + 7 * n;
        S.kc = reinterpret_cast<uint32_t*>(base + 8 * n);
        S.a0 = reinterpret_cast<int32_t*>(S.kc + n);
        S.al = S.a0 + n;
    }
    const int32_t lo = S.lo, nc = S.nc;
    for (int32_t q = lane; q < nc; q += 32) {
        double th, tf, xf; uint32_t h0, h1;
        zz_ld_kin(P.v.kin + lo + q, th, tf, xf, h0, h1);
        const ZzPriv pr = zz_ld_priv(P.v.priv + lo + q);
        S.xf[q] = xf; S.tf[q] = tf; S.th[q] = th; S.tau[q] = __ldcg(P.v.tau + lo + q);
        S.a[q] = pr.a; S.b[q] = pr.b; S.told[q] = pr.told; S.c[q] = pr.c;
        S.kc[q] = __ldcg(P.v.kctr + lo + q);
        if (LOGIT) { const int32_t e0 = __ldg(P.lg.acp + lo + q); S.a0[q] = e0; S.al[q] = __ldg(P.lg.acp + lo + q + 1) - e0; }
    }
    __syncwarp();

    const int phase = Q.phase;
    const double tend = (phase == 2) ? zz_unkey(__ldcg(&C->smin_key[0])) : 0.0;
    const bool rec = P.record_trace && phase != 1;
    const uint64_t seed0 = P.v.seed0, seed1 = P.v.seed1;
    unsigned long long nprop = 0, nflip = 0;
    unsigned long long tr_pos = 0, tr_end = 0;

#ifdef ZZ_SEQ_PROF
    long long pc[6] = { 0, 0, 0, 0, 0, 0 }; long long pt0;
#define ZZ_SP_TIC() pt0 = clock64()
#define ZZ_SP_TOC(k) pc[k] += clock64() - pt0
#else
#define ZZ_SP_TIC()
#define ZZ_SP_TOC(k)
#endif
    for (;;) {
        ZZ_SP_TIC();
        // ---- peek (sfact.jl:77): earliest queue time, ties to the smaller coordinate
        double bt = ZZ_INF; int bi = 0x7fffffff;
        for (int32_t q = lane; q < nc; q += 32) {
            const double t = S.tau[q];
            if (t < bt) { bt = t; bi = q; }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            const double ot = __shfl_xor_sync(0xffffffffu, bt, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ot < bt || (ot == bt && oi < bi)) { bt = ot; bi = oi; }
        }
        if (bi == 0x7fffffff) break;   // nothing will ever happen in this chain
        const double tp = bt;
        const int32_t li = bi, jg = lo + bi;
        if (phase == 0) { if (!(tp < P.T)) break; }
        else if (phase == 2) { if (!(tp <= tend)) break; }

        // ---- room for one trace record
        if (rec && tr_pos == tr_end) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&C->trace_len, (unsigned long long)ZZ_SEQ_RES);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + ZZ_SEQ_RES > P.trace_cap) {   // full: markers into what is left, the host drains and relaunches
                for (unsigned long long p = base + (unsigned long long)lane; p < P.trace_cap; p += 32ULL) {
                    double2* e = reinterpret_cast<double2*>(P.trace + p);
                    e[0] = make_double2(0.0, __longlong_as_double(0LL));
                    e[1] = make_double2(0.0, 0.0);
                }
                if (lane == 0) C->need_drain = 1u;
                break;
            }
            tr_pos = base; tr_end = base + ZZ_SEQ_RES;
        }

        ZZ_SP_TOC(0);
        ZZ_SP_TIC();
        // ---- the proposal of coordinate i at tp (sfact.jl:118-121)
        const int32_t be0 = __ldg(Q.bcp + jg), be1 = __ldg(Q.bcp + jg + 1);
        const double gmu_i = __ldg(P.g.gmu + jg);
        const double th_i = S.th[li], tf_i = S.tf[li], xf_i = S.xf[li];
        const double xi = xf_i + th_i * (tp - tf_i);
        double c = S.c[li];
        uint32_t kc = S.kc[li];
        double gt, sx = 0.0, sth = 0.0;
        bool have_s = false;
        if (LOGIT) {
            gt = zz_seq_logit_grad(P.lg, P.v, S, li, jg, tp, xi, kc, lane);
            kc += (uint32_t)P.lg.k;
        } else if (Q.tcp) {
            double d0, d1;
            zz_seq_col_coop(S, Q.trow, Q.tval, __ldg(Q.tcp + jg), __ldg(Q.tcp + jg + 1), tp, lane, d0, d1);
            gt = P.g.h ? d0 - __ldg(P.g.h + jg) : d0;
        } else {
            zz_seq_col_coop(S, Q.brow, Q.bval, be0, be1, tp, lane, sx, sth);
            have_s = true;
            gt = sx;
        }
        ZZ_SP_TOC(1);
        ZZ_SP_TIC();
        const double l = zz_pos(gt * th_i);                                  // fact_samplers.jl:28-30
        const double lb = zz_pos(S.a[li] + S.b[li] * (tp - S.told[li]));     // sfact.jl:70
        const double u = zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc++));
        nprop++;
        ZZ_SP_TOC(2);
        ZZ_SP_TIC();
        if (u * lb < l) {                                                    // sfact.jl:121
            if (phase == 1) {   // look-ahead only: report the time, leave no trace
                if (lane == 0) atomicMin(&C->smin_key[0], zz_key(tp));
                break;
            }
            if (lane == 0) atomicAdd(P.acc + jg, 1u);                        // :122
            if (l >= lb) {                                                   // :123-128
                if (P.v.adapt) c *= P.v.factor;
                else {
                    if (lane == 0 && atomicExch(&C->viol, 1u) == 0u) { C->viol_i = jg + 1; C->viol_t = tp; C->viol_l = l; C->viol_lb = lb; }
                    break;
                }
            }
            nflip++;
            const int32_t tid = P.trace_map ? __ldg(P.trace_map + jg) : jg + 1;   // id in the trace (0: filtered out, src/trace.jl:275-290)
            __syncwarp();
            if (lane == 0) {
                // moment sums of the segment that ends here (trace.jl:194), the event (sfact.jl:50-52,143), the flip (:130)
                atomicAdd(P.s1 + jg, (xf_i + xi) * (tp - tf_i));
                atomicAdd(P.s2 + jg, (tp - tf_i) * (xf_i * xf_i + xf_i * xi + xi * xi));
                if (rec && tid != 0) {
                    double2* e = reinterpret_cast<double2*>(P.trace + tr_pos);
                    e[0] = make_double2(tp, __longlong_as_double((long long)tid));
                    e[1] = make_double2(xi, -th_i);
                }
                S.xf[li] = xi; S.tf[li] = tp; S.th[li] = -th_i; S.c[li] = c; S.kc[li] = kc;
            }
            if (rec && tid != 0) tr_pos++;
            __syncwarp();
            // reschedule G1[i] = rows of column i of Z.Gamma (i among them), one neighbour per lane (:131-135)
            for (int32_t base = be0; base < be1; base += 32) {
                const int32_t e = base + lane;
                if (e < be1) {
                    const int32_t jj = __ldg(Q.brow + e), lj = jj - lo;
                    const int32_t f0 = __ldg(Q.bcp + jj), f1 = __ldg(Q.bcp + jj + 1);
                    const double gmu_j = __ldg(P.g.gmu + jj);
                    double sxj, stj;
                    zz_seq_col_serial(S, Q.brow, Q.bval, f0, f1, tp, sxj, stj);
                    const double cj = S.c[lj], thj = S.th[lj];
                    const double aj = cj + (sxj - gmu_j) * thj;               // fact_samplers.jl:51-52
                    const double bj = cj / 100 + thj * stj;
                    uint32_t kj = S.kc[lj];
                    const double tj = tp + zz_poisson_time(aj, bj, zz_u01(seed0, seed1, (uint64_t)jj, (uint64_t)(kj++)));
                    S.a[lj] = aj; S.b[lj] = bj; S.told[lj] = tp; S.tau[lj] = tj; S.kc[lj] = kj;
                }
            }
            __syncwarp();
            ZZ_SP_TOC(3);
        } else {                                                             // :137-140
            if (!have_s) zz_seq_col_coop(S, Q.brow, Q.bval, be0, be1, tp, lane, sx, sth);
            const double a = c + (sx - gmu_i) * th_i;
            const double b = c / 100 + th_i * sth;
            const double tau = tp + zz_poisson_time(a, b, zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc++)));
            __syncwarp();
            if (lane == 0) { S.a[li] = a; S.b[li] = b; S.told[li] = tp; S.tau[li] = tau; S.kc[li] = kc; }
            __syncwarp();
            ZZ_SP_TOC(4);
            if (be1 - be0 > 32) { ZZ_SP_TOC(5); }
        }
    }
#ifdef ZZ_SEQ_PROF
    if (lane == 0 && phase == 0) for (int q = 0; q < 6; ++q) atomicAdd(&C->dbg[q], (unsigned long long)pc[q]);
#endif

    if (phase == 1) return;
    __syncwarp();
    for (int32_t q = lane; q < nc; q += 32) {
        double2* kq = reinterpret_cast<double2*>(P.v.kin + lo + q);
        kq[0] = make_double2(S.th[q], S.tf[q]);
        reinterpret_cast<double*>(P.v.kin + lo + q)[2] = S.xf[q];
        double2* pq = reinterpret_cast<double2*>(P.v.priv + lo + q);
        pq[0] = make_double2(S.a[q], S.b[q]);
        pq[1] = make_double2(S.told[q], S.c[q]);
        P.v.tau[lo + q] = S.tau[q];
        P.v.kctr[lo + q] = S.kc[q];
    }
    for (unsigned long long p = tr_pos + (unsigned long long)lane; p < tr_end; p += 32ULL) {   // unused reservations
        double2* e = reinterpret_cast<double2*>(P.trace + p);
        e[0] = make_double2(0.0, __longlong_as_double(0LL));
        e[1] = make_double2(0.0, 0.0);
    }
    if (lane == 0) {
        if (nprop) atomicAdd(&C->num, nprop);
        if (nflip) atomicAdd(&C->nacc, nflip);
    }
}

extern "C" __global__ void __launch_bounds__(32) zz_seq_kernel(const __grid_constant__ ZzParams P, const __grid_constant__ ZzSeq Q)
{
    zz_seq_body<false>(P, Q);
}
extern "C" __global__ void __launch_bounds__(32) zz_seq_kernel_logit(const __grid_constant__ ZzParams P, const __grid_constant__ ZzSeq Q)
{
    zz_seq_body<true>(P, Q);
}

#endif  // ZZ_SEQ_CUH
