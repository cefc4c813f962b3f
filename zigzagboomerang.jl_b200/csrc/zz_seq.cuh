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
#define ZZ_SEQ_LONG 64    // neighbours with longer columns are rescheduled by the whole warp, one after the other

struct ZzSeqSh {
    double *xf, *tf, *th, *tau, *a, *b, *told, *c;
    uint32_t* kc;
    int32_t *a0, *al;
    int32_t* og;           // original id of every coordinate of the chain (draw stream, trace id, per-coordinate arrays)
    double *g, *cx, *ct;   // scratch: terms of the logistic gradient (64), products of one column of Z.Gamma (colmax each)
    int32_t lo, nc;
};

__device__ __forceinline__ double zz_seq_pos(const ZzSeqSh& S, int32_t k, double s)
{
    return S.xf[k] + S.th[k] * (s - S.tf[k]);
}

// In-order sums of a[0..n) and b[0..n) (shared memory; readable up to n + 8): the loads of the next four terms are in flight
// while the two dependent chains consume the current four.
__device__ __forceinline__ void zz_seq_sum2(const double* a, const double* b, int32_t n, double& sa, double& sb)
{
    double x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3], y0 = b[0], y1 = b[1], y2 = b[2], y3 = b[3];
    double ax = 0.0, ay = 0.0;
    for (int32_t q = 0; q < n; q += 4) {
        const double nx0 = a[q + 4], nx1 = a[q + 5], nx2 = a[q + 6], nx3 = a[q + 7];
        const double ny0 = b[q + 4], ny1 = b[q + 5], ny2 = b[q + 6], ny3 = b[q + 7];
        ax += x0; ay += y0;
        if (q + 1 < n) { ax += x1; ay += y1; }
        if (q + 2 < n) { ax += x2; ay += y2; }
        if (q + 3 < n) { ax += x3; ay += y3; }
        x0 = nx0; x1 = nx1; x2 = nx2; x3 = nx3; y0 = ny0; y1 = ny1; y2 = ny2; y3 = ny3;
    }
    sa = ax; sb = ay;
}

// Ordered sums over entries [e0, e1) of a CSC column: sx = sum val * x_k(s), st = sum val * theta_k, storage order.
// Cooperative: the lanes form the products side by side and park them in shared memory, then every lane adds them up in
// order (the same addresses in every lane: broadcast reads, no shuffles).
__device__ __forceinline__ void zz_seq_col_coop(const ZzSeqSh& S, const int32_t* __restrict__ row, const double* __restrict__ val,
                                                int32_t e0, int32_t e1, double s, int lane, double& sx, double& st)
{
    const int32_t n = e1 - e0;
    __syncwarp();   // (earlier readers of the scratch arrays are done)
#pragma unroll 4
    for (int32_t q = lane; q < n; q += 32) {
        const int32_t k = __ldg(row + e0 + q) - S.lo;
        const double w = __ldg(val + e0 + q);
        const double thk = S.th[k];
        S.cx[q] = w * (S.xf[k] + thk * (s - S.tf[k]));
        S.ct[q] = w * thk;
    }
    __syncwarp();
    sx = 0.0; st = 0.0;
    if (n > 0) zz_seq_sum2(S.cx, S.ct, n, sx, st);
}

// The same sums by one lane on its own (four entries in flight at a time).
__device__ __forceinline__ void zz_seq_col_serial(const ZzSeqSh& S, const int32_t* __restrict__ row, const double* __restrict__ val,
                                                  int32_t e0, int32_t e1, double s, double& sx, double& st)
{
    double ax = 0.0, at = 0.0;
    for (int32_t e = e0; e < e1; e += 4) {
        int32_t k[4]; double w[4];
#pragma unroll
        for (int z = 0; z < 4; ++z) {
            const bool ok = e + z < e1;
            k[z] = ok ? __ldg(row + e + z) - S.lo : 0;
            w[z] = ok ? __ldg(val + e + z) : 0.0;
        }
#pragma unroll
        for (int z = 0; z < 4; ++z)
            if (e + z < e1) {
                const double thk = S.th[k[z]];
                ax += w[z] * (S.xf[k[z]] + thk * (s - S.tf[k[z]]));
                at += w[z] * thk;
            }
    }
    sx = ax; st = at;
}

// Lexicographic arg-min of (time, index) over the warp: three REDUX steps on the order-preserving image of the time.
__device__ __forceinline__ void zz_seq_argmin(double bt, int bi, double& tp, int& li)
{
    const unsigned long long key = zz_key(bt);
    const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
    const unsigned int mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    const bool win = (hi == mh) && (lo == ml);
    li = (int)__reduce_min_sync(0xffffffffu, win ? (unsigned int)bi : 0xffffffffu);
    tp = zz_unkey(((unsigned long long)mh << 32) | (unsigned long long)ml);
}

#define ZZ_SEQ_MAXW 8   // warps per chain

// Verdicts of one step, one per warp (shared memory)
struct ZzSeqStep {
    double newtau[ZZ_SEQ_MAXW];   // rejected proposal: the coordinate's next proposal time
    int idx[ZZ_SEQ_MAXW];         // local index of the item the warp took (-1: none)
    int verdict[ZZ_SEQ_MAXW];     // 0 nothing / invalidates what follows, 1 rejected, 2 accepted
    int stop;                     // the chain ends after this step (trace buffer full, bound violation, look-ahead hit)
    int acc_li;                   // local index of the coordinate that flipped in this step (-1: none) and the time of the flip
    double acc_tp;
};

// SPECULATIVE THINNING.  87 % of the proposals of config 3 are rejected, and a rejection changes nothing another coordinate
// reads: only the proposing coordinate's own bound, counter and next time.  So the W warps of a chain take the W EARLIEST queue
// entries side by side; warp w's result stands iff every earlier entry of the step was rejected AND re-queued later than warp
// w's entry (then the sequential loop would have found exactly the state warp w read).  The first accepted flip ends the
// step: its warp reschedules the neighbourhood alone, the later warps' work is discarded and redone in the next step.
// Expected entries per step with rejection probability r: 1 + r + r^2 + ... -- 1.9 for two warps, 3.3 for four at r = 0.87.
// W = 1 is the plain loop.  Nothing of this changes a single bit of the output: every entry that counts is evaluated on
// exactly the state the reference's loop would have at that point.
template <bool LOGIT>
__device__ __forceinline__ void zz_seq_body(const ZzParams& P, const ZzSeq& Q)
{
    extern __shared__ unsigned int zz_dyn[];
    __shared__ ZzSeqStep V;
    const int lane = (int)(threadIdx.x & 31u), wid = (int)(threadIdx.x >> 5), nw = (int)(blockDim.x >> 5);
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
        // original ids: only when the host renumbered the coordinates (Q.orig set); otherwise id = lo + local index
        S.og = Q.orig ? S.al + n : nullptr;
        double* scr = reinterpret_cast<double*>(S.al + n + (Q.orig ? n : (n & 1)));   // (3 n or 4 n words + padding: 8-byte aligned again)
        scr += (size_t)wid * (72 + 2 * ((size_t)Q.colmax + 8));                       // every warp has its own scratch
        S.g = scr; S.cx = S.g + 72; S.ct = S.cx + Q.colmax + 8;   // (the pipelined sums read up to 8 entries past the end)
    }
    const int32_t lo = S.lo, nc = S.nc;
    for (int32_t q = (int32_t)threadIdx.x; q < nc; q += (int32_t)blockDim.x) {
        const int32_t o = Q.orig ? __ldg(Q.orig + lo + q) : lo + q;
        double th, tf, xf; uint32_t h0, h1;
        zz_ld_kin(P.v.kin + o, th, tf, xf, h0, h1);
        const ZzPriv pr = zz_ld_priv(P.v.priv + o);
        S.xf[q] = xf; S.tf[q] = tf; S.th[q] = th; S.tau[q] = __ldcg(P.v.tau + o);
        S.a[q] = pr.a; S.b[q] = pr.b; S.told[q] = pr.told; S.c[q] = pr.c;
        S.kc[q] = __ldcg(P.v.kctr + o);
        if (S.og) S.og[q] = o;
        if (LOGIT) { const int32_t e0 = __ldg(P.lg.acp + o); S.a0[q] = e0; S.al[q] = __ldg(P.lg.acp + o + 1) - e0; }
    }
    if (threadIdx.x == 0) V.stop = 0;

    const int phase = Q.phase;
    const double tend = (phase == 2) ? zz_unkey(__ldcg(&C->smin_key[0])) : 0.0;
    const bool rec = P.record_trace && phase != 1;
    const uint64_t seed0 = P.v.seed0, seed1 = P.v.seed1;
    unsigned long long nprop = 0, nflip = 0;
    unsigned long long tr_pos = 0, tr_end = 0;
    const int half = lane >> 4, r16 = lane & 15;   // logistic gradient: lanes r and r + 16 share sampled row r

#ifdef ZZ_SEQ_PROF
    long long pc[6] = { 0, 0, 0, 0, 0, 0 }; long long pt0;
#define ZZ_SP_TIC() pt0 = clock64()
#define ZZ_SP_TOC(k) pc[k] += clock64() - pt0
#else
#define ZZ_SP_TIC()
#define ZZ_SP_TOC(k)
#endif
    for (;;) {
        __syncthreads();   // the state of the chain is consistent: every warp sees what the previous step applied
        if (V.stop) break;
        ZZ_SP_TIC();
        // ---- peek (sfact.jl:77): the nw earliest queue times, ties to the smaller coordinate.  Every warp computes all of them
        // (same shared memory, same result); a lane keeps its two earliest entries, so a third entry of one lane among the nw
        // earliest merely shortens the step.
        double b1 = ZZ_INF, b2 = ZZ_INF; int i1 = 0x7fffffff, i2 = 0x7fffffff;
        for (int32_t q = lane; q < nc; q += 32) {
            const double t = S.tau[q];
            if (t < b1) { b2 = b1; i2 = i1; b1 = t; i1 = q; }
            else if (t < b2) { b2 = t; i2 = q; }
        }
        double tp = ZZ_INF; int li = -1;       // this warp's entry
        double t0 = ZZ_INF;                     // the earliest entry (decides whether the chain goes on)
        const int my_cnt = (nc - lane + 31) / 32;   // queue entries this lane looks after
        int taken = 0;
        for (int w = 0; w < nw; ++w) {
            double tw; int iw;
            zz_seq_argmin(b1, i1, tw, iw);
            if (w == 0) t0 = tw;
            if (!((tw < ZZ_INF) && (phase == 0 ? tw < P.T : (phase == 2 ? tw <= tend : true)))) break;   // (the same in every lane and warp)
            if (w == wid) { tp = tw; li = iw; }
            const bool mine = (i1 == iw);
            if (mine) { b1 = b2; i1 = i2; b2 = ZZ_INF; i2 = 0x7fffffff; taken++; }
            // a lane that has handed out both of the entries it kept may hide a third one that is earlier than everybody else's
            if (__any_sync(0xffffffffu, mine && taken == 2 && my_cnt > 2)) break;
        }
        if (!(t0 < ZZ_INF)) break;   // nothing will ever happen in this chain
        if (phase == 0) { if (!(t0 < P.T)) break; }
        else if (phase == 2) { if (!(t0 <= tend)) break; }
        const bool have = li >= 0;

        // ---- room for one trace record
        bool room = true;
        if (have && rec && tr_pos == tr_end) {
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
                room = false;
            } else { tr_pos = base; tr_end = base + ZZ_SEQ_RES; }
        }
        ZZ_SP_TOC(0);
        ZZ_SP_TIC();

        // ---- the proposal of coordinate i at tp (sfact.jl:118-121).  Everything that does not depend on the outcome is
        // started first: the loads of the own column of Z.Gamma (bound after a rejection), the uniform of the thinning test and
        // the logarithm of the uniform of the NEXT proposal time (poisson_time is a function of log u, src/poissontime.jl).
        int verdict = 0;
        int32_t jn = 0, jg = 0, be0 = 0, be1 = 0;
        double th_i = 0.0, tf_i = 0.0, xf_i = 0.0, xi = 0.0, c = 0.0, l = 0.0, lb = 0.0;
        double ra = 0.0, rb = 0.0, rtau = 0.0;   // the rejection branch: new bound and next proposal time
        uint32_t kc = 0;
        if (have && room) {
            jn = lo + li; jg = S.og ? S.og[li] : jn;   // chain-order id (matrices) and original id (streams, arrays)
            be0 = __ldg(Q.bcp + jn); be1 = __ldg(Q.bcp + jn + 1);
            const double gmu_i = __ldg(P.g.gmu + jg);
            th_i = S.th[li]; tf_i = S.tf[li]; xf_i = S.xf[li];
            xi = xf_i + th_i * (tp - tf_i);
            c = S.c[li];
            kc = S.kc[li];
            const bool short_col = (be1 - be0) <= 32;
            double gt, sx = 0.0, sth = 0.0;
            bool have_s = false;
            double u, Lrej;
            if (LOGIT) {
                const ZzLogit& L = P.lg;
                const int32_t a0 = S.a0[li], al = S.al[li];
                const double lk = (double)al / (double)L.k;
                double sacc = 0.0;
                for (int32_t r0 = 0; r0 < L.k; r0 += 16) {
                    const int32_t r = r0 + r16;
                    const bool act = r < L.k;
                    // sampled row r (scripts/logistic.jl:83,86): its entry of A and the row's constants
                    int32_t q0 = 0, len = 0; double w = 0.0, ya = 0.0, c0 = 0.0;
                    if (act) {
                        const double ur = zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc + (uint32_t)r));
                        int32_t i = (int32_t)(ur * (double)al);
                        if (i >= al) i = al - 1;
                        const int4 eh = __ldg(reinterpret_cast<const int4*>(Q.ent + a0 + i));
                        w = lk * __ldg(&Q.ent[a0 + i].val);
                        q0 = eh.y; len = eh.z;
                        const double2 rr = __ldg(reinterpret_cast<const double2*>(Q.rowrec + eh.x) + half);   // (y, ny) | (sn0, ns0)
                        const double2 r2 = __ldg(reinterpret_cast<const double2*>(Q.rowrec + eh.x) + (half ^ 1));
                        // lanes 0-15: y and sigmoidn; lanes 16-31: ny and nsigmoid
                        ya = half ? r2.y : rr.x;
                        c0 = half ? rr.y : r2.x;
                    }
                    // own column of Z.Gamma, first round only (speculative: needed when the proposal is rejected)
                    int32_t ck = 0; double cw = 0.0;
                    const bool cin = (r0 == 0) && short_col && (be0 + lane < be1);
                    if (cin) { ck = __ldg(Q.brow + be0 + lane) - lo; cw = __ldg(Q.bval + be0 + lane); }
                    if (r0 == 0) {
                        u = zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc + (uint32_t)L.k));
                        Lrej = zz_log(zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc + (uint32_t)L.k + 1u)));
                    }
                    double uu = 0.0;                                        // idot_moving!(At, row, ...), src/common.jl:33-42
                    for (int32_t qb = 0; qb < len; qb += 8) {
                        int32_t m[8]; double v[8];
#pragma unroll
                        for (int z = 0; z < 8; ++z) {
                            const bool ok = qb + z < len;
                            m[z] = ok ? __ldg(Q.rcoln + q0 + qb + z) - lo : 0;
                            v[z] = ok ? __ldg(L.rval + q0 + qb + z) : 0.0;
                        }
#pragma unroll
                        for (int z = 0; z < 8; ++z)
                            if (qb + z < len) uu += v[z] * zz_seq_pos(S, m[z], tp);
                    }
                    // sigmoidn(u) = 1 / (1 + exp(u)) on lanes 0-15, nsigmoid(u) = -(1 / (1 + exp(-u))) on lanes 16-31 (:34,56-57)
                    const double sg = 1.0 / (1.0 + zz_exp(half ? -uu : uu));
                    __syncwarp();   // (the scratch arrays are free again)
                    if (act) {
                        S.g[4 * r16 + half] = w * ya * (half ? -sg : sg);       // :87 | :88
                        S.g[4 * r16 + 2 + half] = -(w * ya * c0);               // :90 | :91 (control variate at the mode), subtracted
                    }
                    if (cin) { const double thk = S.th[ck]; S.cx[lane] = cw * (S.xf[ck] + thk * (tp - S.tf[ck])); S.ct[lane] = cw * thk; }
                    __syncwarp();
                    const int cnt = 4 * min(16, L.k - r0);
                    const int ccnt = (r0 == 0 && short_col) ? be1 - be0 : 0;
                    // three independent ordered sums, every lane the same (broadcast reads): the 4 k terms of the gradient in the
                    // reference's order, and the two column sums of the bound
                    {
                        double g0 = S.g[0], g1 = S.g[1], g2 = S.g[2], g3 = S.g[3];
                        for (int q = 0; q < cnt; q += 4) {   // (cnt is a multiple of 4; S.g holds 64 + 8 entries)
                            const double n0 = S.g[q + 4], n1 = S.g[q + 5], n2 = S.g[q + 6], n3 = S.g[q + 7];
                            sacc += g0; sacc += g1; sacc += g2; sacc += g3;
                            g0 = n0; g1 = n1; g2 = n2; g3 = n3;
                        }
                    }
                    if (ccnt > 0) zz_seq_sum2(S.cx, S.ct, ccnt, sx, sth);
                }
                have_s = short_col;
                gt = L.gamma0 * xi - sacc;                                      // :107
                kc += (uint32_t)L.k + 1u;
            } else {
                u = zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)kc);
                Lrej = zz_log(zz_u01(seed0, seed1, (uint64_t)jg, (uint64_t)(kc + 1u)));
                kc += 1u;
                if (Q.tcp) {
                    double d0, d1;
                    zz_seq_col_coop(S, Q.trow, Q.tval, __ldg(Q.tcp + jn), __ldg(Q.tcp + jn + 1), tp, lane, d0, d1);
                    gt = P.g.h ? d0 - __ldg(P.g.h + jg) : d0;
                } else {
                    zz_seq_col_coop(S, Q.brow, Q.bval, be0, be1, tp, lane, sx, sth);
                    have_s = true;
                    gt = sx;
                }
            }
            l = zz_pos(gt * th_i);                                  // fact_samplers.jl:28-30
            lb = zz_pos(S.a[li] + S.b[li] * (tp - S.told[li]));     // sfact.jl:70
            if (u * lb < l) verdict = 2;                            // sfact.jl:121
            else {                                                  // :137-140, prepared here, applied when the entry stands
                if (!have_s) zz_seq_col_coop(S, Q.brow, Q.bval, be0, be1, tp, lane, sx, sth);
                ra = c + (sx - gmu_i) * th_i;
                rb = c / 100 + th_i * sth;
                rtau = tp + zz_poisson_time_L(ra, rb, Lrej);
                verdict = 1;
            }
        }
        ZZ_SP_TOC(1);
        ZZ_SP_TIC();
        if (lane == 0) {
            V.idx[wid] = li; V.verdict[wid] = verdict; V.newtau[wid] = rtau;
            if (have && !room) V.stop = 1;
            if (wid == 0) V.acc_li = -1;
        }
        __syncthreads();
        // does this warp's entry stand?  every earlier entry was rejected and re-queued after (tp, li)
        bool valid = verdict != 0;
        for (int v = 0; v < wid; ++v) {
            const double nt = V.newtau[v];
            valid = valid && V.verdict[v] == 1 && (nt > tp || (nt == tp && V.idx[v] > li));
        }
        if (valid) nprop++;
        if (valid && verdict == 1) {
            if (lane == 0) { S.a[li] = ra; S.b[li] = rb; S.told[li] = tp; S.tau[li] = rtau; S.kc[li] = kc + 1u; }
        }
        // The (at most one) accepted flip of the step: its own coordinate is updated now -- nobody else touches it --, the
        // rescheduling of its neighbourhood is shared by all warps after the next barrier.
        if (valid && verdict == 2) {
            if (phase == 1) {   // look-ahead only: report the time, leave no trace
                if (lane == 0) { atomicMin(&C->smin_key[0], zz_key(tp)); V.stop = 1; }
            } else if (l >= lb && !P.v.adapt) {                              // sfact.jl:123-124
                if (lane == 0) {
                    atomicAdd(P.acc + jg, 1u);
                    if (atomicExch(&C->viol, 1u) == 0u) { C->viol_i = jg + 1; C->viol_t = tp; C->viol_l = l; C->viol_lb = lb; }
                    V.stop = 1;
                }
            } else {
                if (l >= lb) c *= P.v.factor;                                // :125-127
                nflip++;
                const int32_t tid = P.trace_map ? __ldg(P.trace_map + jg) : jg + 1;   // id in the trace (0: filtered out, src/trace.jl:275-290)
                if (lane == 0) {
                    atomicAdd(P.acc + jg, 1u);                               // :122
                    // moment sums of the segment that ends here (trace.jl:194), the event (sfact.jl:50-52,143), the flip (:130)
                    atomicAdd(P.s1 + jg, (xf_i + xi) * (tp - tf_i));
                    atomicAdd(P.s2 + jg, (tp - tf_i) * (xf_i * xf_i + xf_i * xi + xi * xi));
                    if (rec && tid != 0) {
                        double2* e = reinterpret_cast<double2*>(P.trace + tr_pos);
                        e[0] = make_double2(tp, __longlong_as_double((long long)tid));
                        e[1] = make_double2(xi, -th_i);
                    }
                    S.xf[li] = xi; S.tf[li] = tp; S.th[li] = -th_i; S.c[li] = c; S.kc[li] = kc;
                    V.acc_li = li; V.acc_tp = tp;
                }
                if (rec && tid != 0) tr_pos++;
            }
        }
        ZZ_SP_TOC(2);
        __syncthreads();   // the rejections of this step and the flip itself are applied
        ZZ_SP_TIC();
        const int ali = V.acc_li;
        if (ali >= 0) {
            // reschedule G1[i] = rows of column i of Z.Gamma (i among them; sfact.jl:131-135): warp 0 takes one neighbour per lane,
            // neighbours with a long column are served cooperatively by the other warps, one each at a time (by warp 0 too when
            // the chain has a single warp).  Every warp reads the column, so no hand-over is needed.
            const double atp = V.acc_tp;
            const int32_t an = lo + ali;
            const int32_t ae0 = __ldg(Q.bcp + an), ae1 = __ldg(Q.bcp + an + 1);
            for (int32_t base = ae0; base < ae1; base += 32) {
                const int32_t e = base + lane;
                const bool vld = e < ae1;
                int32_t jj = 0, lj = 0, f0 = 0, f1 = 0;
                if (vld) {
                    const int32_t jc = __ldg(Q.brow + e);   // chain-order id of the neighbour
                    lj = jc - lo; jj = S.og ? S.og[lj] : jc;
                    f0 = __ldg(Q.bcp + jc); f1 = __ldg(Q.bcp + jc + 1);
                }
                const bool islong = vld && (f1 - f0 > ZZ_SEQ_LONG);
                const unsigned int lm = __ballot_sync(0xffffffffu, islong);
                const bool mine_short = vld && !islong && wid == 0;
                const int nth = islong ? __popc(lm & ((1u << lane) - 1u)) : -1;            // rank among the long neighbours of the chunk
                const bool mine_long = islong && (nw == 1 ? true : (1 + nth % (nw - 1)) == wid);
                double gmu_j = 0.0, Lj = 0.0; uint32_t kj = 0;
                if (mine_short || mine_long) {
                    gmu_j = __ldg(P.g.gmu + jj);
                    kj = S.kc[lj];
                    Lj = zz_log(zz_u01(seed0, seed1, (uint64_t)jj, (uint64_t)(kj++)));
                }
                double sxj = 0.0, stj = 0.0;
                if (mine_short) zz_seq_col_serial(S, Q.brow, Q.bval, f0, f1, atp, sxj, stj);
                for (unsigned int ml = __ballot_sync(0xffffffffu, mine_long); ml; ml &= ml - 1u) {
                    const int src = __ffs((int)ml) - 1;
                    double cx, ct;
                    zz_seq_col_coop(S, Q.brow, Q.bval, __shfl_sync(0xffffffffu, f0, src), __shfl_sync(0xffffffffu, f1, src), atp, lane, cx, ct);
                    if (lane == src) { sxj = cx; stj = ct; }
                }
                if (mine_short || mine_long) {
                    const double cj = S.c[lj], thj = S.th[lj];
                    const double aj = cj + (sxj - gmu_j) * thj;               // fact_samplers.jl:51-52
                    const double bj = cj / 100 + thj * stj;
                    const double tj = atp + zz_poisson_time_L(aj, bj, Lj);
                    S.a[lj] = aj; S.b[lj] = bj; S.told[lj] = atp; S.tau[lj] = tj; S.kc[lj] = kj;
                }
            }
            ZZ_SP_TOC(3);
        }
    }
#ifdef ZZ_SEQ_PROF
    if (lane == 0 && phase == 0) for (int q = 0; q < 6; ++q) atomicAdd(&C->dbg[q], (unsigned long long)pc[q]);
#endif

    // (every warp leaves the loop at the same block barrier or at a test all of them evaluate alike)
    if (phase != 1) {
        __syncthreads();
        for (int32_t q = (int32_t)threadIdx.x; q < nc; q += (int32_t)blockDim.x) {
            const int32_t o = S.og ? S.og[q] : lo + q;
            double2* kq = reinterpret_cast<double2*>(P.v.kin + o);
            kq[0] = make_double2(S.th[q], S.tf[q]);
            reinterpret_cast<double*>(P.v.kin + o)[2] = S.xf[q];
            double2* pq = reinterpret_cast<double2*>(P.v.priv + o);
            pq[0] = make_double2(S.a[q], S.b[q]);
            pq[1] = make_double2(S.told[q], S.c[q]);
            P.v.tau[o] = S.tau[q];
            P.v.kctr[o] = S.kc[q];
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
}

extern "C" __global__ void __launch_bounds__(32 * ZZ_SEQ_MAXW) zz_seq_kernel(const __grid_constant__ ZzParams P, const __grid_constant__ ZzSeq Q)
{
    zz_seq_body<false>(P, Q);
}
extern "C" __global__ void __launch_bounds__(32 * ZZ_SEQ_MAXW) zz_seq_kernel_logit(const __grid_constant__ ZzParams P, const __grid_constant__ ZzSeq Q)
{
    zz_seq_body<true>(P, Q);
}

#endif  // ZZ_SEQ_CUH
