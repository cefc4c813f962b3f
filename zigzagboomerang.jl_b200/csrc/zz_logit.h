// zz_logit.h -- subsampled logistic-regression target for the windowed local ZigZag (config 3 of BASELINE.json).
//
// Reference: scripts/logistic.jl.  The target's partial derivative is the closure `∇ϕmoving` (:107) =
//   gamma0 * x[j] - fdot_moving(A, At, j, t, x, theta, t', F, mu, y, ny, k)                         (:78-95)
// an unbiased estimate of  gamma0 x_j - sum_rows A[row,j] (y sigmoidn(A[row,:] x) + ny nsigmoid(A[row,:] x))  from k rows
// drawn uniformly from column j of the design matrix A, with the same terms at the mode mu subtracted (control variate).
// It is called as spdmp(∇ϕmoving, t0, x0, θ0, T, c, Zdrop, SelfMoving(), A, At, μ, y, m .- y, 10; adapt=true, factor=5)
// (:167): `SelfMoving` routes through the extended closure signature (src/sfact.jl:68) because the sampled rows touch
// coordinates outside G[j]; bounds and rescheduling use Z.Gamma = Gamma_drop (the sparsified Hessian at the mode) only.
//
// On the device the k row indices come from the proposing coordinate's OWN counter stream (the reference draws them from
// Julia's global RNG, :83,86 -- a documented deviation, equal in law), positions of the coordinates of a sampled row are
// flip-anchored like everywhere else (no in-place moves: idot_moving!, src/common.jl:33-42, becomes a read), and a
// coordinate is re-evaluated whenever the flip list of ANY coordinate sharing a design row with it changes (those are the
// ZZ_NB_TGT entries of its neighbour list; the host builds them from the pattern of A'A).
// Compiles for the device and for the host emulation (oracle/zz_window_sim.cpp), like zz_core.h.
#ifndef ZZ_LOGIT_H
#define ZZ_LOGIT_H

#include "zz_core.h"

struct ZzLogit {
    const int32_t* acp;    // [d+1] column offsets of the design matrix A (0-based)
    const int32_t* arow;   // row of every stored entry of A, column by column (0-based)
    const double* aval;
    const int32_t* rp;     // [n+1] offsets of the rows of A (= columns of At), coordinates ascending inside a row
    const int32_t* rcol;
    const double* rval;
    const double* y;       // [n] successes per design row
    const double* ny;      // [n] m - y
    const double* u0;      // [n] idot(At, row, mu), storage order (scripts/logistic.jl:90)
    double gamma0;         // prior precision
    int32_t k;             // rows drawn per evaluation
    int32_t n;             // design rows
};

// gamma0 x_j - fdot_moving(...) at time s; consumes L.k draws of coordinate j's stream (counter k is advanced).
// Every product is evaluated left to right as Julia's n-ary `*` does: ((l/k * vals[i]) * y[row]) * sigmoidn(u).
ZZ_HD double zz_logit_grad(const ZzLogit& L, const ZzView& v, int32_t j, double s, double xown, uint32_t w0,
                           uint32_t cur, uint32_t& k)
{
    const int32_t e0 = L.acp[j];
    const int32_t l = L.acp[j + 1] - e0;
    const double lk = (double)l / (double)L.k;
    double sacc = 0.0;
    for (int32_t r = 0; r < L.k; ++r) {
        const double ur = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
        int32_t i = (int32_t)(ur * (double)l);                  // uniform index into nzrange(A, j) (:83,86)
        if (i >= l) i = l - 1;
        const int32_t row = L.arow[e0 + i];
        const double w = lk * L.aval[e0 + i];
        double u = 0.0;                                         // idot_moving!(At, row, ...), src/common.jl:33-42
        const int32_t q1 = L.rp[row + 1];
        for (int32_t q = L.rp[row]; q < q1; ++q) {
            const int32_t m = L.rcol[q];
            double xm = xown, thm;
            if (m != j) zz_nb_state(v, m, s, j, w0, cur, xm, thm);
            u += L.rval[q] * xm;
        }
        const double yr = L.y[row], nyr = L.ny[row], u0 = L.u0[row];
        sacc += w * yr * zz_sigmoidn(u);                        // :87-88
        sacc += w * nyr * zz_nsigmoid(u);
        sacc -= w * yr * zz_sigmoidn(u0);                       // :90-91 (control variate at the mode)
        sacc -= w * nyr * zz_nsigmoid(u0);
    }
    return L.gamma0 * xown - sacc;                              // :107
}

// idot(Z.Gamma, j, x(s)) and idot(Z.Gamma, j, theta) with +own / -own velocity: the ZZ_NB_BND entries of zz_eval.
ZZ_HD void zz_eval_bnd(const ZzGraph& g, const ZzView& v, int32_t j, double s, int32_t key_idx, double xown,
                       double thown, uint32_t w0, uint32_t cur, double& gx, double& gp, double& gm)
{
    double ax = 0.0, ap = 0.0, am = 0.0;
    const int32_t e1 = g.nptr[j + 1];
    for (int32_t e = g.nptr[j]; e < e1; ++e) {
        if (!(g.nfl[e] & ZZ_NB_BND)) continue;
        const int32_t k = g.nidx[e];
        const double wb = g.nwb[e];
        if (k == j) {
            ax += wb * xown; ap += wb * thown; am += wb * (-thown);
        } else {
            double x, th;
            zz_nb_state(v, k, s, key_idx, w0, cur, x, th);
            ax += wb * x; ap += wb * th; am += wb * th;
        }
    }
    gx = ax; gp = ap; gm = am;
}

// Timeline of coordinate j inside the window (zz_process_node_slow with the logistic target): own proposals evaluate the
// subsampled gradient (src/sfact.jl:118 through the SelfMoving closure), reschedules use ab of fact_samplers.jl:50-54.
// List-walking version: every timeline item re-reads the neighbour records from memory (one dependent round trip per entry);
// used only for columns with more than ZZ_LNB bound / trigger entries.
ZZ_HD void zz_process_node_logit_walk(const ZzGraph& g, const ZzView& v, const ZzLogit& L, int32_t j, double H, int incl,
                                 uint32_t w0, uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    double th, tf, xf; uint32_t hh0, hh1;
    zz_ld_kin(v.kin + j, th, tf, xf, hh0, hh1);
    const ZzPriv pr = zz_ld_priv(v.priv + j);
    double a = pr.a, b = pr.b, told = pr.told, c = pr.c;
    double c100 = c / 100;
    double tau = zz_ld(v.tau + j);
    uint32_t k = zz_ld32(v.kctr + j);
    const double gmu = g.gmu[j];
    uint32_t nprop = 0, nflip = 0, flags = 0;
    double last_t = -ZZ_INF; int32_t last_i = -1;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    o.interior = 0u;
    uint32_t nitems = 0;

    for (int item = 0;; ++item) {
        double nt = ZZ_INF; int32_t ni = 0x7fffffff;
        if (!first_iter) zz_next_trigger(g, v, j, last_t, last_i, w0, cur, nt, ni);
        const bool own = (tau < nt) || (tau == nt && j < ni);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        ++nitems;
        const double xs = xf + th * (s - tf);
        double gx, gp, gm, gth;
        if (own) {
            const double gt = zz_logit_grad(L, v, j, s, xs, w0, cur, k);   // draws k .. k + L.k - 1
            zz_eval_bnd(g, v, j, s, j, xs, th, w0, cur, gx, gp, gm);
            const double l = zz_pos(gt * th);                 // fact_samplers.jl:28-30
            const double lb = zz_pos(a + b * (s - told));     // sfact.jl:70
            const double u = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
            nprop++;
            gth = gp;
            if (u * lb < l) {                                 // sfact.jl:121
                if (l >= lb) {                                // sfact.jl:123-128
                    if (v.adapt) { c *= v.factor; c100 = c / 100; }
                    else if (!(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
                }
                if (nflip == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m)
                    if (m == (int)nflip) o.fl[m] = s;
                nflip++;
                xf = xs; tf = s; th = -th;                    // dynamics.jl:46-49
                gth = gm;
            }
        } else {
            last_t = nt; last_i = ni;
            zz_eval_bnd(g, v, j, s, ni, xs, th, w0, cur, gx, gp, gm);
            gth = gp;
        }
        a = c + (gx - gmu) * th;                              // fact_samplers.jl:51
        b = c100 + th * gth;                                  // fact_samplers.jl:52
        told = s;
        tau = s + zz_poisson_time(a, b, zz_u01(v.seed0, v.seed1, (uint64_t)j, k++));   // sfact.jl:134,139
    }
    o.a = a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k; o.nprop = nprop; o.nflip = nflip; o.flags = flags;
    o.hdr0 = hh0; o.hdr1 = hh1; o.nitems = nitems;
}

// ---------------------------------------------------------------------------------------------------------------
// Gather-first version (what zz_fast.h does for short columns, for columns of up to ZZ_LNB bound / trigger entries): the
// records of the neighbours that enter the bound or reschedule j are fetched ONCE per evaluation, in batches of independent
// loads, into a per-thread scratch (local memory, L1-resident); their recorded flips are merged into one pool ordered by
// (time, coordinate) and applied incrementally, so a timeline item costs arithmetic over the scratch instead of a chain of
// L2 round trips per neighbour.  Same operations in the same order as the list-walking version -- bit-identical results.
// (Measured with the list-walking version alone: 2.6 ms per relaxation pass on config 3, whose dense regressors have 147
// bound entries and are dirtied by every flip.)
#define ZZ_LNB 192     // cached bound / trigger entries of column j (j included)
#define ZZ_LPOOL 32    // neighbour flips merged per coordinate and window
#define ZZ_LBATCH 8    // neighbour records fetched together

struct ZzLHood {
    int n;
    int32_t idx[ZZ_LNB];
    double wb[ZZ_LNB];
    uint8_t fl[ZZ_LNB];
    double th[ZZ_LNB], tf[ZZ_LNB], xf[ZZ_LNB];
};

struct ZzLPool {
    int n;
    double t[ZZ_LPOOL];
    int32_t id[ZZ_LPOOL];   // flipping coordinate
    int32_t m[ZZ_LPOOL];    // its position in the scratch
};

ZZ_HD void zz_lpool_add(ZzLPool& pool, double fs, int32_t id, int32_t m, uint32_t& flags)
{
    int p = pool.n;
    if (p == ZZ_LPOOL) {   // full: the window will be retried shorter; keep the EARLIEST flips so that what is evaluated stays causal
        flags |= ZZ_F_OVERFLOW;
        if (!(pool.t[p - 1] > fs || (pool.t[p - 1] == fs && pool.id[p - 1] > id))) return;
        pool.n = --p;
    }
    while (p > 0 && (pool.t[p - 1] > fs || (pool.t[p - 1] == fs && pool.id[p - 1] > id))) {
        pool.t[p] = pool.t[p - 1]; pool.id[p] = pool.id[p - 1]; pool.m[p] = pool.m[p - 1];
        --p;
    }
    pool.t[p] = fs; pool.id[p] = id; pool.m[p] = m; pool.n++;
}

ZZ_HD void zz_process_node_logit(const ZzGraph& g, const ZzView& v, const ZzLogit& L, int32_t j, double H, int incl,
                                 uint32_t w0, uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    // ---- which entries of the merged list are cached: bound or trigger ones (target-only entries are read on demand)
    ZzLHood hd;
    {
        int n = 0;
        const int32_t e1 = g.nptr[j + 1];
        for (int32_t e = g.nptr[j]; e < e1; ++e) {
            const uint32_t fl = g.nfl[e];
            if (!(fl & (ZZ_NB_BND | ZZ_NB_TRIG))) continue;
            if (n == ZZ_LNB) { zz_process_node_logit_walk(g, v, L, j, H, incl, w0, cur, first_iter, o); return; }
            hd.idx[n] = g.nidx[e]; hd.fl[n] = (uint8_t)fl; hd.wb[n] = g.nwb[e];
            ++n;
        }
        hd.n = n;
    }
    ZzOwn w;
    zz_load_own(v, j, w);
    // ---- gather: ZZ_LBATCH records at a time (independent loads in flight), then their flip lists into the pool
    ZzLPool pool; pool.n = 0;
    uint32_t flags = 0;
    for (int b0 = 0; b0 < hd.n; b0 += ZZ_LBATCH) {
        double bth[ZZ_LBATCH], btf[ZZ_LBATCH], bxf[ZZ_LBATCH]; uint32_t bh0[ZZ_LBATCH], bh1[ZZ_LBATCH];
#pragma unroll
        for (int q = 0; q < ZZ_LBATCH; ++q) {
            bth[q] = 0.0; btf[q] = 0.0; bxf[q] = 0.0; bh0[q] = 0; bh1[q] = 0;
            if (b0 + q < hd.n && hd.idx[b0 + q] != j) zz_ld_kin(zz_kin_at<false>(v, hd.idx[b0 + q]), bth[q], btf[q], bxf[q], bh0[q], bh1[q]);
        }
#pragma unroll
        for (int q = 0; q < ZZ_LBATCH; ++q) {
            if (b0 + q < hd.n) {
                const int m = b0 + q;
                hd.th[m] = bth[q]; hd.tf[m] = btf[q]; hd.xf[m] = bxf[q];
                if (!first_iter && hd.idx[m] != j) {
                    int slot;
                    const uint32_t cnt = zz_pick_slot(bh0[q], bh1[q], w0, cur, slot);
                    if (cnt) {
                        const double* fl = zz_flips_at<false>(v, hd.idx[m]) + slot * ZZ_MAXFLIP;
                        for (uint32_t r = 0; r < cnt; ++r) zz_lpool_add(pool, zz_ld(fl + r), hd.idx[m], m, flags);
                    }
                }
            }
        }
    }

    double th = w.th, tf = w.tf, xf = w.xf;
    double a = w.a, b = w.b, told = w.told, c = w.c;
    double c100 = c / 100;
    double tau = w.tau;
    uint32_t k = w.k;
    const double gmu = g.gmu[j];
    uint32_t nprop = 0, nflip = 0, nitems = 0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    o.interior = 0u;
    int p = 0;
    for (int item = 0;; ++item) {
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int32_t ni = p < pool.n ? pool.id[p] : 0x7fffffff;
        const bool own = (tau < nt) || (tau == nt && j < ni);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        if (!own) {   // a cached neighbour flips at s: advance its anchor; reschedule j only if it is a trigger
            const int m = pool.m[p];
            hd.xf[m] = hd.xf[m] + hd.th[m] * (s - hd.tf[m]);
            hd.tf[m] = s;
            hd.th[m] = -hd.th[m];
            ++p;
            if (!(hd.fl[m] & ZZ_NB_TRIG)) continue;
        }
        ++nitems;
        const double xs = xf + th * (s - tf);
        double gt = 0.0;
        if (own) gt = zz_logit_grad(L, v, j, s, xs, w0, cur, k);   // draws k .. k + L.k - 1
        // idot(Z.Gamma, j, x(s)), idot(Z.Gamma, j, theta) with +own / -own velocity, storage order (common.jl:16-24)
        double ax = 0.0, ap = 0.0, am = 0.0;
        for (int m = 0; m < hd.n; ++m) {
            if (!(hd.fl[m] & ZZ_NB_BND)) continue;
            const double wb = hd.wb[m];
            if (hd.idx[m] == j) { ax += wb * xs; ap += wb * th; am += wb * (-th); }
            else {
                const double x = hd.xf[m] + hd.th[m] * (s - hd.tf[m]);
                ax += wb * x; ap += wb * hd.th[m]; am += wb * hd.th[m];
            }
        }
        double gth = ap;
        if (own) {
            const double l = zz_pos(gt * th);                 // fact_samplers.jl:28-30
            const double lb = zz_pos(a + b * (s - told));     // sfact.jl:70
            const double u = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
            nprop++;
            if (u * lb < l) {                                 // sfact.jl:121
                if (l >= lb) {                                // sfact.jl:123-128
                    if (v.adapt) { c *= v.factor; c100 = c / 100; }
                    else if (!(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
                }
                if (nflip == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m)
                    if (m == (int)nflip) o.fl[m] = s;
                nflip++;
                xf = xs; tf = s; th = -th;                    // dynamics.jl:46-49
                gth = am;
            }
        }
        a = c + (ax - gmu) * th;                              // fact_samplers.jl:51
        b = c100 + th * gth;                                  // fact_samplers.jl:52
        told = s;
        tau = s + zz_poisson_time(a, b, zz_u01(v.seed0, v.seed1, (uint64_t)j, k++));   // sfact.jl:134,139
    }
    o.a = a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k; o.nprop = nprop; o.nflip = nflip; o.flags = flags;
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1; o.nitems = nitems;
}

#endif  // ZZ_LOGIT_H
