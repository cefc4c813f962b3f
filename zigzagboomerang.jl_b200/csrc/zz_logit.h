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
ZZ_HD void zz_process_node_logit(const ZzGraph& g, const ZzView& v, const ZzLogit& L, int32_t j, double H, int incl,
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

#endif  // ZZ_LOGIT_H
