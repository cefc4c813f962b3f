// zz_strong.h -- per-coordinate timeline of the STRONG-BOUND sparse sticky ZigZag (src/sparsestickyzz.jl, BASELINE config 4 as
// the reference runs it) in the windowed scheme, and of its per-coordinate generalisation asynchzz / sspdmp4 (src/asynchzz.jl).
// Kernel: zz_run_kernel_csr_strong (entries zzb_sspdmp3_run / zzb_sspdmp4_run); the same header runs inside the schedule emulation
// (oracle/zz_window_sim.cpp); contracts oracle/zz_oracle.c:zzo_sparsestickyzz_ctr / zzo_strongsticky_ctr.  It is written against
// the shared structures of zz_fast.h: the kernel instantiation is a mode constant.
//
// What is different from every other mode: a reflection reschedules NOBODY but the reflecting coordinate (constant bound
// a = c + grad_i theta_i valid for 1/c, then renewed; sparsestickyzz.jl:136-142,330-340,372-399).  The timeline of j therefore
// has OWN items only -- proposal (reflect), bound expiry (renew), reaching 0 (hit), thaw -- and neighbours enter only through
// the positions read at those items, advanced through their recorded events (time, velocity after; 0 = frozen).
// Record layout in this mode: priv = (bound value a, expiry time, remembered sign p = +-1 (:196-201,316,386-388), c);
// bits 30..31 of the draw counter = pending action.
#ifndef ZZ_STRONG_H
#define ZZ_STRONG_H

#include "zz_core.h"

struct ZzStrong {
    double c;        // > 0: this run uses the strong-bound sampler.  The bound constant itself is read PER COORDINATE from the private
                     // record (sspdmp3: the one constant of SparseStickyUpperBounds, :127-134, in every record; sspdmp4 / asynchzz:
                     // the vector c of StrongUpperBounds, src/asynchzz.jl:2-7,20-28)
    double kappa;    // (unused: the thaw rate of coordinate j is ZzView::kappa[j] -- StickyBarriers.kappa, stickyzz.jl:19-23)
    int32_t rule;    // 0 = :sticky of sparsestickyzz (re-enter with the remembered sign), 1 = :reversible (random sign),
                     // 2 = :sticky of asynchzz (src/asynchzz.jl:206-213: continue with the velocity saved when it froze)
    int32_t pad;
};

#define ZZ_SA_HIT 0u
#define ZZ_SA_REFLECT 1u
#define ZZ_SA_RENEW 2u
#define ZZ_SA_THAW 3u
#define ZZ_SA_SHIFT 30
#define ZZ_SA_KMASK 0x3fffffffu

// ab (:136-142, adapt = false) at time s followed by queue_time! (:144-172): the earliest of bound expiry, proposed
// reflection (rate 0.01 + a^+, poissontime.jl:86-92 with b = 0) and reaching 0 (:119-126)
ZZ_HD void zz_strong_queue(const ZzView& v, double cj, int32_t j, double s, double xs, double th, double gi,
                           uint32_t& k, double& ba, double& bexp, double& tau, uint32_t& act)
{
    ba = cj + gi * th;
    bexp = s + 1.0 / cj;
    const double trefl = s + zz_poisson_time3(ba, 0.0, 0.01, zz_u01(v.seed0, v.seed1, (uint64_t)j, k++));
    const double thit = (th * xs >= 0.0) ? ZZ_INF : s - xs / th;
    double t = bexp < trefl ? bexp : trefl;
    if (thit < t) t = thit;
    act = (thit == t) ? ZZ_SA_HIT : (trefl == t) ? ZZ_SA_REFLECT : ZZ_SA_RENEW;
    tau = t;
}

template <int NB>
ZZ_HD double zz_strong_grad(const ZzHood<NB>& hd, const ZzGraph& g, int32_t j, double s, double xown, double thown)
{   // idot(Gamma, j, u) over the sparse state (:42-51) [- h_j]: frozen coordinates contribute 0
    double gt, gx, gp, gm;
    zz_eval_hood<NB>(hd, g.same != 0, s, xown, thown, gt, gx, gp, gm);
    if (!g.same && g.h) gt = gt - g.h[j];
    return gt;
}

template <int NB>
ZZ_HD void zz_timeline_strong(ZzHood<NB>& hd, const ZzPool& pool, const ZzOwn& w, const ZzGraph& g, const ZzView& v,
                              const ZzStrong& S, int32_t j, double H, int incl, uint32_t flags0, ZzNodeOut& o)
{
    double th = w.th, tf = w.tf, xf = w.xf;
    double ba = w.a, bexp = w.b, psign = w.told, tau = w.tau;
    uint32_t act = w.k >> ZZ_SA_SHIFT, k = w.k & ZZ_SA_KMASK;
    uint32_t nprop = 0, nev = 0, flags = flags0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;
    for (int item = 0;; ++item) {
        const double s = tau;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        // neighbour events that precede (s, j): advance their anchors (no reschedule of j in this sampler)
        while (p < pool.n && (pool.t[p] < s || (pool.t[p] == s && pool.m[p] < hd.self))) {
            const double fs = pool.t[p], tha = pool.th[p];
            const int nm = pool.m[p];
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == nm) {
                    if (tha == 0.0) hd.xf[m] = -0.0 * hd.th[m];                                      // hit: frozen at 0
                    else if (hd.th[m] != 0.0) hd.xf[m] = hd.xf[m] + hd.th[m] * (fs - hd.tf[m]);      // reflection
                    hd.tf[m] = fs;                                                                   // (thaw: x stays 0)
                    hd.th[m] = tha;
                }
            }
            ++p;
        }
        if (act == ZZ_SA_THAW) {                                   // :291-329
            const double vi = S.rule == 1 ? (zz_u01(v.seed0, v.seed1, (uint64_t)j, k++) < 0.5 ? -1.0 : 1.0) : psign;
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = vi; }
            nev++;
            tf = s; th = vi;
            const double xs = xf + th * (s - tf);
            zz_strong_queue(v, w.c, j, s, xs, th, zz_strong_grad<NB>(hd, g, j, s, xs, th), k, ba, bexp, tau, act);
            continue;
        }
        const double xs = xf + th * (s - tf);
        if (act == ZZ_SA_HIT) {                                    // :341-371 (clusteralpha = 1)
            if ((xs < 0.0 ? -xs : xs) > 1e-7) flags |= ZZ_F_STICKY_ERR;
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = 0.0; }
            nev++;
            xf = -0.0 * th; tf = s; th = 0.0;
            act = ZZ_SA_THAW;
            tau = s - zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, k++)) / v.kappa[j];
            continue;
        }
        const double gi = zz_strong_grad<NB>(hd, g, j, s, xs, th);
        const double l = zz_pos(gi * th), lb = zz_pos(ba);          // lambda, :129-134 (b = 0)
        if (act == ZZ_SA_RENEW) {                                  // :330-340
            if (l > lb && !(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
            zz_strong_queue(v, w.c, j, s, xs, th, gi, k, ba, bexp, tau, act);
            continue;
        }
        nprop++;                                                   // :372-399
        if (zz_u01(v.seed0, v.seed1, (uint64_t)j, k++) * lb < l) {
            if (l > lb && !(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = -th; }
            nev++;
            xf = xs; tf = s; th = -th;
            if (S.rule == 0) psign = th > 0.0 ? 1.0 : -1.0;
            else if (S.rule == 2) psign = th;
            const double xn = xf + th * (s - tf);
            zz_strong_queue(v, w.c, j, s, xn, th, zz_strong_grad<NB>(hd, g, j, s, xn, th), k, ba, bexp, tau, act);
        } else {
            zz_strong_queue(v, w.c, j, s, xs, th, gi, k, ba, bexp, tau, act);
        }
    }
    o.a = ba; o.b = bexp; o.told = psign; o.tau = tau; o.c = w.c;
    o.k = k | (act << ZZ_SA_SHIFT); o.nprop = nprop; o.nflip = nev; o.flags = flags;
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1; o.nitems = 0; o.interior = 0u;
}

// columns of at most ZZ_NB entries (the chain and lattice targets of config 4)
ZZ_HD bool zz_process_node_strong(const ZzGraph& g, const ZzView& v, const ZzStrong& S, int32_t j, double H, int incl,
                                  uint32_t w0, uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    ZzHood<ZZ_NB> hd; ZzPool pool; uint32_t flags = 0; ZzOwn w;
    zz_load_own(v, j, w);
    if (!zz_gather_csr<ZZ_NB, false, true>(g, v, j, w0, cur, first_iter, hd, pool, flags)) return false;
    zz_timeline_strong<ZZ_NB>(hd, pool, w, g, v, S, j, H, incl, flags, o);
    return true;
}

// initial state (sparsestickystate :10-12, u.p :196-201, thaw clocks :218, queue fill :224-228): x0 == 0 starts frozen
ZZ_HD bool zz_init_node_strong(const ZzGraph& g, const ZzView& v, const ZzStrong& S, int32_t j, double t0)
{
    ZzHood<ZZ_NB> hd; ZzPool pool; uint32_t flags = 0;
    if (!zz_gather_csr<ZZ_NB, false, true>(g, v, j, 1u, 1u, true, hd, pool, flags)) return false;
    double th, tf, xf; uint32_t h0, h1;
    zz_ld_kin(v.kin + j, th, tf, xf, h0, h1);
    ZzPriv pr = zz_ld_priv(v.priv + j);
    uint32_t k = 0, act;
    double tau;
    if (th == 0.0) {
        pr.a = 0.0; pr.b = 0.0;
        if (S.rule != 2) pr.told = 1.0;   // (rule 2: zz_setup_kernel has left the velocity to continue with in the record)
        act = ZZ_SA_THAW;
        tau = t0 - zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, k++)) / v.kappa[j];
    } else {
        pr.told = S.rule == 2 ? th : (th > 0.0 ? 1.0 : -1.0);
        const double xs = xf + th * (t0 - tf);
        zz_strong_queue(v, pr.c, j, t0, xs, th, zz_strong_grad<ZZ_NB>(hd, g, j, t0, xs, th), k, pr.a, pr.b, tau, act);
    }
    v.priv[j] = pr;
    v.tau[j] = tau;
    v.kctr[j] = k | (act << ZZ_SA_SHIFT);
    return true;
}

#endif  // ZZ_STRONG_H
