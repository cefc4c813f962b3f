// zz_fast.h -- gather-first evaluation of one coordinate's timeline (included by zz_core.h).
//
// zz_process_node_slow walks the neighbour lists in global memory once per timeline item: a chain of ~100
// dependent L2 round trips (tens of microseconds per coordinate on a B200).  Here the whole neighbourhood
// (kinematic records + recorded flips) is fetched FIRST with independent loads -- four round trips in total --
// and the timeline then runs out of registers: neighbour flips are merged into one time-ordered pool and applied
// incrementally, so every evaluation sees exactly the neighbour state the slow path would compute, bit for bit.
#ifndef ZZ_FAST_H
#define ZZ_FAST_H

#if defined(ZZ_PROF_NODE) && defined(__CUDA_ARCH__)
__device__ unsigned long long* zz_dbg_ptr;   // set by the kernel; segment k accumulates cycles since segment k-1
__device__ long long zz_dbg_last;
__device__ int zz_dbg_on;     // segments are accumulated only while set (ZZ_PROF_TAIL: during the single-CTA tail passes)
#define ZZ_SEG(k) do { if (blockIdx.x == 0 && threadIdx.x == 0 && zz_dbg_on) { long long _c = clock64(); if (k) zz_dbg_ptr[k] += (unsigned long long)(_c - zz_dbg_last); zz_dbg_last = _c; } } while (0)
#define ZZ_SEGCOUNT() do { if (blockIdx.x == 0 && threadIdx.x == 0 && zz_dbg_on) zz_dbg_ptr[6] += 1; } while (0)
#else
#define ZZ_SEG(k) do { } while (0)
#define ZZ_SEGCOUNT() do { } while (0)
#endif

#define ZZ_NB 8     // neighbourhood capacity of the gathered path (entries of column j, j included)
#define ZZ_NB_WIDE 32   // sticky / Boomerang / refreshment samplers: columns of 9 .. 32 entries take a second, out-of-line instantiation
#define ZZ_POOL 16  // neighbour flips merged per coordinate and window

// frontier state of the coordinate being evaluated (loaded before the neighbourhood so that all loads overlap)
struct ZzOwn {
    double th, tf, xf, a, b, told, c, tau;
    uint32_t k, hdr0, hdr1;
};

ZZ_HD void zz_load_own(const ZzView& v, int32_t j, ZzOwn& w)
{
    zz_ld_kin(v.kin + j, w.th, w.tf, w.xf, w.hdr0, w.hdr1);
    const ZzPriv pr = zz_ld_priv(v.priv + j);
    w.a = pr.a; w.b = pr.b; w.told = pr.told; w.c = pr.c;
    w.tau = zz_ld(v.tau + j);
    w.k = zz_ld32(v.kctr + j);
}

template <int NB>
struct ZzHood {
    int n;        // entries in storage order
    int self;     // position of j itself
    double wb[NB], wt[NB];
    uint32_t fl[NB];
    double th[NB], tf[NB], xf[NB];
};

// FactBoomerang only: centre of the rotation of each neighbour (Z.mu).  Kept out of ZzHood so that the ZigZag kernels
// compile exactly as before.
template <int NB>
struct ZzHoodMu {
    double mu[NB];
};

struct ZzPool {
    int n;
    double t[ZZ_POOL];
    int m[ZZ_POOL];      // neighbourhood position of the flipping coordinate
    double th[ZZ_POOL];  // sticky only: its velocity after the event (0 = freeze)
};

// merge the recorded flips of neighbour position m into the pool, ordered by (time, position)
template <bool VEL = true>
ZZ_HD void zz_pool_add(ZzPool& pool, double fs, int m, uint32_t& flags, double tha = 0.0)
{
    int p = pool.n;
    if (p == ZZ_POOL) { flags |= ZZ_F_OVERFLOW; return; }
    while (p > 0 && (pool.t[p - 1] > fs || (pool.t[p - 1] == fs && pool.m[p - 1] > m))) {
        pool.t[p] = pool.t[p - 1]; pool.m[p] = pool.m[p - 1];
        if (VEL) pool.th[p] = pool.th[p - 1];
        --p;
    }
    pool.t[p] = fs; pool.m[p] = m; pool.n++;
    if (VEL) pool.th[p] = tha;
}

template <int NB, bool MG, bool VEL>
ZZ_HD void zz_gather_flips(const ZzView& v, const int32_t (&idx)[NB], const uint32_t (&h0)[NB], const uint32_t (&h1)[NB],
                           int n, int self, uint32_t w0, uint32_t cur, ZzPool& pool, uint32_t& flags)
{
    pool.n = 0;
    if constexpr (VEL) {   // sticky / Boomerang lists carry the velocity after each event (compile-time: the plain kernels stay small)
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (m < n && m != self) {
                int slot;
                const uint32_t cnt = zz_pick_slot(h0[m], h1[m], w0, cur, slot);
                if (cnt) {
                    const double* fl = zz_flips_at<MG>(v, idx[m]) + slot * ZZ_MAXFLIP;
                    const double* ft = v.fth + ((size_t)idx[m] * 2 + slot) * ZZ_MAXFLIP;
                    for (uint32_t q = 0; q < cnt; ++q) zz_pool_add(pool, zz_ld(fl + q), m, flags, zz_ld(ft + q));
                }
            }
        }
    } else {
        // Plain lists: the first two entries of every non-empty neighbour list are fetched together (independent loads, one
        // round trip) before anything is merged; longer lists (rare) continue entry by entry.  Insertion order -- neighbour by
        // neighbour, entries ascending -- is unchanged, so the merged pool is the same.
        uint32_t cnt[NB]; const double* fl[NB]; double f0[NB], f1[NB];
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            cnt[m] = 0; fl[m] = nullptr; f0[m] = 0.0; f1[m] = 0.0;
            if (m < n && m != self) {
                int slot;
                cnt[m] = zz_pick_slot(h0[m], h1[m], w0, cur, slot);
                if (cnt[m]) fl[m] = zz_flips_at<MG>(v, idx[m]) + slot * ZZ_MAXFLIP;
            }
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (cnt[m]) {
#if defined(__CUDA_ARCH__)
                const double2 u = __ldcg(reinterpret_cast<const double2*>(fl[m]));   // slots are 48-byte aligned
                f0[m] = u.x; f1[m] = u.y;
#else
                f0[m] = fl[m][0]; f1[m] = fl[m][1];
#endif
            }
        }
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            if (cnt[m]) {
                zz_pool_add<false>(pool, f0[m], m, flags);
                if (cnt[m] > 1) zz_pool_add<false>(pool, f1[m], m, flags);
                for (uint32_t q = 2; q < cnt[m]; ++q) zz_pool_add<false>(pool, zz_ld(fl[m] + q), m, flags);
            }
        }
    }
}

// General sparse column (<= NB entries).  Returns false when the column is longer (caller uses the slow path).
template <int NB, bool MG, bool VEL = false>
ZZ_HD bool zz_gather_csr(const ZzGraph& g, const ZzView& v, int32_t j, uint32_t w0, uint32_t cur, bool first_iter,
                         ZzHood<NB>& hd, ZzPool& pool, uint32_t& flags, ZzHoodMu<NB>* hm = nullptr)
{
    const int32_t e0 = g.nptr[j];
    const int n = g.nptr[j + 1] - e0;
    if (n > NB) return false;
    hd.n = n; hd.self = 0;
    int32_t idx[NB]; uint32_t h0[NB], h1[NB];
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        if (m < n) {
            idx[m] = g.nidx[e0 + m]; hd.fl[m] = g.nfl[e0 + m]; hd.wb[m] = g.nwb[e0 + m];
            hd.wt[m] = g.same ? 0.0 : g.nwt[e0 + m];
        } else { idx[m] = -1; hd.fl[m] = 0; hd.wb[m] = 0.0; hd.wt[m] = 0.0; }
    }
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        h0[m] = 0; h1[m] = 0; hd.th[m] = 0.0; hd.tf[m] = 0.0; hd.xf[m] = 0.0;
        if (m < n) {
            if (idx[m] == j) hd.self = m;
            else zz_ld_kin(zz_kin_at<MG>(v, idx[m]), hd.th[m], hd.tf[m], hd.xf[m], h0[m], h1[m]);
        }
    }
    if (hm) {
#pragma unroll
        for (int m = 0; m < NB; ++m) hm->mu[m] = (m < n) ? v.bmu[idx[m]] : 0.0;
    }
    pool.n = 0;
    if (!first_iter) zz_gather_flips<NB, MG, VEL>(v, idx, h0, h1, n, hd.self, w0, cur, pool, flags);
    return true;
}

// 5-point lattice: column j = {j-M, j-1, j, j+1, j+M} (those that exist), weights -1 and shift + degree.
template <bool MG, bool VEL = false>
ZZ_HD void zz_gather_grid(const ZzGraph& g, const ZzView& v, int32_t j, uint32_t w0, uint32_t cur, bool first_iter,
                          ZzHood<5>& hd, ZzPool& pool, uint32_t& flags, ZzHoodMu<5>* hm = nullptr)
{
    const int32_t M = g.grid_m, N = g.grid_n;
    const int32_t col = zz_grid_col(g, j), row = j - col * M;
    int32_t idx[5]; uint32_t h0[5], h1[5];
    int n = 0;
    if (col > 0) idx[n++] = j - M;
    if (row > 0) idx[n++] = j - 1;
    const int self = n;
    idx[n++] = j;
    if (row < M - 1) idx[n++] = j + 1;
    if (col < N - 1) idx[n++] = j + M;
    hd.n = n; hd.self = self;
    const double diag = g.grid_diag[n - 1];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        if (m >= n) idx[m] = -1;
        hd.wb[m] = (m == self) ? diag : -1.0; hd.wt[m] = 0.0;
        hd.fl[m] = (m < n) ? ((m == self) ? (ZZ_NB_TGT | ZZ_NB_BND) : (ZZ_NB_TGT | ZZ_NB_BND | ZZ_NB_TRIG)) : 0u;
        h0[m] = 0; h1[m] = 0; hd.th[m] = 0.0; hd.tf[m] = 0.0; hd.xf[m] = 0.0;
    }
#pragma unroll
    for (int m = 0; m < 5; ++m)
        if (m < n && m != self) zz_ld_kin(zz_kin_at<MG>(v, idx[m]), hd.th[m], hd.tf[m], hd.xf[m], h0[m], h1[m]);
    if (hm) {
#pragma unroll
        for (int m = 0; m < 5; ++m) hm->mu[m] = (m < n) ? v.bmu[idx[m]] : 0.0;
    }
    pool.n = 0;
    ZZ_SEG(1);
    if (!first_iter) zz_gather_flips<5, MG, VEL>(v, idx, h0, h1, n, self, w0, cur, pool, flags);
}

// idot's over the gathered column at time s, storage order (common.jl:16-24)
template <int NB>
ZZ_HD void zz_eval_hood(const ZzHood<NB>& hd, bool same, double s, double xown, double thown, double& gt,
                        double& gx, double& gp, double& gm)
{
    double at = 0.0, ax = 0.0, ap = 0.0, am = 0.0;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        if (m < hd.n) {
            const bool self = (m == hd.self);
            const double th = self ? thown : hd.th[m];
            const double x = self ? xown : (hd.xf[m] + hd.th[m] * (s - hd.tf[m]));
            if (hd.fl[m] & ZZ_NB_BND) {
                const double wb = hd.wb[m];
                ax += wb * x; ap += wb * th; am += wb * (self ? -th : th);
            }
            if (!same && (hd.fl[m] & ZZ_NB_TGT)) at += hd.wt[m] * x;
        }
    }
    gt = same ? ax : at; gx = ax; gp = ap; gm = am;
}

// Timeline of coordinate j from gathered data; identical arithmetic to zz_process_node_slow.
template <int NB, bool LB>
ZZ_HD void zz_timeline(ZzHood<NB>& hd, const ZzPool& pool, const ZzOwn& w, const ZzGraph& g, const ZzView& v,
                       int32_t j, double H, int incl, uint32_t flags0, ZzNodeOut& o)
{
    double th = w.th, tf = w.tf, xf = w.xf;
    const uint32_t hh0 = w.hdr0, hh1 = w.hdr1;
    double a = w.a, b = w.b, told = w.told, c = w.c;
    double c100 = c / 100;
    double tau = w.tau;
    const bool lbm = LB;   // compile-time: the plain ZigZag kernels carry none of the LocalBound logic
    bool renew = LB && (w.k & ZZ_RENEW_BIT) != 0;
    uint32_t k = w.k & ~ZZ_RENEW_BIT;
    const double gmu = g.grid_m ? 0.0 : g.gmu[j];
    const double hj = (!g.same && g.h) ? g.h[j] : 0.0;
    const bool has_h = (!g.same && g.h);
    uint32_t nprop = 0, nflip = 0, flags = flags0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;
    int item = 0;
    uint32_t nitems = 0;
    for (;; ++item) {
        ZZ_SEGCOUNT();
        ++nitems;
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int nm = p < pool.n ? pool.m[p] : 0x7fffffff;
        const bool own = (tau < nt) || (tau == nt && hd.self < nm);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        if (!own) {
            // neighbour at position nm flips at s: advance its anchor; reschedule j only if it triggers us
            bool trig = false;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == nm) {
                    hd.xf[m] = hd.xf[m] + hd.th[m] * (s - hd.tf[m]);
                    hd.tf[m] = s;
                    hd.th[m] = -hd.th[m];
                    trig = (hd.fl[m] & ZZ_NB_TRIG) != 0;
                }
            }
            ++p;
            if (!trig) continue;
        }
        // the draws of this item depend only on the counter: start them (and the logarithm of the rescheduling draw)
        // before the neighbourhood evaluation so that the two dependency chains overlap
        const bool prop = own && !(lbm && renew);             // an expired LocalBound is renewed without thinning
        const uint32_t kr = prop ? k + 1u : k;
        const double L2 = zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, kr));
        const double u1 = prop ? zz_u01(v.seed0, v.seed1, (uint64_t)j, k) : 0.0;
        k = kr + 1u;
        // one evaluation of the column at s shared by both kinds of item (keeps diverged lanes on the same code)
        const double xs = xf + th * (s - tf);
        double gt, gx, gp, gm;
        zz_eval_hood<NB>(hd, g.same != 0, s, xs, th, gt, gx, gp, gm);
        double gth = gp;
        if (has_h) gt = gt - hj;
        if (prop) {
            const double l = zz_pos(gt * th);                 // fact_samplers.jl:28-30
            const double lb = zz_pos(a + b * (s - told));     // sfact.jl:70
            nprop++;
            if (u1 * lb < l) {                                // sfact.jl:121
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
        }
        a = c + (lbm ? gt : gx - gmu) * th;                   // fact_samplers.jl:51 / local.jl:3
        b = c100 + th * gth;                                  // fact_samplers.jl:52 / local.jl:4 (c100 = c / 100)
        told = s;
        tau = zz_next_time(s, zz_poisson_time_L(a, b, L2), c, th, lbm, renew);   // sfact.jl:134,139 / local.jl:38-39
    }
    o.a = a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k | (renew ? ZZ_RENEW_BIT : 0u); o.nprop = nprop; o.nflip = nflip; o.flags = flags;
    o.hdr0 = hh0; o.hdr1 = hh1; o.nitems = nitems;
}

// ---------------------------------------------------------------------------------------------------------------
// Sticky ZigZag (sspdmp, src/ss_fact.jl:78-157) seen from coordinate j.  Own items: proposal (:124-152), freeze when the
// coordinate reaches 0 (:87-107), thaw after its Exp(kappa_j) clock (:108-123).  Neighbour items: any velocity change
// of a neighbour (flip / freeze / thaw) reschedules j if j is moving (:100-106,117-123,140-146).
ZZ_HD double zz_freezing_time(double x, double th)
{   // ss_fact.jl:10-16
    if (th * x >= 0.0) return ZZ_INF;
    return -x / th;
}

template <int NB>
ZZ_HD void zz_timeline_sticky(ZzHood<NB>& hd, const ZzPool& pool, const ZzOwn& w, const ZzGraph& g, const ZzView& v,
                              int32_t j, double H, int incl, uint32_t flags0, ZzNodeOut& o)
{
    double th = w.th, tf = w.tf, xf = w.xf;
    double a = w.a, b = w.b, told = w.told, c = w.c;
    double c100 = c / 100;
    double tau = w.tau;
    bool fbit = (w.k & ZZ_RENEW_BIT) != 0;
    uint32_t k = w.k & ~ZZ_RENEW_BIT;
    bool frozen = (th == 0.0);
    double thf = frozen ? w.a : 0.0;          // the saved velocity of a frozen coordinate travels in the `a` slot
    const double kap = v.kappa[j];
    const double gmu = g.grid_m ? 0.0 : g.gmu[j];
    const bool has_h = (!g.same && g.h);
    const double hj = has_h ? g.h[j] : 0.0;
    uint32_t nprop = 0, nflip = 0, nev = 0, flags = flags0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;

    for (int item = 0;; ++item) {
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int nm = p < pool.n ? pool.m[p] : 0x7fffffff;
        const bool own = (tau < nt) || (tau == nt && hd.self < nm);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        bool propose = false;
        if (!own) {
            const double tha = pool.th[p];
            bool trig = false;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == nm) {
                    if (tha == 0.0) { hd.xf[m] = -0.0 * hd.th[m]; }                          // freeze: x = -0*theta (:92)
                    else if (hd.th[m] != 0.0) { hd.xf[m] = hd.xf[m] + hd.th[m] * (s - hd.tf[m]); }  // flip
                    hd.tf[m] = s;
                    hd.th[m] = tha;
                    trig = (hd.fl[m] & ZZ_NB_TRIG) != 0;
                }
            }
            ++p;
            if (!trig || frozen) continue;           // frozen coordinates are not rescheduled
            if ((v.sticky & ZZ_STICKY_STRONG_UB) && tha == 0.0) continue;   // strong_upperbounds: a freeze reschedules nobody (:97)
        } else if (frozen) {                         // thaw (:108-116): restore the speed, then reschedule below
            th = thf; thf = 0.0; tf = s; frozen = false;
            if (v.sticky & ZZ_STICKY_REVERSIBLE)     // reversible: re-enter with a random sign (:111-113)
                th *= (zz_u01(v.seed0, v.seed1, (uint64_t)j, k++) < 0.5 ? -1.0 : 1.0);
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = th; }
            nev++;
        } else if (fbit) {                           // freeze (:87-96)
            const double xs = xf + th * (s - tf);
            if ((xs < 0.0 ? -xs : xs) > 1e-8) flags |= ZZ_F_STICKY_ERR;
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = 0.0; }
            nev++;
            xf = -0.0 * th; tf = s; thf = th; th = 0.0; frozen = true; fbit = false; told = s;
            tau = s - zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, k++)) / kap;   // :96
            continue;
        } else {
            propose = true;
        }
        const double xs = xf + th * (s - tf);
        double gt, gx, gp, gm;
        zz_eval_hood<NB>(hd, g.same != 0, s, xs, th, gt, gx, gp, gm);
        double gth = gp;
        double xnow = xs;
        if (propose) {
            if (has_h) gt = gt - hj;
            const double l = zz_pos(gt * th);
            const double lb = zz_pos(a + b * (s - told));
            const double u1 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
            nprop++;
            if (u1 * lb < l) {                       // :130
                if (l > lb) {                        // :132-136
                    if (v.adapt) { c *= v.factor; c100 = c / 100; }
                    else if (!(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
                }
                if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = -th; }
                nev++; nflip++;
                xf = xs; tf = s; th = -th;
                gth = gm;
            }
        }
        a = c + (gx - gmu) * th;                     // fact_samplers.jl:51
        b = c100 + th * gth;                         // fact_samplers.jl:52
        told = s;
        // queue_time! (:54-66): the earlier of the proposed reflection and the hitting time of 0
        const double ur = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
        const double trefl = (v.sticky & ZZ_STICKY_ZZ) ? zz_poisson_time3(a, b, 0.01, ur) : zz_poisson_time(a, b, ur);   // stickyzz.jl:144-147
        const double tfreeze = zz_freezing_time(xnow, th);
        if (tfreeze <= trefl) { fbit = true; tau = s + tfreeze; }
        else { fbit = false; tau = s + trefl; }
    }
    o.a = frozen ? thf : a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k | (fbit ? ZZ_RENEW_BIT : 0u); o.nprop = nprop; o.nflip = nev; o.flags = flags;
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1;
    (void)nflip;
}

// ---------------------------------------------------------------------------------------------------------------
// Factorised Boomerang in spdmp (F::FactBoomerang, src/sfact.jl:29-48,73-145) seen from coordinate j.  Between events a
// coordinate rotates around mu_j (zz_boom_at); its record holds the anchor (tf, xf, theta at tf).  Own items: proposal
// (rate fact_samplers.jl:37-39, constant bound :58-65) and velocity refreshment (sfact.jl:100-108; one clock of rate
// lambda_ref/d per coordinate -- the superposition of the reference's single clock, see oracle/zz_oracle.c).  Neighbour
// items: any recorded event of a trigger neighbour (reflection or refreshment) reschedules j (:110-114,131-135).
// Record layout in this mode: priv/spec = (a, next refreshment time, next proposal time, c); tau = the earlier of the two.
template <int NB>
ZZ_HD void zz_boom_eval(const ZzHood<NB>& hd, const ZzHoodMu<NB>& hm, bool same, double s, double xown, double thown, double& gt,
                        double& zsum)
{
    double at = 0.0, zs = 0.0;
    bool first = true;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        if (m < hd.n) {
            double x, th;
            if (m == hd.self) { x = xown; th = thown; }
            else zz_boom_at(hd.tf[m], hd.xf[m], hd.th[m], hm.mu[m], s, &x, &th);
            if (hd.fl[m] & ZZ_NB_BND) {                          // sum((x_j - mu_j)^2 + theta_j^2 for j in nhd), left fold
                const double term = (x - hm.mu[m]) * (x - hm.mu[m]) + th * th;
                zs = first ? term : zs + term;
                first = false;
            }
            if (same) { if (hd.fl[m] & ZZ_NB_BND) at += hd.wb[m] * x; }
            else if (hd.fl[m] & ZZ_NB_TGT) at += hd.wt[m] * x;   // idot(Gamma, j, x), common.jl:16-24
        }
    }
    gt = at; zsum = zs;
}

ZZ_HD double zz_boom_a(double c, double xi, double thi, double zsum, double diag)
{   // ab(G, i, x, theta, c, Z::FactBoomerang), fact_samplers.jl:58-65 (b = 0)
    const double z = zz_sqrt(zsum);
    const double z2 = xi * xi + thi * thi;
    return c * zz_sqrt(z2) * z + z2 * diag;
}

template <int NB>
ZZ_HD void zz_timeline_boom(ZzHood<NB>& hd, const ZzHoodMu<NB>& hm, const ZzPool& pool, const ZzOwn& w, const ZzGraph& g, const ZzView& v,
                            int32_t j, double H, int incl, uint32_t flags0, ZzNodeOut& o)
{
    double th = w.th, tf = w.tf, xf = w.xf;               // own anchor
    double a = w.a, tref = w.b, tau = w.told, c = w.c;
    uint32_t k = w.k;
    const double muj = v.bmu[j], sigj = v.bsig[j];
    const bool has_h = (!g.same && g.h);
    const double hj = has_h ? g.h[j] : 0.0;
    double diag = 0.0;                                     // Z.Gamma[j,j] (0 when not stored)
#pragma unroll
    for (int m = 0; m < NB; ++m) if (m == hd.self && (hd.fl[m] & ZZ_NB_BND)) diag = hd.wb[m];
    uint32_t nprop = 0, nev = 0, nrefl = 0, flags = flags0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;

    for (int item = 0;; ++item) {
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int nm = p < pool.n ? pool.m[p] : 0x7fffffff;
        const bool isprop = tau <= tref;
        const double town = isprop ? tau : tref;
        const bool own = (town < nt) || (town == nt && hd.self < nm);
        const double s = own ? town : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        double xs, ths;
        zz_boom_at(tf, xf, th, muj, s, &xs, &ths);
        bool propose = false;
        if (!own) {
            const double tha = pool.th[p];
            bool trig = false;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == nm) {                             // re-anchor the neighbour at its event
                    double xm, tm;
                    zz_boom_at(hd.tf[m], hd.xf[m], hd.th[m], hm.mu[m], s, &xm, &tm);
                    hd.xf[m] = xm; hd.tf[m] = s; hd.th[m] = tha;
                    trig = (hd.fl[m] & ZZ_NB_TRIG) != 0;
                }
            }
            ++p;
            if (!trig) continue;
        } else if (!isprop) {                              // refreshment, sfact.jl:100-108
            const double u1 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k);
            const double u2 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k + 1u);
            const double u3 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k + 2u);
            k += 3u;
            const double thn = v.brho * ths + v.brhobar * sigj * zz_randn(u1, u2);
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = thn; }
            nev++;
            xf = xs; tf = s; th = thn; ths = thn;
            tref = s - zz_log(u3) / v.bref_rate;
        } else {
            propose = true;
        }
        double gt, zsum;
        zz_boom_eval<NB>(hd, hm, g.same != 0, s, xs, ths, gt, zsum);
        if (propose) {
            if (has_h) gt = gt - hj;
            const double l = zz_pos((gt - (xs - muj) * diag) * ths);      // fact_samplers.jl:37-39
            const double lb = zz_pos(a);                                   // sfact.jl:70 with b = 0
            const double u1 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
            nprop++;
            if (u1 * lb < l) {                                             // sfact.jl:121
                if (l >= lb) {
                    if (v.adapt) c *= v.factor;
                    else if (!(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
                }
                if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = -ths; }
                nev++; nrefl++;
                xf = xs; tf = s; th = -ths; ths = -ths;                    // dynamics.jl:46-49
            }
        }
        a = zz_boom_a(c, xs, ths, zsum, diag);
        tau = s + zz_poisson_time_L(a, 0.0, zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, k++)));   // sfact.jl:134,139
    }
    o.a = a; o.b = tref; o.told = tau; o.tau = tau <= tref ? tau : tref; o.c = c;
    o.k = k; o.nprop = nprop; o.nflip = nev; o.flags = flags | (nrefl << 3);
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1;
}

// initial bound, first proposal and first refreshment time (sfact.jl:184-190: neither is offset by t0)
template <int NB>
ZZ_HD void zz_boom_init(ZzHood<NB>& hd, const ZzHoodMu<NB>& hm, const ZzGraph& g, const ZzView& v, int32_t j, double t0)
{
    double th, tf, xf; uint32_t h0, h1;
    zz_ld_kin(v.kin + j, th, tf, xf, h0, h1);
    double xs, ths;
    zz_boom_at(tf, xf, th, v.bmu[j], t0, &xs, &ths);
    double diag = 0.0;
#pragma unroll
    for (int m = 0; m < NB; ++m) if (m == hd.self && (hd.fl[m] & ZZ_NB_BND)) diag = hd.wb[m];
    double gt, zsum;
    zz_boom_eval<NB>(hd, hm, g.same != 0, t0, xs, ths, gt, zsum);
    ZzPriv pr;
    pr.c = zz_ld_priv(v.priv + j).c;
    pr.a = zz_boom_a(pr.c, xs, ths, zsum, diag);
    const double tau = zz_poisson_time_L(pr.a, 0.0, zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, 0)));
    const double tref = -zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, 1)) / v.bref_rate;
    pr.b = tref; pr.told = tau;
    v.priv[j] = pr;
    v.tau[j] = tau <= tref ? tau : tref;
    v.kctr[j] = 2u;
}

// ---------------------------------------------------------------------------------------------------------------
// ZigZag with velocity refreshments (Z.lambdaref > 0: src/sfact.jl:78-114,188-190) seen from coordinate j.  Own items: proposal
// (as zz_timeline) and refreshment -- theta_j <- sigma_j * (+-1) (:100-101), next refreshment after Exp(lambdaref / d), then the
// reschedule of :109-113.  Neighbour items: any recorded event of a trigger neighbour (reflection or refreshment; the lists carry
// the velocity after).  Contract: zzo_spdmp_refresh in mode ctr | lazy -- per coordinate the draws are: first proposal (k = 0),
// first refreshment time (k = 1); a proposal consumes (thinning, reschedule); a neighbour's event (reschedule); a refreshment
// (sign, next refreshment time, reschedule).
template <int NB>
ZZ_HD void zz_timeline_refresh(ZzHood<NB>& hd, const ZzPool& pool, const ZzOwn& w, const ZzGraph& g, const ZzView& v,
                               int32_t j, double H, int incl, uint32_t flags0, ZzNodeOut& o)
{
    double th = w.th, tf = w.tf, xf = w.xf;
    double a = w.a, b = w.b, told = w.told, c = w.c;
    double c100 = c / 100;
    double tauP = zz_ld(v.rst + 2 * (size_t)j), tref = zz_ld(v.rst + 2 * (size_t)j + 1);
    uint32_t k = w.k;
    const double gmu = g.grid_m ? 0.0 : g.gmu[j];
    const bool has_h = (!g.same && g.h);
    const double hj = has_h ? g.h[j] : 0.0;
    const double sigj = v.rsig[j];
    uint32_t nprop = 0, nev = 0, nrefl = 0, flags = flags0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;
    for (int item = 0;; ++item) {
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int nm = p < pool.n ? pool.m[p] : 0x7fffffff;
        const bool isprop = tauP <= tref;
        const double town = isprop ? tauP : tref;
        const bool own = (town < nt) || (town == nt && hd.self < nm);
        const double s = own ? town : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        if (!own) {
            const double tha = pool.th[p];
            bool trig = false;
#pragma unroll
            for (int m = 0; m < NB; ++m) {
                if (m == nm) {
                    hd.xf[m] = hd.xf[m] + hd.th[m] * (s - hd.tf[m]);
                    hd.tf[m] = s;
                    hd.th[m] = tha;
                    trig = (hd.fl[m] & ZZ_NB_TRIG) != 0;
                }
            }
            ++p;
            if (!trig) continue;
        }
        const double xs = xf + th * (s - tf);
        if (own && !isprop) {                              // refreshment, sfact.jl:100-108
            const double u1 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k);
            const double u2 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k + 1u);
            k += 2u;
            const double thn = sigj * (u1 < 0.5 ? -1.0 : 1.0);
            if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = thn; }
            nev++;
            xf = xs; tf = s; th = thn;
            tref = s - zz_log(u2) / v.rlam1;
        }
        double gt, gx, gp, gm;
        zz_eval_hood<NB>(hd, g.same != 0, s, xs, th, gt, gx, gp, gm);
        double gth = gp;
        if (own && isprop) {
            if (has_h) gt = gt - hj;
            const double l = zz_pos(gt * th);                 // fact_samplers.jl:28-30
            const double lb = zz_pos(a + b * (s - told));     // sfact.jl:70
            const double u1 = zz_u01(v.seed0, v.seed1, (uint64_t)j, k++);
            nprop++;
            if (u1 * lb < l) {                                // sfact.jl:121
                if (l >= lb) {
                    if (v.adapt) { c *= v.factor; c100 = c / 100; }
                    else if (!(flags & ZZ_F_VIOL)) { flags |= ZZ_F_VIOL; o.viol_t = s; o.viol_l = l; o.viol_lb = lb; }
                }
                if (nev == ZZ_MAXFLIP) { flags |= ZZ_F_OVERFLOW; break; }
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m) if (m == (int)nev) { o.fl[m] = s; o.fth[m] = -th; }
                nev++; nrefl++;
                xf = xs; tf = s; th = -th;                    // dynamics.jl:46-49
                gth = gm;
            }
        }
        a = c + (gx - gmu) * th;                              // fact_samplers.jl:51
        b = c100 + th * gth;                                  // fact_samplers.jl:52
        told = s;
        tauP = s + zz_poisson_time(a, b, zz_u01(v.seed0, v.seed1, (uint64_t)j, k++));   // sfact.jl:112,134,139
    }
    o.a = a; o.b = b; o.told = told; o.tau = tauP <= tref ? tauP : tref; o.c = c;
    o.tprop = tauP; o.tref = tref;
    o.k = k; o.nprop = nprop; o.nflip = nev; o.flags = flags | (nrefl << 3);
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1;
}

// ---------------------------------------------------------------------------------------------------------------
// Lattice interior (plain ZigZag): the 5-point column {j-M, j-1, j, j+1, j+M} with weights {-1, -1, diag, -1, -1},
// target == sampler matrix, no linear term, mu = 0.  Same arithmetic as zz_gather_grid + zz_timeline<5, false>, operation
// for operation -- (-1) * x is written -x, gx - 0 is written gx, both exact -- with every flag test, bounds test and weight
// load folded away: the common case costs about a third fewer instructions, which is what a lone warp in a late
// relaxation pass is made of.  Neighbour positions follow the storage order of the column: 0, 1, (2 = self), 3, 4.
template <bool MG>
ZZ_HD void zz_process_interior(const ZzGraph& g, const ZzView& v, int32_t j, double H, int incl, uint32_t w0, uint32_t cur,
                               bool first_iter, ZzNodeOut& o)
{
    ZzOwn w;
    zz_load_own(v, j, w);
    const int32_t M = g.grid_m;
    const int32_t idx[5] = { j - M, j - 1, j, j + 1, j + M };
    double nth[5], ntf[5], nxf[5]; uint32_t h0[5], h1[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        nth[m] = 0.0; ntf[m] = 0.0; nxf[m] = 0.0; h0[m] = 0; h1[m] = 0;
        if (m != 2) zz_ld_kin(zz_kin_at<MG>(v, idx[m]), nth[m], ntf[m], nxf[m], h0[m], h1[m]);
    }
    ZzPool pool; pool.n = 0;
    uint32_t flags = 0;
    if (!first_iter) zz_gather_flips<5, MG, false>(v, idx, h0, h1, 5, 2, w0, cur, pool, flags);

    const double diag = g.grid_diag[4];
    double th = w.th, tf = w.tf, xf = w.xf;
    double a = w.a, b = w.b, told = w.told, c = w.c;
    double c100 = c / 100;
    double tau = w.tau;
    uint32_t k = w.k;
    uint32_t nprop = 0, nflip = 0, nitems = 0;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;
    int p = 0;
    for (int item = 0;; ++item) {
        ++nitems;
        const double nt = p < pool.n ? pool.t[p] : ZZ_INF;
        const int nm = p < pool.n ? pool.m[p] : 0x7fffffff;
        const bool own = (tau < nt) || (tau == nt && 2 < nm);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        if (!own) {   // the neighbour at position nm flips at s (every lattice neighbour is a trigger)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                if (m != 2 && m == nm) {
                    nxf[m] = nxf[m] + nth[m] * (s - ntf[m]);
                    ntf[m] = s;
                    nth[m] = -nth[m];
                }
            }
            ++p;
        }
        const uint32_t kr = own ? k + 1u : k;
        const double L2 = zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, kr));
        const double u1 = own ? zz_u01(v.seed0, v.seed1, (uint64_t)j, k) : 0.0;
        k = kr + 1u;
        const double xs = xf + th * (s - tf);
        // idot over the column in storage order (common.jl:16-24): positions, velocities with +own / -own
        double ax = 0.0, ap = 0.0, am = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            if (m == 2) {
                ax += diag * xs; ap += diag * th; am += diag * (-th);
            } else {
                const double x = nxf[m] + nth[m] * (s - ntf[m]);
                ax += -x; ap += -nth[m]; am += -nth[m];
            }
        }
        double gth = ap;
        if (own) {
            const double l = zz_pos(ax * th);                 // fact_samplers.jl:28-30
            const double lb = zz_pos(a + b * (s - told));     // sfact.jl:70
            nprop++;
            if (u1 * lb < l) {                                // sfact.jl:121
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
        a = c + ax * th;                                      // fact_samplers.jl:51 (mu = 0)
        b = c100 + th * gth;                                  // fact_samplers.jl:52
        told = s;
        tau = s + zz_poisson_time_L(a, b, L2);                // sfact.jl:134,139
    }
    o.a = a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k; o.nprop = nprop; o.nflip = nflip; o.flags = flags;
    o.hdr0 = w.hdr0; o.hdr1 = w.hdr1; o.nitems = nitems; o.interior = 1u;
}

// Entry points.  KIND 0: 5-point lattice (index arithmetic); KIND 1: general sparse columns.
#define ZZ_KIND_GRID 0
#define ZZ_KIND_CSR 1
#define ZZ_MODE_PLAIN 0
#define ZZ_MODE_LB 1       // LocalBound (src/local.jl)
#define ZZ_MODE_STICKY 2   // sticky ZigZag (src/ss_fact.jl)
#define ZZ_MODE_BOOM 3     // factorised Boomerang (F::FactBoomerang in src/sfact.jl)
#define ZZ_MODE_LOGIT 4    // plain ZigZag with the subsampled logistic target (zz_logit.h; general sparse kernels only)
#define ZZ_MODE_STRONG 5   // strong-bound sparse sticky ZigZag (zz_strong.h); only in -DZZ_ENABLE_STRONG builds of the image
#define ZZ_MODE_REFRESH 6  // ZigZag with velocity refreshments (Z.lambdaref > 0, src/sfact.jl:78-114)
#define ZZ_MODE_HAS_VEL(M) ((M) == ZZ_MODE_STICKY || (M) == ZZ_MODE_BOOM || (M) == ZZ_MODE_STRONG || (M) == ZZ_MODE_REFRESH)   // flip lists carry the velocity after each event
// Columns of ZZ_NB + 1 .. ZZ_NB_WIDE entries for the samplers that have no list-walking fallback (sticky, Boomerang,
// refreshments): the same gather + timeline with a wider neighbourhood.  Out of line (its arrays live in local memory): the
// register allocation of the common path is untouched, long columns are merely slower.
template <int MODE, bool MG>
ZZ_HD_NOINLINE void zz_process_node_wide(const ZzGraph& g, const ZzView& v, int32_t j, double H, int incl, uint32_t w0,
                                         uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    ZzPool pool; uint32_t flags = 0;
    ZzOwn w;
    ZzHood<ZZ_NB_WIDE> hd;
    zz_load_own(v, j, w);
    if (MODE == ZZ_MODE_BOOM) {
        ZzHoodMu<ZZ_NB_WIDE> hm;
        zz_gather_csr<ZZ_NB_WIDE, MG, true>(g, v, j, w0, cur, first_iter, hd, pool, flags, &hm);
        zz_timeline_boom<ZZ_NB_WIDE>(hd, hm, pool, w, g, v, j, H, incl, flags, o);
        return;
    }
    zz_gather_csr<ZZ_NB_WIDE, MG, true>(g, v, j, w0, cur, first_iter, hd, pool, flags);
    if (MODE == ZZ_MODE_REFRESH) zz_timeline_refresh<ZZ_NB_WIDE>(hd, pool, w, g, v, j, H, incl, flags, o);
    else zz_timeline_sticky<ZZ_NB_WIDE>(hd, pool, w, g, v, j, H, incl, flags, o);
}

template <int KIND, int MODE, bool MG = true>
ZZ_HD void zz_process_node_k(const ZzGraph& g, const ZzView& v, int32_t j, double H, int incl, uint32_t w0,
                             uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    ZzPool pool; uint32_t flags = 0;
    ZzOwn w;
    o.interior = 0u;
    if (KIND == ZZ_KIND_GRID) {
        if (MODE == ZZ_MODE_PLAIN) {   // lattice interior: the specialised evaluation
            const int32_t M = g.grid_m, N = g.grid_n;
            const int32_t col = zz_grid_col(g, j), row = j - col * M;
            if (col > 0 && col < N - 1 && row > 0 && row < M - 1) { zz_process_interior<MG>(g, v, j, H, incl, w0, cur, first_iter, o); return; }
        }
        ZzHood<5> hd;
        ZZ_SEG(0);
        zz_load_own(v, j, w);
        if (MODE == ZZ_MODE_BOOM) {
            ZzHoodMu<5> hm;
            zz_gather_grid<MG, true>(g, v, j, w0, cur, first_iter, hd, pool, flags, &hm);
            zz_timeline_boom<5>(hd, hm, pool, w, g, v, j, H, incl, flags, o);
            return;
        }
        zz_gather_grid<MG, MODE == ZZ_MODE_STICKY || MODE == ZZ_MODE_REFRESH>(g, v, j, w0, cur, first_iter, hd, pool, flags);
        ZZ_SEG(2);
        if (MODE == ZZ_MODE_REFRESH) zz_timeline_refresh<5>(hd, pool, w, g, v, j, H, incl, flags, o);
        else if (MODE == ZZ_MODE_STICKY) zz_timeline_sticky<5>(hd, pool, w, g, v, j, H, incl, flags, o);
        else zz_timeline<5, MODE == ZZ_MODE_LB>(hd, pool, w, g, v, j, H, incl, flags, o);
        ZZ_SEG(5);
        return;
    }
    ZzHood<ZZ_NB> hd;
    if (g.nptr[j + 1] - g.nptr[j] <= ZZ_NB) {
        zz_load_own(v, j, w);
        if (MODE == ZZ_MODE_BOOM) {
            ZzHoodMu<ZZ_NB> hm;
            zz_gather_csr<ZZ_NB, MG, true>(g, v, j, w0, cur, first_iter, hd, pool, flags, &hm);
            zz_timeline_boom<ZZ_NB>(hd, hm, pool, w, g, v, j, H, incl, flags, o);
            return;
        }
        zz_gather_csr<ZZ_NB, MG, MODE == ZZ_MODE_STICKY || MODE == ZZ_MODE_REFRESH>(g, v, j, w0, cur, first_iter, hd, pool, flags);
        if (MODE == ZZ_MODE_REFRESH) zz_timeline_refresh<ZZ_NB>(hd, pool, w, g, v, j, H, incl, flags, o);
        else if (MODE == ZZ_MODE_STICKY) zz_timeline_sticky<ZZ_NB>(hd, pool, w, g, v, j, H, incl, flags, o);
        else zz_timeline<ZZ_NB, MODE == ZZ_MODE_LB>(hd, pool, w, g, v, j, H, incl, flags, o);
        return;
    }
    if (MODE == ZZ_MODE_STICKY || MODE == ZZ_MODE_BOOM || MODE == ZZ_MODE_REFRESH) {   // (the host refuses columns over ZZ_NB_WIDE)
        o.interior = 0u;
        zz_process_node_wide<(MODE == ZZ_MODE_STICKY || MODE == ZZ_MODE_BOOM || MODE == ZZ_MODE_REFRESH) ? MODE : ZZ_MODE_STICKY, MG>(
            g, v, j, H, incl, w0, cur, first_iter, o);
        return;
    }
    zz_process_node_slow(g, v, j, H, incl, w0, cur, first_iter, o);
}

// used by the host-side schedule emulation
ZZ_HD void zz_process_node(const ZzGraph& g, const ZzView& v, int32_t j, double H, int incl, uint32_t w0,
                           uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    if (v.refresh) {
        if (g.grid_m) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_REFRESH>(g, v, j, H, incl, w0, cur, first_iter, o);
        else zz_process_node_k<ZZ_KIND_CSR, ZZ_MODE_REFRESH>(g, v, j, H, incl, w0, cur, first_iter, o);
    } else if (v.boom) {
        if (g.grid_m) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_BOOM>(g, v, j, H, incl, w0, cur, first_iter, o);
        else zz_process_node_k<ZZ_KIND_CSR, ZZ_MODE_BOOM>(g, v, j, H, incl, w0, cur, first_iter, o);
    } else if (v.sticky) {
        if (g.grid_m) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_STICKY>(g, v, j, H, incl, w0, cur, first_iter, o);
        else zz_process_node_k<ZZ_KIND_CSR, ZZ_MODE_STICKY>(g, v, j, H, incl, w0, cur, first_iter, o);
    } else if (v.local_bound) {
        if (g.grid_m) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_LB>(g, v, j, H, incl, w0, cur, first_iter, o);
        else zz_process_node_k<ZZ_KIND_CSR, ZZ_MODE_LB>(g, v, j, H, incl, w0, cur, first_iter, o);
    } else {
        if (g.grid_m && v.nranks <= 1) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_PLAIN, false>(g, v, j, H, incl, w0, cur, first_iter, o);   // as the single-GPU kernel
        else if (g.grid_m) zz_process_node_k<ZZ_KIND_GRID, ZZ_MODE_PLAIN>(g, v, j, H, incl, w0, cur, first_iter, o);
        else zz_process_node_k<ZZ_KIND_CSR, ZZ_MODE_PLAIN>(g, v, j, H, incl, w0, cur, first_iter, o);
    }
}

#endif  // ZZ_FAST_H
