// zz_core.h -- per-coordinate timeline evaluation of the windowed local-ZigZag scheme.
//
// Compiles for the device (zz_kernels.cu) and for the host (oracle/zz_window_sim.cpp, a TEST-ONLY
// emulation of the schedule used to debug the scheme without a GPU).  DESIGN.md explains the scheme;
// in short, for one time window [F, H):
//   * every coordinate j owns a "timeline": its own proposals (src/sfact.jl:77-140, restricted to i == j)
//     interleaved with the reschedules imposed by accepted flips of its neighbours (sfact.jl:131-135,
//     seen from the receiving side j instead of the firing side i);
//   * the timeline of j is a pure function of j's frontier state and of the flip lists of its neighbours,
//     so all timelines are relaxed Jacobi-style until no flip list changes; the fixed point is exactly the
//     sequential event history (causality is strictly forward in (time, coordinate) order);
//   * positions are flip-anchored ("lazy"): x_k(s) = xf_k + theta_k (s - tf_k), rewritten only by the owner.
#ifndef ZZ_CORE_H
#define ZZ_CORE_H

#include "zz_math.h"

#define ZZ_MAXFLIP 6     // accepted flips one coordinate may record per window (4-bit count field)
#define ZZ_MAXITEMS 48   // timeline items (proposals + reschedules) per coordinate per window
#define ZZ_MAXRANKS 8     // GPUs of one node
#define ZZ_TAG_LIMIT 0x0f000000u  // iteration tags are rebased before they reach 2^28

// neighbour entry flags
#define ZZ_NB_TGT 1u   // entry of the target precision column (enters grad phi_j)
#define ZZ_NB_BND 2u   // entry of the sampler's Z.Gamma column (enters the bound a_j, b_j)
#define ZZ_NB_TRIG 4u  // a flip of this neighbour reschedules j (j in G1[k], sfact.jl:131)

// per-node result flags
#define ZZ_F_OVERFLOW 1u  // more than ZZ_MAXFLIP flips / ZZ_MAXITEMS items: the window must be shortened
#define ZZ_F_VIOL 2u      // accepted with l >= lb and adapt == false  (sfact.jl:123-124)
#define ZZ_F_STICKY_ERR 4u  // a freezing coordinate was not at 0 (ss_fact.jl:89-91)
#define ZZ_STICKY_REVERSIBLE 2   // bits of ZzView::sticky (bit 0: sticky sampler): sspdmp options reversible / strong_upperbounds
#define ZZ_STICKY_STRONG_UB 4
#define ZZ_STICKY_ZZ 8            // the dense sticky sampler stickyzz / sspdmp2 (src/stickyzz.jl): proposal times at rate 0.01 + (a + b t)^+,
                                 // coordinates that start at 0 start frozen
#define ZZ_RENEW_BIT 0x80000000u  // in the draw counter: the queued time is a bound expiry, not a proposal (local.jl:34)

// Kinematic record of one coordinate, read by its neighbours: 32 B = one DRAM/L2 sector.
// hdr[s] = (tag << 4) | count describes flip-list slot s: `count` flips recorded by the relaxation
// iteration `tag`.  A reader running iteration `cur` of a window whose first tag is `w0` uses the slot
// with the largest tag in [w0, cur) -- never the slot the owner may be rewriting during `cur`.
struct __attribute__((aligned(32))) ZzKin {
    double theta, tf, xf;
    uint32_t hdr[2];
};

// Private bound record of one coordinate (only its owner reads or writes it).
struct __attribute__((aligned(32))) ZzPriv {
    double a, b, told, c;
};

// Speculative end-of-window state written by every relaxation pass, folded into the frontier at commit.
struct __attribute__((aligned(16))) ZzSpec {
    double a, b, told, tau, c;
    uint32_t k;       // draw counter after the window
    uint16_t nprop;   // proposals inside the window
    uint8_t nflip;    // accepted ones
    uint8_t flags;
};

struct ZzGraph {
    const int32_t* nptr;   // [d+1] neighbour list offsets
    const int32_t* nidx;   // neighbour ids (0-based), ascending, self included (storage order of column j)
    const double* nwt;     // target weight Gamma_t[k,j]
    const double* nwb;     // bound weight  Gamma_b[k,j]
    const uint8_t* nfl;    // ZZ_NB_* flags
    const double* gmu;     // idot(Z.Gamma, j, Z.mu)   (fact_samplers.jl:51)
    const double* h;       // linear term of the target, may be null
    int32_t same;          // target and bound matrices are the same object and h == 0, mu == 0
    // 5-point lattice fast path (0.01 I + gridlaplacian(m, n), scripts/gridlaplace.jl): neighbours and weights
    // follow from index arithmetic, no index/weight loads.  grid_m == 0 -> general CSR lists above.
    int32_t grid_m, grid_n;
    double grid_diag[5];   // diagonal entry by node degree (2, 3, 4), exactly as stored in the matrix
    // j / grid_m for 0 <= j < 2^31 as one multiply and shift (zz_grid_set_magic / zz_grid_col)
    uint64_t grid_magic;
    int32_t grid_shift, grid_pad;
};

// Lattice column of coordinate j: floor(j / M) = (j * magic) >> shift with magic = floor(2^shift / M) + 1 and
// shift = 31 + ceil(log2 M); exact for every 0 <= j < 2^31 (round-up method of Granlund & Montgomery; tested exhaustively
// against the division for the lattice sizes of the test-suite).
ZZ_HD void zz_grid_set_magic(ZzGraph& g)
{
    g.grid_magic = 0; g.grid_shift = 0; g.grid_pad = 0;
    if (g.grid_m <= 0) return;
    int l = 0;
    while ((1LL << l) < (long long)g.grid_m) ++l;
    g.grid_shift = 31 + l;
    g.grid_magic = (uint64_t)((((unsigned __int128)1) << g.grid_shift) / (unsigned __int128)g.grid_m) + 1ULL;
}
ZZ_HD int32_t zz_grid_col(const ZzGraph& g, int32_t j)
{
    return (int32_t)(((uint64_t)(uint32_t)j * g.grid_magic) >> g.grid_shift);
}

struct ZzView {
    int32_t d;
    ZzKin* kin;
    double* flips;   // [d][2][ZZ_MAXFLIP]
    ZzPriv* priv;
    double* tau;
    uint32_t* kctr;
    uint64_t seed0, seed1;
    int32_t adapt;
    double factor;
    // LocalBound variant (src/local.jl): bounds from the target's own derivatives, valid for Delta = 2/c/|theta|, then
    // renewed; the renew flag of a coordinate travels in bit 31 of its draw counter
    int32_t local_bound;
    // Sticky ZigZag (src/ss_fact.jl): coordinates freeze at 0 and thaw after Exp(kappa); the lists of a window then hold
    // (time, velocity after the event) pairs -- fth is the velocity array parallel to `flips`; a frozen coordinate has
    // theta == 0 in its record and keeps its saved velocity in priv.a; bit 31 of the draw counter = "next own event is a
    // freeze" (the f flag of ss_fact.jl:54-66)
    int32_t sticky;
    double* fth;
    const double* kappa;
    // coordinate sharding across GPUs (one process per GPU): rank r owns the global ids [r*shard, (r+1)*shard).
    // Every rank allocates full-length arrays and indexes them globally; a record is valid only in its owner's
    // copy, reached through the peer mappings below (NVLink loads).  nranks == 1: the plain pointers above.
    int32_t nranks, rank, shard;
    int32_t lo, hi;                 // owned range
    ZzKin* kin_peer[ZZ_MAXRANKS];
    double* flips_peer[ZZ_MAXRANKS];
    double* fth_peer[ZZ_MAXRANKS];   // (samplers whose lists carry the velocity after each event: sticky, Boomerang)
    // Factorised Boomerang (F::FactBoomerang in src/sfact.jl): rotation around bmu, velocity refreshment
    // theta <- brho theta + brhobar bsig N(0,1) at rate bref_rate = lambda_ref / d per coordinate; lists carry
    // (time, velocity after) like the sticky ones (fth)
    // ZigZag with velocity refreshments (hasrefresh(Z): Z.lambdaref > 0, src/sfact.jl:78-114,188-190; mode ZZ_MODE_REFRESH): one
    // refreshment clock of rate rlam1 = lambdaref / d per coordinate (superposition of the reference's single clock), refreshed
    // velocity = rsig[j] * (+-1); lists carry the velocity after each event (fth).  rst[j] = (next proposal time, next
    // refreshment time) of the frontier -- tau[j] is the earlier of the two -- and rspec[j] the speculative end-of-window pair.
    int32_t refresh, pad_refresh;
    const double* rsig;
    double rlam1;
    double* rst;      // [d][2]
    double* rspec;    // [d][2]
    int32_t boom, pad_boom;
    const double* bmu;
    const double* bsig;
    double bref_rate, brho, brhobar;
};

// where the authoritative record / flip lists of coordinate k live
// (MG = false: single-GPU kernels, the test is compiled away)
template <bool MG = true>
ZZ_HD const ZzKin* zz_kin_at(const ZzView& v, int32_t k)
{
    if (MG && v.nranks > 1) return v.kin_peer[k / v.shard] + k;
    return v.kin + k;
}
template <bool MG = true>
ZZ_HD const double* zz_flips_at(const ZzView& v, int32_t k)
{
    if (MG && v.nranks > 1) return v.flips_peer[k / v.shard] + (size_t)k * 2 * ZZ_MAXFLIP;
    return v.flips + (size_t)k * 2 * ZZ_MAXFLIP;
}

struct ZzNodeOut {
    double a, b, told, tau, c;
    uint32_t k, nprop, nflip, flags;
    double fl[ZZ_MAXFLIP];
    double fth[ZZ_MAXFLIP];  // sticky only: velocity after each recorded event
    double viol_t, viol_l, viol_lb;
    uint32_t hdr0, hdr1;   // own flip-list headers as read at entry
    double tprop, tref;    // ZZ_MODE_REFRESH: next proposal / refreshment time after the window
    uint32_t nitems;       // timeline items processed by this evaluation (statistics of the host emulation)
    uint32_t interior;     // evaluated by the lattice-interior path: all four readers exist
};

#if defined(__CUDA_ARCH__)
// data that other SMs rewrite between grid barriers is always read through L2
ZZ_HD double zz_ld(const double* p) { return __ldcg(p); }
ZZ_HD void zz_ld_kin(const ZzKin* p, double& th, double& tf, double& xf, uint32_t& h0, uint32_t& h1)
{
    double2 u = __ldcg(reinterpret_cast<const double2*>(p));
    double2 w = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    th = u.x; tf = u.y; xf = w.x;
    unsigned long long hh = (unsigned long long)__double_as_longlong(w.y);
    h0 = (uint32_t)hh; h1 = (uint32_t)(hh >> 32);
}
ZZ_HD uint32_t zz_ld32(const uint32_t* p) { return __ldcg(p); }
ZZ_HD ZzPriv zz_ld_priv(const ZzPriv* p)
{
    double2 u = __ldcg(reinterpret_cast<const double2*>(p));
    double2 w = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    ZzPriv r; r.a = u.x; r.b = u.y; r.told = w.x; r.c = w.y;
    return r;
}
#else
ZZ_HD double zz_ld(const double* p) { return *p; }
ZZ_HD uint32_t zz_ld32(const uint32_t* p) { return *p; }
ZZ_HD ZzPriv zz_ld_priv(const ZzPriv* p) { return *p; }
ZZ_HD void zz_ld_kin(const ZzKin* p, double& th, double& tf, double& xf, uint32_t& h0, uint32_t& h1)
{
    th = p->theta; tf = p->tf; xf = p->xf; h0 = p->hdr[0]; h1 = p->hdr[1];
}
#endif

// Which flip-list slot is valid for a reader at iteration `cur` of a window starting at tag w0?
// Returns the slot in `slot` and its count; count 0 when neither slot belongs to this window.
ZZ_HD uint32_t zz_pick_slot(uint32_t h0, uint32_t h1, uint32_t w0, uint32_t cur, int& slot)
{
    uint32_t t0 = h0 >> 4, t1 = h1 >> 4;
    bool v0 = (t0 >= w0) && (t0 < cur), v1 = (t1 >= w0) && (t1 < cur);
    if (v0 && (!v1 || t0 > t1)) { slot = 0; return h0 & 15u; }
    if (v1) { slot = 1; return h1 & 15u; }
    slot = -1;
    return 0;
}

// Position and velocity of neighbour k at event key (s, key_idx): frontier anchor advanced through those
// of k's recorded flips that precede the key in (time, coordinate) order.
ZZ_HD void zz_nb_state(const ZzView& v, int32_t k, double s, int32_t key_idx, uint32_t w0, uint32_t cur,
                       double& x, double& th)
{
    double tf, xf; uint32_t h0, h1;
    zz_ld_kin(zz_kin_at(v, k), th, tf, xf, h0, h1);
    int slot;
    uint32_t cnt = zz_pick_slot(h0, h1, w0, cur, slot);
    if (cnt) {
        const double* fl = zz_flips_at(v, k) + slot * ZZ_MAXFLIP;
        for (uint32_t m = 0; m < cnt; ++m) {
            double fs = zz_ld(fl + m);
            if (fs < s || (fs == s && k <= key_idx)) {
                xf = xf + th * (fs - tf);
                tf = fs;
                th = -th;
            } else break;
        }
    }
    x = xf + th * (s - tf);
}

// One pass over column j at event key (s, key_idx).  Own state is passed in (it lives in registers).
//   gt  = idot(Gamma_t, j, x(s)) [- h_j]                  (target partial derivative, common.jl:16-24)
//   gx  = idot(Z.Gamma, j, x(s))
//   gp/gm = idot(Z.Gamma, j, theta) with +own / -own velocity (the caller picks after thinning)
ZZ_HD void zz_eval(const ZzGraph& g, const ZzView& v, int32_t j, double s, int32_t key_idx, double xown,
                   double thown, uint32_t w0, uint32_t cur, double& gt, double& gx, double& gp, double& gm)
{
    double at = 0.0, ax = 0.0, ap = 0.0, am = 0.0;
    const int32_t e1 = g.nptr[j + 1];
    for (int32_t e = g.nptr[j]; e < e1; ++e) {
        int32_t k = g.nidx[e];
        uint32_t fl = g.nfl[e];
        double x, th;
        if (k == j) {
            x = xown; th = thown;
            if (fl & ZZ_NB_BND) {
                double wb = g.nwb[e];
                ax += wb * x; ap += wb * th; am += wb * (-th);
            }
            if (!g.same && (fl & ZZ_NB_TGT)) at += g.nwt[e] * x;
        } else {
            if (!(fl & (ZZ_NB_TGT | ZZ_NB_BND))) continue;
            zz_nb_state(v, k, s, key_idx, w0, cur, x, th);
            if (fl & ZZ_NB_BND) {
                double wb = g.nwb[e];
                ax += wb * x; ap += wb * th; am += wb * th;
            }
            if (!g.same && (fl & ZZ_NB_TGT)) at += g.nwt[e] * x;
        }
    }
    if (g.same) at = ax;
    else if (g.h) at = at - g.h[j];
    gt = at; gx = ax; gp = ap; gm = am;
}

// First flip of a trigger neighbour strictly after key (last_t, last_i); (+inf, INT_MAX) if none.
ZZ_HD void zz_next_trigger(const ZzGraph& g, const ZzView& v, int32_t j, double last_t, int32_t last_i,
                           uint32_t w0, uint32_t cur, double& nt, int32_t& ni)
{
    nt = ZZ_INF; ni = 0x7fffffff;
    const int32_t e1 = g.nptr[j + 1];
    for (int32_t e = g.nptr[j]; e < e1; ++e) {
        if (!(g.nfl[e] & ZZ_NB_TRIG)) continue;
        int32_t k = g.nidx[e];
        if (k == j) continue;
        const ZzKin* p = zz_kin_at(v, k);
#if defined(__CUDA_ARCH__)
        unsigned long long hh = (unsigned long long)__double_as_longlong(__ldcg(reinterpret_cast<const double*>(p) + 3));
        uint32_t h0 = (uint32_t)hh, h1 = (uint32_t)(hh >> 32);
#else
        uint32_t h0 = p->hdr[0], h1 = p->hdr[1];
#endif
        int slot;
        uint32_t cnt = zz_pick_slot(h0, h1, w0, cur, slot);
        if (!cnt) continue;
        const double* fl = zz_flips_at(v, k) + slot * ZZ_MAXFLIP;
        for (uint32_t m = 0; m < cnt; ++m) {
            double fs = zz_ld(fl + m);
            if (fs > last_t || (fs == last_t && k > last_i)) {
                if (fs < nt || (fs == nt && k < ni)) { nt = fs; ni = k; }
                break;
            }
        }
    }
}

// next_time(t, abc, z) of src/not_fact_samplers.jl:43-50 for the LocalBound variant: the proposal is capped at the
// validity horizon Delta = 2/c/|theta| of the bound (local.jl:5); plain `s + dt` otherwise.
ZZ_HD double zz_next_time(double s, double dt, double c, double th, bool lbm, bool& renew)
{
    if (!lbm) return s + dt;
    const double Delta = 2.0 / c / (th < 0.0 ? -th : th);
    if (dt > Delta) { renew = true; return s + Delta; }
    renew = false;
    return s + dt;
}

// The timeline of coordinate j inside the window ending at H (exclusive, or inclusive when incl != 0),
// starting from its frontier state.  See the file header; per item this is exactly the arithmetic of
// spdmp_inner! (sfact.jl:118-139) and ab (fact_samplers.jl:50-54) for coordinate j.
ZZ_HD void zz_process_node_slow(const ZzGraph& g, const ZzView& v, int32_t j, double H, int incl, uint32_t w0,
                           uint32_t cur, bool first_iter, ZzNodeOut& o)
{
    double th, tf, xf; uint32_t hh0, hh1;
    zz_ld_kin(v.kin + j, th, tf, xf, hh0, hh1);
    const ZzPriv pr = zz_ld_priv(v.priv + j);
    double a = pr.a, b = pr.b, told = pr.told, c = pr.c;
    double c100 = c / 100;
    double tau = zz_ld(v.tau + j);
    uint32_t k = zz_ld32(v.kctr + j);
    const bool lbm = v.local_bound != 0;
    bool renew = (k & ZZ_RENEW_BIT) != 0;
    k &= ~ZZ_RENEW_BIT;
    const double gmu = g.gmu[j];
    uint32_t nprop = 0, nflip = 0, flags = 0;
    double last_t = -ZZ_INF; int32_t last_i = -1;
    o.viol_t = 0.0; o.viol_l = 0.0; o.viol_lb = 0.0;

    for (int item = 0;; ++item) {
        double nt = ZZ_INF; int32_t ni = 0x7fffffff;
        if (!first_iter) zz_next_trigger(g, v, j, last_t, last_i, w0, cur, nt, ni);
        const bool own = (tau < nt) || (tau == nt && j < ni);
        const double s = own ? tau : nt;
        if (!(s < H || (incl && s == H))) break;
        if (item >= ZZ_MAXITEMS) { flags |= ZZ_F_OVERFLOW; break; }
        const double xs = xf + th * (s - tf);
        double gt, gx, gp, gm, gth;
        if (own && lbm && renew) {                            // local.jl:34-41: bound expired, renew it
            zz_eval(g, v, j, s, j, xs, th, w0, cur, gt, gx, gp, gm);
            gth = gp;
        } else if (own) {
            zz_eval(g, v, j, s, j, xs, th, w0, cur, gt, gx, gp, gm);
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
#if defined(__CUDA_ARCH__)
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m)  // keeps o.fl in registers (no dynamic indexing)
                    if (m == (int)nflip) o.fl[m] = s;
                nflip++;
#else
                o.fl[nflip++] = s;
#endif
                xf = xs; tf = s; th = -th;                    // dynamics.jl:46-49
                gth = gm;
            } else {
                gth = gp;
            }
        } else {
            last_t = nt; last_i = ni;
            zz_eval(g, v, j, s, ni, xs, th, w0, cur, gt, gx, gp, gm);
            gth = gp;
        }
        a = c + (lbm ? gt : gx - gmu) * th;                   // fact_samplers.jl:51 / local.jl:3
        b = c100 + th * gth;                                  // fact_samplers.jl:52 / local.jl:4 (c100 = c / 100)
        told = s;
        const double dt = zz_poisson_time(a, b, zz_u01(v.seed0, v.seed1, (uint64_t)j, k++));  // sfact.jl:134,139
        tau = zz_next_time(s, dt, c, th, lbm, renew);         // not_fact_samplers.jl:43-50 when LocalBound
    }
    o.a = a; o.b = b; o.told = told; o.tau = tau; o.c = c;
    o.k = k | (renew ? ZZ_RENEW_BIT : 0u); o.nprop = nprop; o.nflip = nflip; o.flags = flags;
    o.hdr0 = hh0; o.hdr1 = hh1;
}

#include "zz_fast.h"

// Initial bound and first proposal time of coordinate j (sfact.jl:184-187; note: no "+ t0").
// FactBoomerang: every anchor is at t0 and there are no lists yet -- gather and evaluate once (kept apart from zz_init_node
// so that the ZigZag initialisation kernel stays small)
ZZ_HD void zz_init_node_boom(const ZzGraph& g, const ZzView& v, int32_t j, double t0)
{
    ZzPool pool; uint32_t fl = 0;
    if (g.grid_m) { ZzHood<5> hd; ZzHoodMu<5> hm; zz_gather_grid<false, true>(g, v, j, 1u, 1u, true, hd, pool, fl, &hm); zz_boom_init<5>(hd, hm, g, v, j, t0); }
    else {
        ZzHood<ZZ_NB> hd; ZzHoodMu<ZZ_NB> hm;
        if (g.nptr[j + 1] - g.nptr[j] > ZZ_NB) {   // columns of 9 .. ZZ_NB_WIDE entries (longer ones are refused by the host)
            ZzHood<ZZ_NB_WIDE> hdw; ZzHoodMu<ZZ_NB_WIDE> hmw;
            if (!zz_gather_csr<ZZ_NB_WIDE, false, true>(g, v, j, 1u, 1u, true, hdw, pool, fl, &hmw)) return;
            zz_boom_init<ZZ_NB_WIDE>(hdw, hmw, g, v, j, t0);
            return;
        }
        if (!zz_gather_csr<ZZ_NB, false, true>(g, v, j, 1u, 1u, true, hd, pool, fl, &hm)) return;
        zz_boom_init<ZZ_NB>(hd, hm, g, v, j, t0);
    }
}

ZZ_HD void zz_init_node(const ZzGraph& g, const ZzView& v, int32_t j, double t0)
{
    double th, tf, xf; uint32_t h0, h1;
    zz_ld_kin(v.kin + j, th, tf, xf, h0, h1);
    double gt, gx, gp, gm;
    zz_eval(g, v, j, t0, j, xf + th * (t0 - tf), th, 1u, 1u, gt, gx, gp, gm);
    ZzPriv pr;
    const ZzPriv pr0 = zz_ld_priv(v.priv + j);
    pr.c = pr0.c;
    const bool lbm = v.local_bound != 0;
    const bool zzmode = v.sticky && (v.sticky & ZZ_STICKY_ZZ);   // stickyzz / sspdmp2
    pr.a = pr.c + (lbm ? gt : gx - g.gmu[j]) * th;
    pr.b = pr.c / 100 + th * gp;
    pr.told = t0;
    if (zzmode && th == 0.0) pr.a = pr0.a;   // starts frozen: zz_setup_kernel has left the velocity to continue with in the `a` slot
    v.priv[j] = pr;
    const double u0 = zz_u01(v.seed0, v.seed1, (uint64_t)j, 0);
    const double dt = zzmode ? zz_poisson_time3(pr.a, pr.b, 0.01, u0) : zz_poisson_time(pr.a, pr.b, u0);   // (floor: poissontime.jl:93-99)
    bool renew = false;
    if (zzmode && th == 0.0) {   // stickyzz.jl:198-206: thaw clock only
        v.tau[j] = t0 - zz_log(u0) / v.kappa[j];
    } else if (v.sticky) {   // ss_fact.jl:178-188: the earlier of the first proposal and the hitting time of 0, from t0
        const double x0 = xf + th * (t0 - tf);
        const double tfreez = (th * x0 >= 0.0) ? ZZ_INF : -x0 / th;
        renew = dt > tfreez;
        v.tau[j] = t0 + (renew ? tfreez : dt);
    } else {
        v.tau[j] = lbm ? zz_next_time(t0, dt, pr.c, th, true, renew) : dt;   // sfact.jl:186 has no "+ t0"; local.jl:122 has
    }
    v.kctr[j] = 1u | (renew ? ZZ_RENEW_BIT : 0u);
    if (v.refresh) {   // sfact.jl:188-190 (per-coordinate clock, contract zzo_spdmp_refresh in mode ctr): second draw of the stream, no "+ t0"
        const double tref = -zz_log(zz_u01(v.seed0, v.seed1, (uint64_t)j, 1)) / v.rlam1;
        v.rst[2 * (size_t)j] = dt; v.rst[2 * (size_t)j + 1] = tref;
        v.tau[j] = dt <= tref ? dt : tref;
        v.kctr[j] = 2u;
    }
}

#endif  // ZZ_CORE_H
