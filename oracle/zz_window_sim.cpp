// zz_window_sim.cpp -- single-threaded HOST EMULATION of the windowed relaxation schedule the sm_100a
// kernel runs (zigzagboomerang.jl_b200/csrc/zz_kernels.cu), built on the very same per-coordinate code
// (zz_core.h, zz_ctl.h, zz_host_graph.h).
//
// TEST INFRASTRUCTURE ONLY (lives under oracle/ for that reason): it exists so the scheme can be checked
// against the sequential oracle (zz_oracle.c, mode ctr|lazy) on a machine without a GPU.  It is not a
// fallback: the product library (csrc/zzb200.cpp) never links or calls it and fails loudly without CUDA.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../zigzagboomerang.jl_b200/csrc/zz_core.h"
#include "../zigzagboomerang.jl_b200/csrc/zz_ctl.h"
#include "../zigzagboomerang.jl_b200/csrc/zz_host_graph.h"
#include "../zigzagboomerang.jl_b200/csrc/zz_host_logit.h"
#include "../zigzagboomerang.jl_b200/csrc/zz_strong.h"

struct zzw_event { double t; int64_t i; double x; double th; };

struct zzw_run {
    int64_t d = 0;
    std::vector<zzw_event> ev;
    std::vector<int64_t> acc;
    int64_t num = 0;
    std::vector<double> t, x, th, c, s1, s2;
    int status = 0; int64_t err_i = 0; double err_t = 0, err_l = 0, err_lb = 0;
    // schedule statistics
    int64_t windows = 0, retries = 0, iters = 0, node_evals = 0, max_iters = 0;
    std::vector<int64_t> pass_hist = std::vector<int64_t>(64, 0);  // work-list size summed per pass index
    std::vector<int64_t> item_hist = std::vector<int64_t>(64, 0);  // timeline items processed, summed per pass index
    std::string msg;
};

static zzw_run* zzw_impl(int64_t d, const int64_t* tcp, const int64_t* trv, const double* tnz, const double* h,
                   const int64_t* bcp, const int64_t* brv, const double* bnz, const double* mu, double t0,
                   const double* x0, const double* th0, double T, const double* c_in, const uint64_t* seed,
                   int adapt, double factor, double delta0, double target_frac, uint32_t tag_limit, int local_bound, const double* kappa,
                   const double* boom_sigma, double boom_lambdaref, double boom_rho, const ZzHostLogit* hl, const ZzStrong* st = nullptr,
                   const double* refresh_sigma = nullptr, double refresh_lambdaref = 0.0);

// Which schedule the emulation runs: 0 = pass-synchronous (Jacobi) relaxation of round 1; T > 0 = the ASYNCHRONOUS tile-local
// relaxation of zz_run_body_async with T tiles: per-tile queues (lattice: one per checkerboard colour, processed alternately),
// dedupe bits, inboxes for marks that cross a tile boundary (delivered with a random delay), list tags owned by the publisher,
// every evaluation with the freshest lists (cur = ~0).  Tiles take turns in a pseudo-random order derived from `seed`, so the
// tests exercise many interleavings; the fixed point -- hence every output bit -- must not depend on it.
static int g_async_tiles = 0;
static uint64_t g_async_seed = 1;

extern "C" {

void zzw_set_schedule(int tiles, uint64_t seed) { g_async_tiles = tiles; g_async_seed = seed ? seed : 1; }

// exhaustive check of the multiply-shift lattice-column formula against the division, for tests
int64_t zzw_check_grid_col(int32_t M, int64_t jmax)
{
    ZzGraph g; memset(&g, 0, sizeof g); g.grid_m = M; zz_grid_set_magic(g);
    int64_t bad = 0;
    for (int64_t j = 0; j <= jmax; ++j) bad += (zz_grid_col(g, (int32_t)j) != (int32_t)(j / M));
    const int32_t edge[] = { 0x7fffffff, 0x7ffffffe, 0x40000000 };
    for (int32_t j : edge) bad += (zz_grid_col(g, j) != j / M);
    return bad;
}

zzw_run* zzw_spdmp(int64_t d, const int64_t* tcp, const int64_t* trv, const double* tnz, const double* h,
                   const int64_t* bcp, const int64_t* brv, const double* bnz, const double* mu, double t0,
                   const double* x0, const double* th0, double T, const double* c_in, const uint64_t* seed,
                   int adapt, double factor, double delta0, double target_frac, uint32_t tag_limit, int local_bound, const double* kappa,
                   const double* boom_sigma, double boom_lambdaref, double boom_rho)
{
    return zzw_impl(d, tcp, trv, tnz, h, bcp, brv, bnz, mu, t0, x0, th0, T, c_in, seed, adapt, factor, delta0, target_frac,
                    tag_limit, local_bound, kappa, boom_sigma, boom_lambdaref, boom_rho, nullptr);
}

// ZigZag with velocity refreshments (Z.lambdaref > 0, src/sfact.jl:78-114); contract: zzo_spdmp_refresh in mode ctr | lazy
zzw_run* zzw_spdmp_refresh(int64_t d, const int64_t* tcp, const int64_t* trv, const double* tnz, const double* h,
                           const int64_t* bcp, const int64_t* brv, const double* bnz, const double* mu, const double* sigma, double lambdaref,
                           double t0, const double* x0, const double* th0, double T, const double* c_in, const uint64_t* seed,
                           int adapt, double factor, double delta0, double target_frac, uint32_t tag_limit)
{
    return zzw_impl(d, tcp, trv, tnz, h, bcp, brv, bnz, mu, t0, x0, th0, T, c_in, seed, adapt, factor, delta0, target_frac,
                    tag_limit, 0, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, sigma, lambdaref);
}

// the schedule with the subsampled logistic target (zz_logit.h; arguments as zzb_problem_create_logistic + zzb_spdmp_run)
zzw_run* zzw_spdmp_logistic(int64_t d, int64_t n, const int64_t* acp, const int64_t* arv, const double* anz,
                            const int64_t* atcp, const int64_t* atrv, const double* atnz, const double* y, const double* ny,
                            const double* mu_cv, double gamma0, int64_t k,
                            const int64_t* bcp, const int64_t* brv, const double* bnz, const double* mu, double t0,
                            const double* x0, const double* th0, double T, const double* c_in, const uint64_t* seed,
                            int adapt, double factor, double delta0, double target_frac, uint32_t tag_limit)
{
    ZzHostLogit hl;
    std::string e = zz_build_logit(hl, d, n, acp, arv, anz, atcp, atrv, atnz, y, ny, mu_cv, gamma0, k);
    if (!e.empty()) { zzw_run* r = new zzw_run(); r->d = d; r->status = 4; r->msg = e; return r; }
    return zzw_impl(d, hl.dep_cp.data(), hl.dep_rv.data(), hl.dep_nz.data(), nullptr, bcp, brv, bnz, mu, t0, x0, th0, T, c_in,
                    seed, adapt, factor, delta0, target_frac, tag_limit | 0x80000000u, 0, nullptr, nullptr, 0.0, 0.0, &hl);
}

// the schedule with the strong-bound sparse sticky timeline (zz_strong.h; contract: zzo_sparsestickyzz_ctr).  Coordinates with
// x0 == 0 start frozen (velocity 0 in their record); scalar c and kappa.
zzw_run* zzw_sparsesticky(int64_t d, const int64_t* gcp, const int64_t* grv, const double* gnz, const double* h, const double* x0,
                          const double* th0, double T, double c, double kappa, int rule, const uint64_t* seed, double delta0,
                          double target_frac, uint32_t tag_limit)
{
    ZzStrong st; st.c = c; st.kappa = kappa; st.rule = rule; st.pad = 0;
    std::vector<double> th((size_t)d), cv((size_t)d, c), kap((size_t)d, kappa), zero((size_t)d, 0.0);
    for (int64_t j = 0; j < d; ++j) th[j] = (x0[j] != 0.0) ? th0[j] : 0.0;
    return zzw_impl(d, gcp, grv, gnz, h, gcp, grv, gnz, zero.data(), 0.0, x0, th.data(), T, cv.data(), seed, 0, 1.0, delta0, target_frac,
                    tag_limit | 0x80000000u, 0, kap.data(), nullptr, 0.0, 0.0, nullptr, &st);
}

// the same with a bound constant and a thaw rate PER coordinate, a start time and rule 2 = "continue with the velocity saved at the
// freeze" (asynchzz / sspdmp4; contract: zzo_strongsticky_ctr).  th0 is passed as given: the record set-up below freezes the
// coordinates that start at 0 and keeps their velocity, like zz_setup_kernel.
zzw_run* zzw_strongsticky(int64_t d, const int64_t* gcp, const int64_t* grv, const double* gnz, const double* h, double t0,
                          const double* x0, const double* th0, double T, const double* c, const double* kappa, int rule,
                          const uint64_t* seed, double delta0, double target_frac, uint32_t tag_limit)
{
    ZzStrong st; st.c = c[0]; st.kappa = kappa[0]; st.rule = rule; st.pad = 0;
    std::vector<double> th(th0, th0 + d), zero((size_t)d, 0.0);
    if (rule != 2) for (int64_t j = 0; j < d; ++j) if (x0[j] == 0.0) th[j] = 0.0;
    return zzw_impl(d, gcp, grv, gnz, h, gcp, grv, gnz, zero.data(), t0, x0, th.data(), T, c, seed, 0, 1.0, delta0, target_frac,
                    tag_limit | 0x80000000u, 0, kappa, nullptr, 0.0, 0.0, nullptr, &st);
}

}  // extern "C"

static zzw_run* zzw_impl(int64_t d, const int64_t* tcp, const int64_t* trv, const double* tnz, const double* h,
                   const int64_t* bcp, const int64_t* brv, const double* bnz, const double* mu, double t0,
                   const double* x0, const double* th0, double T, const double* c_in, const uint64_t* seed,
                   int adapt, double factor, double delta0, double target_frac, uint32_t tag_limit, int local_bound, const double* kappa,
                   const double* boom_sigma, double boom_lambdaref, double boom_rho, const ZzHostLogit* hl, const ZzStrong* st,
                   const double* refresh_sigma, double refresh_lambdaref)
{
    zzw_run* r = new zzw_run();
    r->d = d;
    const int sticky_opts = local_bound & (ZZ_STICKY_REVERSIBLE | ZZ_STICKY_STRONG_UB | ZZ_STICKY_ZZ);   // (the sspdmp option bits travel in `local_bound`)
    local_bound &= 1;
    ZzHostGraph G;
    std::vector<double> zero_mu((size_t)d, 0.0);
    // LocalBound (src/local.jl): the bound is built from the TARGET's derivatives; the sampler matrix is not used
    std::string e = local_bound ? zz_build_graph(G, d, tcp, trv, tnz, h, tcp, trv, tnz, zero_mu.data())
                                : zz_build_graph(G, d, tcp, trv, tnz, h, bcp, brv, bnz, mu);
    if (!e.empty()) { r->status = 4; r->msg = e; return r; }
    std::vector<ZzKin> kin(d);
    std::vector<double> flips((size_t)d * 2 * ZZ_MAXFLIP, 0.0), fth((size_t)d * 2 * ZZ_MAXFLIP, 0.0);
    std::vector<ZzPriv> priv(d);
    std::vector<double> tau(d);
    std::vector<uint32_t> kctr(d), dstamp(d, 0);
    std::vector<ZzSpec> spec(d);
    std::vector<double> vt(d), vl(d), vlb(d);
    r->acc.assign(d, 0); r->s1.assign(d, 0.0); r->s2.assign(d, 0.0);

    ZzGraph g; g.nptr = G.nptr.data(); g.nidx = G.nidx.data(); g.nwt = G.nwt.data(); g.nwb = G.nwb.data();
    g.nfl = G.nfl.data(); g.gmu = G.gmu.data(); g.h = G.has_h ? G.h.data() : nullptr; g.same = hl ? 0 : G.same;
    ZzLogit lg; memset(&lg, 0, sizeof lg);
    if (hl) {
        lg.acp = hl->acp.data(); lg.arow = hl->arow.data(); lg.aval = hl->aval.data(); lg.rp = hl->rp.data();
        lg.rcol = hl->rcol.data(); lg.rval = hl->rval.data(); lg.y = hl->y.data(); lg.ny = hl->ny.data(); lg.u0 = hl->u0.data();
        lg.gamma0 = hl->gamma0; lg.k = hl->k; lg.n = hl->n;
    }
    g.grid_m = (tag_limit & 0x80000000u) ? 0 : G.grid_m; g.grid_n = G.grid_n;  // top bit of tag_limit: force the CSR path
    tag_limit &= 0x7fffffffu;
    for (int q = 0; q < 5; ++q) g.grid_diag[q] = G.grid_diag[q];
    zz_grid_set_magic(g);
    ZzView v; memset(&v, 0, sizeof v); v.nranks = 1; v.hi = (int32_t)d; v.shard = (int32_t)d; v.d = (int32_t)d; v.kin = kin.data(); v.flips = flips.data(); v.priv = priv.data();
    v.tau = tau.data(); v.kctr = kctr.data(); v.seed0 = seed[0]; v.seed1 = seed[1]; v.adapt = adapt; v.factor = factor; v.local_bound = local_bound;
    v.sticky = kappa ? (1 | sticky_opts) : 0; v.fth = (kappa || boom_sigma) ? fth.data() : nullptr; v.kappa = kappa;
    std::vector<double> rst((size_t)d * 2, 0.0), rspec((size_t)d * 2, 0.0);
    v.refresh = refresh_sigma ? 1 : 0; v.rsig = refresh_sigma; v.rlam1 = refresh_lambdaref / (double)d; v.rst = rst.data(); v.rspec = rspec.data();
    if (refresh_sigma) v.fth = fth.data();
    if (refresh_sigma && !g.grid_m && G.maxdeg > ZZ_NB_WIDE) { r->status = 1; r->msg = "column too long"; return r; }
    const bool vel = kappa || boom_sigma || refresh_sigma;   // lists carry the velocity after each event
    v.boom = boom_sigma ? 1 : 0; v.bmu = mu; v.bsig = boom_sigma; v.bref_rate = boom_lambdaref / (double)d; v.brho = boom_rho;
    v.brhobar = sqrt(1 - boom_rho * boom_rho);
    if (boom_sigma && !g.grid_m && G.maxdeg > ZZ_NB_WIDE) { r->status = 1; r->msg = "column too long"; return r; }

    for (int64_t j = 0; j < d; ++j) {
        kin[j].theta = th0[j]; kin[j].tf = t0; kin[j].xf = x0[j]; kin[j].hdr[0] = kin[j].hdr[1] = 0;
        priv[j].c = c_in[j];
        if (st && st->rule == 2 && x0[j] == 0.0) { priv[j].told = th0[j]; kin[j].theta = 0.0; }   // (as zz_setup_kernel)
        if (kappa && !st && (sticky_opts & ZZ_STICKY_ZZ) && x0[j] == 0.0) { priv[j].a = th0[j]; kin[j].theta = 0.0; }
    }
    double F0 = ZZ_INF;
    for (int64_t j = 0; j < d; ++j) {
        if (st) { if (!zz_init_node_strong(g, v, *st, (int32_t)j, t0)) { r->status = 1; r->msg = "column too long"; return r; } }
        else if (v.boom) zz_init_node_boom(g, v, (int32_t)j, t0);
        else zz_init_node(g, v, (int32_t)j, t0);
        F0 = std::min(F0, tau[j]);
    }
    if (!(t0 < T)) goto finish;  // while t' < T never entered (sfact.jl:199)
    {
        ZzCtl ctl;
        double target = std::max(target_frac * (double)d, 4.0);
        zz_ctl_init(ctl, std::min(F0, T), T, delta0, target, std::max(0.25 * target, 2.0));
        uint32_t cur = 0;
        std::vector<int32_t> wl, next, touched;
        bool ovf_seen = false;   // like the kernel (ZZ_OVF_BIT): an overflowing evaluation ends the window at the next pass boundary
        auto handle = [&](int32_t j, const ZzNodeOut& o, uint32_t w0, uint32_t curtag) {
            int slot;
            uint32_t cnt = zz_pick_slot(kin[j].hdr[0], kin[j].hdr[1], w0, curtag, slot);
            bool same = (cnt == o.nflip);
            if (same && cnt) {
                const double* fl = &flips[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                for (uint32_t m = 0; m < cnt && same; ++m) same = (zz_d2u(fl[m]) == zz_d2u(o.fl[m]));
                const double* ft = &fth[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                for (uint32_t m = 0; vel && m < cnt && same; ++m) same = (zz_d2u(ft[m]) == zz_d2u(o.fth[m]));
            }
            if (!same) {
                int ws = (slot == 0) ? 1 : 0;
                double* fl = &flips[((size_t)j * 2 + ws) * ZZ_MAXFLIP];
                for (uint32_t m = 0; m < o.nflip; ++m) fl[m] = o.fl[m];
                double* ft = &fth[((size_t)j * 2 + ws) * ZZ_MAXFLIP];
                for (uint32_t m = 0; vel && m < o.nflip; ++m) ft[m] = o.fth[m];
                kin[j].hdr[ws] = (curtag << 4) | o.nflip;
                for (int32_t q = G.dptr[j]; q < G.dptr[j + 1]; ++q) {
                    int32_t k = G.didx[q];
                    uint32_t old = dstamp[k];
                    if (old != curtag + 1) {
                        dstamp[k] = curtag + 1;
                        next.push_back(k);
                        if (old < w0) touched.push_back(k);
                    }
                }
            }
            ZzSpec& s = spec[j];
            s.a = o.a; s.b = o.b; s.told = o.told; s.tau = o.tau; s.c = o.c; s.k = o.k;
            s.nprop = (uint16_t)o.nprop; s.nflip = (uint8_t)o.nflip; s.flags = (uint8_t)o.flags;
            if (refresh_sigma) { rspec[2 * (size_t)j] = o.tprop; rspec[2 * (size_t)j + 1] = o.tref; }
            if (o.flags & ZZ_F_OVERFLOW) ovf_seen = true;
            vt[j] = o.viol_t; vl[j] = o.viol_l; vlb[j] = o.viol_lb;
            r->node_evals++;
        };
        while (ctl.phase != ZZ_PH_DONE && ctl.phase != ZZ_PH_FAIL) {
            if (cur > tag_limit) {  // tag rebase
                for (int64_t j = 0; j < d; ++j) { kin[j].hdr[0] = kin[j].hdr[1] = 0; dstamp[j] = 0; }
                cur = 0;
            }
            zz_ctl_begin(ctl);
            const uint32_t w0 = cur + 1;
            cur = w0;
            wl.clear(); next.clear(); touched.clear(); ovf_seen = false;
            ZzNodeOut o;
            int64_t it = 1;
            if (g_async_tiles > 0) {
                // ---------------- asynchronous tile-local relaxation (emulation of zz_run_body_async) ----------------
                const int NT = g_async_tiles;
                const int64_t per = (d + NT - 1) / NT;
                const uint32_t STRIDE = 256;
                uint64_t rs = g_async_seed * 0x9E3779B97F4A7C15ULL + (uint64_t)w0;
                auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; };
                auto colour = [&](int32_t k) -> int { if (!g.grid_m) return -1; const int32_t c = k / g.grid_m; return (int)(((k - c * g.grid_m) + c) & 1); };
                std::vector<std::vector<int32_t>> q[2] = { std::vector<std::vector<int32_t>>(NT), std::vector<std::vector<int32_t>>(NT) };
                std::vector<std::vector<int32_t>> inbox(NT);
                std::vector<uint8_t> dirty(d, 0), tb(d, 0);
                std::vector<int> cq(NT, 0);
                for (int64_t j = 0; j < d; ++j) {
                    const double tj = tau[j];
                    if (tj < ctl.H || (ctl.incl && tj == ctl.H)) {
                        const int c = colour((int32_t)j);
                        q[c < 0 ? 0 : c][j / per].push_back((int32_t)j);
                        dirty[j] = 1; tb[j] = 1; touched.push_back((int32_t)j);
                    }
                }
                auto mark = [&](int tile, int32_t k, int nxt) {   // a local mark of tile `tile` (or a drained inbox entry)
                    if (dirty[k]) return;
                    dirty[k] = 1;
                    const int c = colour(k);
                    q[c < 0 ? nxt : c][tile].push_back(k);
                    if (!tb[k]) { tb[k] = 1; touched.push_back(k); }
                };
                auto publish = [&](int tile, int32_t j, const ZzNodeOut& oo, int nxt) {
                    int slot;
                    uint32_t cnt = zz_pick_slot(oo.hdr0, oo.hdr1, w0, 0xffffffffu, slot);
                    bool same = (cnt == oo.nflip);
                    if (same && cnt) {
                        const double* fl = &flips[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                        for (uint32_t m = 0; m < cnt && same; ++m) same = (zz_d2u(fl[m]) == zz_d2u(oo.fl[m]));
                        const double* ft = &fth[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                        for (uint32_t m = 0; vel && m < cnt && same; ++m) same = (zz_d2u(ft[m]) == zz_d2u(oo.fth[m]));
                    }
                    uint32_t fl_ = oo.flags;
                    if (!same) {
                        const int wsl = (slot == 0) ? 1 : 0;
                        const uint32_t newtag = (slot < 0) ? w0 : (((slot == 0) ? oo.hdr0 : oo.hdr1) >> 4) + 1u;
                        if (newtag - w0 >= STRIDE) fl_ |= ZZ_F_OVERFLOW;
                        else {
                            double* fl = &flips[((size_t)j * 2 + wsl) * ZZ_MAXFLIP];
                            for (uint32_t m = 0; m < oo.nflip; ++m) fl[m] = oo.fl[m];
                            double* ft = &fth[((size_t)j * 2 + wsl) * ZZ_MAXFLIP];
                            for (uint32_t m = 0; vel && m < oo.nflip; ++m) ft[m] = oo.fth[m];
                            kin[j].hdr[wsl] = (newtag << 4) | oo.nflip;
                            for (int32_t qd = G.dptr[j]; qd < G.dptr[j + 1]; ++qd) {
                                const int32_t k = G.didx[qd];
                                if (k / per == tile) mark(tile, k, nxt);
                                else inbox[k / per].push_back(k);        // delivered when the owner drains
                            }
                        }
                    }
                    ZzSpec& sp = spec[j];
                    sp.a = oo.a; sp.b = oo.b; sp.told = oo.told; sp.tau = oo.tau; sp.c = oo.c; sp.k = oo.k;
                    sp.nprop = (uint16_t)oo.nprop; sp.nflip = (uint8_t)oo.nflip; sp.flags = (uint8_t)fl_;
                    if (refresh_sigma) { rspec[2 * (size_t)j] = oo.tprop; rspec[2 * (size_t)j + 1] = oo.tref; }
                    if (fl_ & ZZ_F_OVERFLOW) ovf_seen = true;
                    vt[j] = oo.viol_t; vl[j] = oo.viol_l; vlb[j] = oo.viol_lb;
                    r->node_evals++;
                };
                for (;;) {
                    // pick a tile with work (queue or inbox) at random; none left -> quiescent
                    std::vector<int> cand;
                    for (int tl = 0; tl < NT; ++tl) if (!q[0][tl].empty() || !q[1][tl].empty() || !inbox[tl].empty()) cand.push_back(tl);
                    if (cand.empty() || ovf_seen) break;
                    const int tl = cand[rnd() % cand.size()];
                    if (!inbox[tl].empty() && (rnd() & 1)) {            // drain (sometimes later: messages are asynchronous)
                        std::vector<int32_t> in; in.swap(inbox[tl]);
                        for (int32_t k : in) mark(tl, k, cq[tl]);
                    }
                    if (q[cq[tl]][tl].empty()) { if (!q[cq[tl] ^ 1][tl].empty()) cq[tl] ^= 1; else continue; }
                    std::vector<int32_t> cur_items; cur_items.swap(q[cq[tl]][tl]);
                    for (size_t a = cur_items.size(); a > 1; --a) std::swap(cur_items[a - 1], cur_items[rnd() % a]);   // order inside a round
                    ++it;
                    for (int32_t j : cur_items) {
                        dirty[j] = 0;
                        o.nitems = 0;
                        if (st) zz_process_node_strong(g, v, *st, j, ctl.H, ctl.incl, w0, 0xffffffffu, false, o);
                        else if (hl) zz_process_node_logit(g, v, lg, j, ctl.H, ctl.incl, w0, 0xffffffffu, false, o);
                        else zz_process_node(g, v, j, ctl.H, ctl.incl, w0, 0xffffffffu, false, o);
                        publish(tl, j, o, cq[tl] ^ 1);
                    }
                    cq[tl] ^= 1;
                }
                cur = w0 + STRIDE;   // every tag of this attempt lies below
            } else {

            for (int64_t j = 0; j < d; ++j) {
                double tj = tau[j];
                if (tj < ctl.H || (ctl.incl && tj == ctl.H)) {
                    // dirtied already by an earlier node of this pass? then it is in `next` as well; fine.
                    if (dstamp[j] < w0) { dstamp[j] = cur; touched.push_back((int32_t)j); }
                    if (st) zz_process_node_strong(g, v, *st, (int32_t)j, ctl.H, ctl.incl, w0, cur, true, o);
                    else if (hl) zz_process_node_logit(g, v, lg, (int32_t)j, ctl.H, ctl.incl, w0, cur, true, o);
                    else zz_process_node(g, v, (int32_t)j, ctl.H, ctl.incl, w0, cur, true, o);
                    handle((int32_t)j, o, w0, cur);
                }
            }
            for (;;) {
                cur++;
                wl.swap(next); next.clear();
                if (wl.empty() || ovf_seen) break;
                ++it;
                r->pass_hist[std::min<int64_t>(it, 63)] += (int64_t)wl.size();
                for (int32_t j : wl) {
                    o.nitems = 0;
                    if (st) zz_process_node_strong(g, v, *st, j, ctl.H, ctl.incl, w0, cur, false, o);
                    else if (hl) zz_process_node_logit(g, v, lg, j, ctl.H, ctl.incl, w0, cur, false, o);
                    else zz_process_node(g, v, j, ctl.H, ctl.incl, w0, cur, false, o);
                    r->item_hist[std::min<int64_t>(it, 63)] += o.nitems;
                    handle(j, o, w0, cur);
                }
            }
            }
            r->iters += it; r->max_iters = std::max(r->max_iters, it);
            bool overflow = ovf_seen; double smin = ZZ_INF; unsigned long long nprop = 0;  // accepted flips (length controller)
            for (int32_t j : touched) {
                const ZzSpec& s = spec[j];
                if (s.flags & ZZ_F_OVERFLOW) overflow = true;
                nprop += (unsigned long long)s.nprop | ((unsigned long long)s.nflip << 32);
                if (s.nflip) {
                    int slot; zz_pick_slot(kin[j].hdr[0], kin[j].hdr[1], w0, g_async_tiles > 0 ? 0xffffffffu : cur, slot);
                    smin = std::min(smin, flips[((size_t)j * 2 + slot) * ZZ_MAXFLIP]);
                }
            }
            int act = zz_ctl_end(ctl, overflow, smin, nprop);
            if (act != ZZ_ACT_COMMIT) { r->retries++; continue; }
            r->windows++;
            size_t seg0 = r->ev.size();
            for (int32_t j : touched) {
                const ZzSpec& s = spec[j];
                if ((s.flags & ZZ_F_STICKY_ERR) && r->status == 0) r->status = 9;
                if ((s.flags & ZZ_F_VIOL) && r->status == 0) {
                    r->status = 3; r->err_i = j + 1; r->err_t = vt[j]; r->err_l = vl[j]; r->err_lb = vlb[j];
                }
                priv[j].a = s.a; priv[j].b = s.b; priv[j].told = s.told; priv[j].c = s.c;
                tau[j] = s.tau; kctr[j] = s.k;
                if (refresh_sigma) { rst[2 * (size_t)j] = rspec[2 * (size_t)j]; rst[2 * (size_t)j + 1] = rspec[2 * (size_t)j + 1]; }
                r->num += s.nprop;
                if (s.nflip) {
                    int slot; zz_pick_slot(kin[j].hdr[0], kin[j].hdr[1], w0, g_async_tiles > 0 ? 0xffffffffu : cur, slot);
                    const double* fl = &flips[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                    double th = kin[j].theta, tf = kin[j].tf, xf = kin[j].xf;
                    const double* ft = &fth[((size_t)j * 2 + slot) * ZZ_MAXFLIP];
                    for (uint32_t m = 0; m < s.nflip; ++m) {
                        double fs = fl[m];
                        double xs, thn;
                        if (boom_sigma) {   // reflection / refreshment (zz_commit_node)
                            double tho; zz_boom_at(tf, xf, th, mu[j], fs, &xs, &tho);
                            thn = ft[m];
                            if (m == 0) r->acc[j] += (s.flags >> 3) & 7u;
                        } else if (refresh_sigma) {   // reflection or refreshment: the recorded velocity; reflections counted by the timeline
                            xs = xf + th * (fs - tf); thn = ft[m];
                            if (m == 0) r->acc[j] += (s.flags >> 3) & 7u;
                        } else if (kappa) {   // sticky: flip / freeze / thaw (zz_commit_node)
                            thn = ft[m];
                            if (thn == 0.0) xs = -0.0 * th;
                            else if (th == 0.0) xs = xf;
                            else { xs = xf + th * (fs - tf); r->acc[j] += 1; }
                        } else { xs = xf + th * (fs - tf); thn = -th; r->acc[j] += 1; }
                        if (!boom_sigma) {
                            r->s1[j] += (xf + xs) * (fs - tf);                    // trace.jl:194 (unscaled)
                            r->s2[j] += (fs - tf) * (xf * xf + xf * xs + xs * xs);
                        }
                        th = thn; tf = fs; xf = xs;
                        r->ev.push_back(zzw_event{ fs, j + 1, xs, th });          // sfact.jl:50-52
                    }
                    kin[j].theta = th; kin[j].tf = tf; kin[j].xf = xf;
                }
            }
            std::sort(r->ev.begin() + seg0, r->ev.end(), [](const zzw_event& a, const zzw_event& b) {
                return a.t < b.t || (a.t == b.t && a.i < b.i);
            });
            if (r->status) break;
        }
        if (ctl.phase == ZZ_PH_FAIL) { r->status = 9; r->msg = "window controller failed"; }
    }
finish:
    r->t.resize(d); r->x.resize(d); r->th.resize(d); r->c.resize(d);
    for (int64_t j = 0; j < d; ++j) { r->t[j] = kin[j].tf; r->x[j] = kin[j].xf; r->th[j] = kin[j].theta; r->c[j] = priv[j].c; }
    return r;
}

extern "C" {

int zzw_status(const zzw_run* r) { return r->status; }
void zzw_error_info(const zzw_run* r, int64_t* i, double* t, double* l, double* lb) { *i = r->err_i; *t = r->err_t; *l = r->err_l; *lb = r->err_lb; }
int64_t zzw_trace_len(const zzw_run* r) { return (int64_t)r->ev.size(); }
void zzw_trace_copy(const zzw_run* r, zzw_event* dst, int64_t first, int64_t count) { memcpy(dst, r->ev.data() + first, (size_t)count * sizeof(zzw_event)); }
void zzw_counts(const zzw_run* r, int64_t* acc, int64_t* num) { memcpy(acc, r->acc.data(), (size_t)r->d * 8); *num = r->num; }
void zzw_final_state(const zzw_run* r, double* t, double* x, double* th, double* c)
{
    size_t nb = (size_t)r->d * 8;
    memcpy(t, r->t.data(), nb); memcpy(x, r->x.data(), nb); memcpy(th, r->th.data(), nb); memcpy(c, r->c.data(), nb);
}
void zzw_sums(const zzw_run* r, double* s1, double* s2) { memcpy(s1, r->s1.data(), (size_t)r->d * 8); memcpy(s2, r->s2.data(), (size_t)r->d * 8); }
void zzw_stats(const zzw_run* r, int64_t* out) { out[0] = r->windows; out[1] = r->retries; out[2] = r->iters; out[3] = r->node_evals; out[4] = r->max_iters; }
void zzw_pass_hist(const zzw_run* r, int64_t* out) { memcpy(out, r->pass_hist.data(), 64 * 8); }
void zzw_item_hist(const zzw_run* r, int64_t* out) { memcpy(out, r->item_hist.data(), 64 * 8); }
void zzw_free(zzw_run* r) { delete r; }
}
