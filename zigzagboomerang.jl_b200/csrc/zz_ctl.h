// zz_ctl.h -- window controller shared by the persistent kernel (every CTA keeps an identical copy in
// registers) and by the host-side schedule emulation used in tests.
//
// Phases reproduce the termination rule of the reference's outer loop (src/sfact.jl:199-202): run while
// t' < T, where t' only advances on ACCEPTED events -- so every event before T is simulated (phase A),
// then proposals continue until the first accepted one at or after T (phase B finds its time s_min in a
// trial window that is discarded, phase C re-runs that window closed at s_min and commits it).
#ifndef ZZ_CTL_H
#define ZZ_CTL_H

#include "zz_math.h"

#define ZZ_PH_A 0
#define ZZ_PH_B 1
#define ZZ_PH_C 2
#define ZZ_PH_DONE 3
#define ZZ_PH_FAIL 4

#define ZZ_ACT_COMMIT 1
#define ZZ_ACT_RETRY 0

struct ZzCtl {
    double F;      // frontier: every event with time < F is final
    double H;      // end of the current window
    double T;      // user horizon
    double delta;  // current window length
    double target; // proposals per window the length controller aims for; accepted flips are capped at target_flips
    double target_flips;
    int phase;
    int incl;      // window closed on the right (phase C only)
};

ZZ_HD void zz_ctl_init(ZzCtl& c, double F0, double T, double delta0, double target, double target_flips)
{
    c.F = F0; c.T = T; c.delta = delta0; c.target = target; c.target_flips = target_flips; c.incl = 0; c.H = F0;
    c.phase = (F0 < T) ? ZZ_PH_A : ZZ_PH_B;
}

ZZ_HD void zz_ctl_begin(ZzCtl& c)
{
    if (c.phase == ZZ_PH_C) { c.incl = 1; return; }  // H was set to s_min
    c.incl = 0;
    c.H = c.F + c.delta;
    if (!(c.H > c.F)) c.H = zz_u2d(zz_d2u(c.F) + 1);  // delta underflowed against F: one ulp
    if (c.phase == ZZ_PH_A && c.H >= c.T) c.H = c.T;
}

// Called once the relaxation of the current window has converged.
//   overflow  some coordinate exceeded ZZ_MAXFLIP / ZZ_MAXITEMS
//   smin      earliest accepted flip inside the window (+inf if none); only used in phase B
//   nprop     proposals (low 32 bits) and accepted flips (high 32 bits) of the previously committed window; they
//             steer the window length only -- the events do not depend on it
ZZ_HD int zz_ctl_end(ZzCtl& c, bool overflow, double smin, unsigned long long nprop)
{
    if (overflow) {
        if (c.phase == ZZ_PH_C) { c.phase = ZZ_PH_FAIL; return ZZ_ACT_RETRY; }  // cannot happen: subset of a good window
        c.delta *= 0.5;
        return ZZ_ACT_RETRY;
    }
    if (c.phase == ZZ_PH_C) { c.F = c.H; c.phase = ZZ_PH_DONE; return ZZ_ACT_COMMIT; }
    if (c.phase == ZZ_PH_B && smin < ZZ_INF) { c.H = smin; c.phase = ZZ_PH_C; return ZZ_ACT_RETRY; }
    // commit and adapt the window length (any policy gives the same events; this one only steers speed)
    const unsigned long long np = nprop & 0xffffffffULL, nf = nprop >> 32;
    double f = c.target / (double)(np > 0 ? np : 1ULL);
    const double ff = c.target_flips / (double)(nf > 0 ? nf : 1ULL);
    f = ff < f ? ff : f;
    f = f < 0.5 ? 0.5 : (f > 2.0 ? 2.0 : f);
    const bool clipped = (c.phase == ZZ_PH_A && c.H >= c.T);
    if (!clipped) c.delta *= 0.5 * (1.0 + f);
    c.F = c.H;
    if (c.phase == ZZ_PH_A && c.F >= c.T) c.phase = ZZ_PH_B;
    return ZZ_ACT_COMMIT;
}

#endif  // ZZ_CTL_H
