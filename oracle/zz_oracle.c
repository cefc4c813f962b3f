/* zz_oracle.c -- CPU oracle for the factorised local-ZigZag event loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference (ZigZagBoomerang.jl @ 691afe2) is pure Julia, Julia is not
 * installed in this image, and the reference's tests hold no golden vector for this path
 * (SURVEY.md 8(c)): every sampler test is a law-of-large-numbers check with an unpinned seed.
 * This file is therefore a line-by-line RESTATEMENT of the algorithm, pinned only by the
 * reference's property tests (test/poisson.jl) and statistical tests (test/maintest.jl), which
 * tests/ re-runs against it.
 *
 * What is restated (file:line in /root/reference):
 *   spdmp setup + outer loop          src/sfact.jl:162-212   (zzo_spdmp)
 *   spdmp_inner!                      src/sfact.jl:73-145    (inner_event)
 *   smove_forward! (ZigZag)           src/sfact.jl:6-28      (move_nbhd, move_all)
 *   lambda / lambda-bar               src/fact_samplers.jl:28-30, src/sfact.jl:69-70
 *   ab (ZigZag)                       src/fact_samplers.jl:41,50-54   (ab_zigzag)
 *   adapt!                            src/fact_samplers.jl:67-70
 *   idot                              src/common.jl:16-24    (idot)
 *   poisson_time                      src/poissontime.jl:8-30 (o_poisson_time)
 *   SPriorityQueue                    src/priorityqueue.jl:44-117 (heap_*)
 *   reflect! (scalar)                 src/dynamics.jl:46-49
 *   event / Trace push                src/sfact.jl:50-52, src/trace.jl:38,73
 *   mean(::Trace)                     src/trace.jl:182-200   (zzo_moments, first moment)
 *   RNG                               src/ZigZagBoomerang.jl:6-10 -> RandomNumbers.jl 1.5.3
 *                                     Xorshifts.Xoroshiro128Plus (third-party, absent from
 *                                     /root/reference; restated from the published
 *                                     xoroshiro128+ algorithm, 2016 constants 55/14/36)
 *
 * Two independent switches (mode bits):
 *   RNG    0 = "seq": ONE xoroshiro128+ stream consumed in global event order exactly where the
 *              reference calls rand(rng) (sfact.jl:121,134,139,186).
 *          1 = "ctr": per-coordinate counter streams u(j,k) (zz_math.h): coordinate j consumes
 *              one draw every time IT is (re)scheduled and one every time IT proposes.  Same
 *              roles as the reference, but schedule-independent -- the GPU parity contract.
 *   ARITH  0 = "inplace": neighbours are advanced in place at every proposal like the reference.
 *          2 = "lazy": positions are anchored at the coordinate's own last flip,
 *              x_j(s) = xf_j + th_j (s - tf_j); only the owner ever rewrites them.
 *   GRAPH  4 = All() neighbourhood (pdmp, sfact.jl:236) instead of Matched().
 *   BOUND  8 = LocalBound variant (src/local.jl:2-6,10-78,95-149; next_time: src/not_fact_samplers.jl:43-50): the
 *              bound comes from the target's own first and second directional derivatives, expires after
 *              Delta = 2/c/|theta| and is then renewed.  The target descriptor plays the role of the extended-form
 *              closure: (grad phi_i, v_i) = (idot(G,i,x) - h_i, theta_i idot(G,i,theta)) (local.jl:7).
 * GPU results must equal mode ctr|lazy (|8) bit for bit.
 * Logistic target (zzo_spdmp_logistic): the same loop with the subsampled logistic-regression partial derivative of
 * scripts/logistic.jl:78-107 (fdot_moving / grad_phi_moving, sigmoid helpers :34,56-57, idot_moving! src/common.jl:33-42)
 * reached through the SelfMoving closure signature (src/sfact.jl:64-68); call site scripts/logistic.jl:167.
 * Further entry points below: zzo_sspdmp (sticky ZigZag, src/ss_fact.jl), zzo_spdmp_boom (FactBoomerang) and
 * zzo_parallel_spdmp (the reference's multithreaded parallel_spdmp, src/parallel.jl -- CPU baseline of bench.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#include "../zigzagboomerang.jl_b200/csrc/zz_math.h" /* zz_log, zz_u01, zz_sincos / zz_boom_at (shared primitives) */

#define ZZO_RNG_CTR 1
#define ZZO_ARITH_LAZY 2
#define ZZO_GRAPH_ALL 4
#define ZZO_LOCAL_BOUND 8 /* spdmp(..., C::LocalBound, ...) of src/local.jl:95-149 */
#define ZZO_STICKY_REVERSIBLE 16 /* sspdmp(...; reversible = true): a thawing coordinate re-enters with a random sign, ss_fact.jl:111-113 */
#define ZZO_STICKY_STRONG_UB 32  /* sspdmp(...; strong_upperbounds = true): a freeze reschedules nobody, ss_fact.jl:97-107 */
#define ZZO_STICKYZZ 64          /* the dense sticky sampler stickyzz / sspdmp2 (src/stickyzz.jl:176-338): the same loop with (i) proposal
                                    times drawn at rate 0.01 + (a + b t)^+ ("guarantee minimum rate", poissontime.jl:93-99 through
                                    queue_time! stickyzz.jl:144-165) and (ii) coordinates that start AT 0 start frozen with a thaw clock
                                    (:198-206) instead of being queued */

#define ZZO_OK 0
#define ZZO_E_BOUND 3 /* "Tuning parameter `c` too small." sfact.jl:124 */
#define ZZO_E_GRAPH 4
#define ZZO_E_ARG 1

typedef struct { double t; int64_t i; double x; double th; } zzo_event; /* trace.jl:38 */

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

typedef struct {
    int64_t d;
    int mode;
    /* results */
    zzo_event *ev; int64_t nev, cap;
    int64_t *acc; int64_t num;
    double *t, *x, *th; /* final state */
    double *c;
    double *x0; double t0;
    int status; int64_t err_i; double err_t, err_l, err_lb;
    double loop_seconds;   /* wall time of the event loop alone (setup excluded), for the CPU baselines of bench.py */
} zzo_run;

/* ---- poisson_time, src/poissontime.jl:8-30 ------------------------------------------------ */
static double o_poisson_time(double a, double b, double u)
{
    if (b > 0) {
        if (a < 0)
            return sqrt(-zz_log(u) * 2.0 / b) - a / b;
        else
            return sqrt((a / b) * (a / b) - zz_log(u) * 2.0 / b) - a / b;
    } else if (b == 0) {
        if (a > 0)
            return -zz_log(u) / a;
        else
            return INFINITY;
    } else {
        if (a <= 0)
            return INFINITY;
        else if (-zz_log(u) <= -(a * a) / b + (a * a) / (2 * b))
            return -sqrt((a / b) * (a / b) - zz_log(u) * 2.0 / b) - a / b;
        else
            return INFINITY;
    }
}
double zzo_poisson_time(double a, double b, double u) { return o_poisson_time(a, b, u); }

/* three-parameter form c + (a+bt)^+, src/poissontime.jl:39-65 */
double zzo_poisson_time3(double a, double b, double c, double u)
{
    if (b > 0) {
        if (a < 0) {
            if (-c * a / b + zz_log(u) < 0.0)
                return sqrt(-2 * b * zz_log(u) + c * c + 2 * a * c) / b - (a + c) / b;
            else
                return -zz_log(u) / c;
        } else
            return sqrt(-zz_log(u) * 2.0 * b + (a + c) * (a + c)) / b - (a + c) / b;
    } else if (b == 0) {
        if (a > 0)
            return -zz_log(u) / (a + c);
        else
            return -zz_log(u) / c;
    } else {
        if (a <= 0.0)
            return -zz_log(u) / c;
        else if (-c * a / b - (a * a) / (2 * b) + zz_log(u) > 0.0)
            return +sqrt((a + c) * (a + c) - 2.0 * zz_log(u) * b) / b - (a + c) / b;
        else
            return (-zz_log(u) + (a * a) / (2 * b)) / c;
    }
}
double zzo_log(double x) { return zz_log(x); }
double zzo_u01(uint64_t s0, uint64_t s1, uint64_t i, uint64_t k) { return zz_u01(s0, s1, i, k); }

/* HOST build of the shared primitives on arrays: the counterpart of zzb_math_probe (device build), same kinds */
void zzo_math_probe(int kind, int64_t n, const double* x, const double* y, const double* z, double* o1, double* o2)
{
    for (int64_t k = 0; k < n; ++k) {
        if (kind == 0) o1[k] = zz_log(x[k]);
        else if (kind == 1) o1[k] = zz_exp(x[k]);
        else if (kind == 2) zz_sincos(x[k], &o1[k], &o2[k]);
        else if (kind == 3) o1[k] = zz_poisson_time(x[k], y[k], z[k]);
        else o1[k] = zz_u01(zz_d2u(x[0]), zz_d2u(y[0]), (uint64_t)k, (uint64_t)(k ^ 0x5bd1));
    }
}

/* ---- xoroshiro128+ (RandomNumbers.jl Xorshifts.Xoroshiro128Plus; restated, unpinned) ------- */
typedef struct { uint64_t x, y; } xoro;
static inline uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
static inline uint64_t xoro_next(xoro *r)
{
    uint64_t s0 = r->x, s1 = r->y, p = s0 + s1;
    s1 ^= s0;
    r->x = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    r->y = rotl64(s1, 36);
    return p;
}
static inline double xoro_rand(xoro *r)
{ /* 52 high bits into the mantissa of [1,2), minus one -> [0,1) */
    uint64_t v = (xoro_next(r) >> 12) | 0x3ff0000000000000ULL;
    double d; memcpy(&d, &v, 8);
    return d - 1.0;
}

/* ---- SPriorityQueue, src/priorityqueue.jl ------------------------------------------------- */
typedef struct { int64_t n; int64_t *key; double *val; int64_t *index; int lex; } heapq; /* 1-based */
static inline int h_lt(const heapq *q, double va, int64_t ka, double vb, int64_t kb)
{ /* lt(o, a, b) on the priorities (:44-75); with lex set, ties are broken by key so that the
     ctr-mode event order is a strict total order (DESIGN.md "ties"). */
    if (va < vb) return 1;
    if (q->lex && va == vb) return ka < kb;
    return 0;
}
static void h_down(heapq *q, int64_t i)
{ /* percolate_down! :44-59 */
    int64_t xk = q->key[i]; double xv = q->val[i];
    int64_t l;
    while ((l = 2 * i) <= q->n) {
        int64_t r = 2 * i + 1;
        int64_t j = (r > q->n || h_lt(q, q->val[l], q->key[l], q->val[r], q->key[r])) ? l : r;
        if (h_lt(q, q->val[j], q->key[j], xv, xk)) {
            q->index[q->key[j]] = i; q->key[i] = q->key[j]; q->val[i] = q->val[j];
            i = j;
        } else break;
    }
    q->index[xk] = i; q->key[i] = xk; q->val[i] = xv;
}
static void h_up(heapq *q, int64_t i)
{ /* percolate_up! :61-75 */
    int64_t xk = q->key[i]; double xv = q->val[i];
    while (i > 1) {
        int64_t j = i / 2;
        if (h_lt(q, xv, xk, q->val[j], q->key[j])) {
            q->index[q->key[j]] = i; q->key[i] = q->key[j]; q->val[i] = q->val[j];
            i = j;
        } else break;
    }
    q->index[xk] = i; q->key[i] = xk; q->val[i] = xv;
}
static void h_enqueue(heapq *q, int64_t key, double v)
{ /* enqueue! :105-115 (keys arrive in order 1..n) */
    q->n += 1; q->key[q->n] = key; q->val[q->n] = v; q->index[key] = q->n;
    h_up(q, q->n);
}
static void h_set(heapq *q, int64_t key, double v)
{ /* setindex! :93-103 */
    int64_t i = q->index[key];
    double old = q->val[i];
    q->val[i] = v;
    if (h_lt(q, old, key, v, key)) h_down(q, i); else h_up(q, i);
}

/* ---- sparse helpers ------------------------------------------------------------------------ */
typedef struct { const int64_t *colptr, *rowval; const double *nzval; } csc; /* 1-based, Julia layout */

static double idot(const csc *A, int64_t j, const double *x)
{ /* src/common.jl:16-24; j and rowval 1-based, x 0-based storage */
    double s = 0.0;
    for (int64_t p = A->colptr[j - 1]; p < A->colptr[j]; ++p)
        s += A->nzval[p - 1] * x[A->rowval[p - 1] - 1];
    return s;
}

typedef struct {
    int64_t d; int mode;
    csc tg, bd; const double *h, *mu;
    double *t, *x, *th, *t_old, *ba, *bb, *c;
    double *tf, *xf;           /* lazy anchors */
    uint32_t *kctr;            /* ctr-mode per-coordinate draw counters */
    uint64_t s0, s1; xoro rng;
    int64_t *g2ptr, *g2idx;    /* G2 neighbourhoods (sfact.jl:178), 1-based ids */
    double *scratch;           /* lazy mode: positions/velocities gathered for idot */
} ctx;

static inline double draw(ctx *z, int64_t j /*1-based coordinate consuming the draw*/)
{
    if (z->mode & ZZO_RNG_CTR) return zz_u01(z->s0, z->s1, (uint64_t)(j - 1), z->kctr[j - 1]++);
    return xoro_rand(&z->rng);
}

/* position of coordinate k (1-based) at time s */
static inline double pos_at(const ctx *z, int64_t k, double s)
{
    if (z->mode & ZZO_ARITH_LAZY) return z->xf[k - 1] + z->th[k - 1] * (s - z->tf[k - 1]);
    return z->x[k - 1]; /* inplace: caller has moved it to s already */
}

/* idot over positions at time s (lazy: evaluate on the fly in storage order) */
static double idot_x(const ctx *z, const csc *A, int64_t j, double s)
{
    if (!(z->mode & ZZO_ARITH_LAZY)) return idot(A, j, z->x);
    double acc = 0.0;
    for (int64_t p = A->colptr[j - 1]; p < A->colptr[j]; ++p)
        acc += A->nzval[p - 1] * pos_at(z, A->rowval[p - 1], s);
    return acc;
}

/* ab(G,i,x,th,c,Z::ZigZag), src/fact_samplers.jl:50-54 with loosen(c,x) = c + x (:41) */
static void ab_zigzag(ctx *z, int64_t i, double s)
{
    double a = z->c[i - 1] + (idot_x(z, &z->bd, i, s) - idot(&z->bd, i, z->mu)) * z->th[i - 1];
    double b = z->c[i - 1] / 100 + z->th[i - 1] * idot(&z->bd, i, z->th);
    z->ba[i - 1] = a; z->bb[i - 1] = b;
}

/* smove_forward!(G,i,...) over column i of the bound matrix (G = G1, Matched), sfact.jl:6-12 */
static void move_nbhd(ctx *z, const int64_t *idx, int64_t n, double tp)
{
    for (int64_t q = 0; q < n; ++q) {
        int64_t k = idx[q] - 1;
        z->x[k] = z->x[k] + z->th[k] * (tp - z->t[k]);
        z->t[k] = tp;
    }
}
static void move_all(ctx *z, double tp)
{ /* sfact.jl:23-28 */
    for (int64_t k = 0; k < z->d; ++k) {
        z->x[k] = z->x[k] + z->th[k] * (tp - z->t[k]);
        z->t[k] = tp;
    }
}

static void push_event(zzo_run *r, double t, int64_t i, double x, double th)
{
    if (r->nev == r->cap) {
        r->cap = r->cap ? 2 * r->cap : 1024;
        r->ev = (zzo_event *)realloc(r->ev, (size_t)r->cap * sizeof(zzo_event));
    }
    zzo_event e = { t, i, x, th };
    r->ev[r->nev++] = e;
}

/* Build G2[i] = setdiff(union(G1[j] for j in G1[i]), G[i]) with G = G1 (sfact.jl:178). */
static void build_g2(ctx *z)
{
    int64_t d = z->d;
    z->g2ptr = (int64_t *)calloc((size_t)d + 1, sizeof(int64_t));
    int64_t *mark = (int64_t *)calloc((size_t)d + 1, sizeof(int64_t));
    int64_t cap = 16, n = 0;
    z->g2idx = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
    for (int64_t i = 1; i <= d; ++i) {
        z->g2ptr[i - 1] = n;
        for (int64_t p = z->bd.colptr[i - 1]; p < z->bd.colptr[i]; ++p) mark[z->bd.rowval[p - 1]] = -i; /* in G[i] */
        for (int64_t p = z->bd.colptr[i - 1]; p < z->bd.colptr[i]; ++p) {
            int64_t j = z->bd.rowval[p - 1];
            for (int64_t q = z->bd.colptr[j - 1]; q < z->bd.colptr[j]; ++q) {
                int64_t k = z->bd.rowval[q - 1];
                if (mark[k] == -i || mark[k] == i) continue;
                mark[k] = i;
                if (n == cap) { cap *= 2; z->g2idx = (int64_t *)realloc(z->g2idx, (size_t)cap * sizeof(int64_t)); }
                z->g2idx[n++] = k;
            }
        }
    }
    z->g2ptr[d] = n;
    free(mark);
}

/* ab(G,i,x,th,C::LocalBound,grad_i,v_i,Z::ZigZag), src/local.jl:2-6; returns Delta */
static double ab_local(ctx *z, int64_t i, double s)
{
    double gi = idot_x(z, &z->tg, i, s);
    if (z->h) gi = gi - z->h[i - 1];
    double vi = idot(&z->tg, i, z->th) * z->th[i - 1];
    z->ba[i - 1] = z->c[i - 1] + gi * z->th[i - 1];
    z->bb[i - 1] = z->c[i - 1] / 100 + vi;
    return 2.0 / z->c[i - 1] / fabs(z->th[i - 1]);
}

/* next_time(t, abc, z), src/not_fact_samplers.jl:43-50 */
static double next_time(double t, double a, double b, double Delta, double u, char *renew)
{
    double dt = o_poisson_time(a, b, u);
    if (dt > Delta) { *renew = 1; return t + Delta; }
    *renew = 0;
    return t + dt;
}

/* ---- subsampled logistic target, scripts/logistic.jl (config 3) --------------------------------
 * grad phi_j = gamma0*x[j] - fdot_moving(A, At, j, t, x, theta, t', F, mu, y, ny, k)   (:78-95,107), called through the
 * SelfMoving / ExtendedForm signature (src/sfact.jl:68).  A is the n x d design matrix (CSC), At = A' (CSC: column r =
 * row r of A), y / ny the successes / failures per row, mu the mode (control variate), k the number of rows drawn.
 * The reference draws the row indices from Julia's GLOBAL RNG (Random.SamplerRangeNDL, :83,86): mode seq uses a second
 * xoroshiro stream for them, mode ctr the proposing coordinate's own counter stream (k draws before the thinning draw). */
typedef struct {
    int64_t n; csc A, At; const double *y, *ny, *mu; double gamma0; int64_t k; xoro grng;
} logit;

static double sigmoid_(double x) { return 1.0 / (1.0 + zz_exp(-x)); } /* :34 inv(one(x) + exp(-x)) */
static double sigmoidn_(double x) { return sigmoid_(-x); }              /* :56 */
static double nsigmoid_(double x) { return -sigmoid_(x); }              /* :57 */

static double logit_grad(ctx *z, logit *L, int64_t j, double tp)
{
    const int64_t r0 = L->A.colptr[j - 1], l = L->A.colptr[j] - r0; /* nzrange(A, j) */
    double s = 0.0;
    for (int64_t it = 0; it < L->k; ++it) {
        double ur = (z->mode & ZZO_RNG_CTR) ? draw(z, j) : xoro_rand(&L->grng);
        int64_t i = (int64_t)(ur * (double)l);
        if (i >= l) i = l - 1;
        i += r0; /* 1-based position in rowvals(A) */
        int64_t row = L->A.rowval[i - 1];
        double val = L->A.nzval[i - 1];
        double u = 0.0; /* idot_moving!(At, row, t, x, theta, t', F), src/common.jl:33-42 */
        for (int64_t q = L->At.colptr[row - 1]; q < L->At.colptr[row]; ++q) {
            int64_t m = L->At.rowval[q - 1];
            if (!(z->mode & ZZO_ARITH_LAZY)) { /* smove_forward!(m, t, x, theta, t', F), sfact.jl:14-16 */
                z->x[m - 1] = z->x[m - 1] + z->th[m - 1] * (tp - z->t[m - 1]);
                z->t[m - 1] = tp;
            }
            u += L->At.nzval[q - 1] * pos_at(z, m, tp);
        }
        double lk = (double)l / (double)L->k;
        s += lk * val * L->y[row - 1] * sigmoidn_(u);   /* :87 */
        s += lk * val * L->ny[row - 1] * nsigmoid_(u);  /* :88 */
        double u0 = idot(&L->At, row, L->mu);           /* :89 */
        s -= lk * val * L->y[row - 1] * sigmoidn_(u0);  /* :90 */
        s -= lk * val * L->ny[row - 1] * nsigmoid_(u0); /* :91 */
    }
    return L->gamma0 * pos_at(z, j, tp) - s;            /* :107 */
}

/* velocity refreshments of the ZigZag in spdmp (hasrefresh(Z) = Z.lambdaref > 0, fact_samplers.jl:19): sfact.jl:78-114,188-190.
 * One clock of rate lambdaref at queue key n+1; it picks a coordinate with the GLOBAL rng -- twice (:80,:84: the
 * neighbourhood of the first pick is moved, the second pick is refreshed) -- sets theta_i = sigma_i * rand(rng, (-1, 1))
 * (:100-101), re-arms with waiting_time_ref(F) (global rng, :107; dynamics.jl:100-101) and reschedules G1[i] (:109-113); the
 * refreshment is a trace event (:143).  Mode seq: both generators are the one xoroshiro stream (documented deviation for
 * the global-rng draws).  Mode ctr: one clock of rate lambdaref/d PER coordinate (superposition), its draws from the
 * coordinate's own stream.  No device kernel runs this yet (the host API refuses lambdaref > 0). */
typedef struct { double lambdaref; const double *sigma; } refr;

static zzo_run *spdmp_impl(int64_t d,
                   const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                   const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                   double t0, const double *x0, const double *th0, double T, const double *c_in,
                   const uint64_t *seed, int adapt, double factor, int mode, logit *lg, const refr *rf)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    ctx zs; ctx *z = &zs; memset(z, 0, sizeof(ctx));
    r->d = d; r->mode = mode; r->t0 = t0;
    z->d = d; z->mode = mode;
    z->tg.colptr = tg_colptr; z->tg.rowval = tg_rowval; z->tg.nzval = tg_nzval;
    z->bd.colptr = bd_colptr; z->bd.rowval = bd_rowval; z->bd.nzval = bd_nzval;
    z->h = h; z->mu = mu;
    size_t nb = (size_t)d * sizeof(double);
    z->t = (double *)malloc(nb); z->x = (double *)malloc(nb); z->th = (double *)malloc(nb);
    z->t_old = (double *)malloc(nb); z->ba = (double *)malloc(nb); z->bb = (double *)malloc(nb);
    z->c = (double *)malloc(nb); z->tf = (double *)malloc(nb); z->xf = (double *)malloc(nb);
    z->kctr = (uint32_t *)calloc((size_t)d, sizeof(uint32_t));
    r->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    z->s0 = seed[0]; z->s1 = seed[1]; z->rng.x = seed[0]; z->rng.y = seed[1];
    const int lazy = (mode & ZZO_ARITH_LAZY) != 0, all = (mode & ZZO_GRAPH_ALL) != 0;

    /* sfact.jl:167-169,180 */
    double tp = t0;
    for (int64_t k = 0; k < d; ++k) {
        z->t[k] = t0; z->t_old[k] = t0; z->x[k] = x0[k]; z->th[k] = th0[k]; z->c[k] = c_in[k];
        z->tf[k] = t0; z->xf[k] = x0[k];
    }
    if (!all && !lazy) build_g2(z); /* sfact.jl:171-179 (lazy arithmetic needs no moves at all) */

    heapq Q; Q.n = 0; Q.lex = (mode & ZZO_RNG_CTR) != 0;
    Q.key = (int64_t *)malloc((2 * (size_t)d + 2) * sizeof(int64_t));
    Q.val = (double *)malloc((2 * (size_t)d + 2) * sizeof(double));
    Q.index = (int64_t *)malloc((2 * (size_t)d + 2) * sizeof(int64_t));
    const int lbnd = (mode & ZZO_LOCAL_BOUND) != 0;
    const int ctr = (mode & ZZO_RNG_CTR) != 0;
    const double lam1 = rf ? (ctr ? rf->lambdaref / (double)d : rf->lambdaref) : 0.0;
    char *renew = (char *)calloc((size_t)d, 1);
    if (lbnd) {
        /* local.jl:118-123: bound and first time per coordinate, t[i] + ... (here t0 IS added) */
        for (int64_t i = 1; i <= d; ++i) {
            double Delta = ab_local(z, i, t0);
            h_enqueue(&Q, i, next_time(t0, z->ba[i - 1], z->bb[i - 1], Delta, draw(z, i), &renew[i - 1]));
        }
    } else {
        /* sfact.jl:184-187: bounds, then the queue is filled in coordinate order, NO "+ t0" */
        for (int64_t i = 1; i <= d; ++i) ab_zigzag(z, i, t0);
        for (int64_t i = 1; i <= d; ++i) h_enqueue(&Q, i, o_poisson_time(z->ba[i - 1], z->bb[i - 1], draw(z, i)));
    }

    if (rf) { /* sfact.jl:188-190: enqueue!(Q, (n + 1) => waiting_time_ref(rng, F)), no "+ t0" */
        if (ctr) { for (int64_t i = 1; i <= d; ++i) h_enqueue(&Q, d + i, -zz_log(draw(z, i)) / lam1); }
        else h_enqueue(&Q, d + 1, -zz_log(xoro_rand(&z->rng)) / lam1);
    }
    int64_t num = 0;
    const double tl0 = now_s();
    /* sfact.jl:199 outer loop; body = spdmp_inner! (:73-145) */
    while (tp < T && r->status == ZZO_OK) {
        for (;;) {
            int64_t i = Q.key[1]; tp = Q.val[1]; /* peek :77 */
            const int refresh = i > d;           /* :78 */
            const int64_t qkey = i;
            if (refresh) i = ctr ? i - d : (int64_t)(xoro_rand(&z->rng) * (double)d) + 1;   /* :80 rand(1:n) */
            const int64_t *nb = &z->bd.rowval[z->bd.colptr[i - 1] - 1];
            int64_t nnb = z->bd.colptr[i] - z->bd.colptr[i - 1];
            if (!lazy) { if (all) move_all(z, tp); else move_nbhd(z, nb, nnb, tp); } /* :82 */
            if (refresh) {
                if (!ctr) i = (int64_t)(xoro_rand(&z->rng) * (double)d) + 1;                 /* :84 second pick */
                nb = &z->bd.rowval[z->bd.colptr[i - 1] - 1]; nnb = z->bd.colptr[i] - z->bd.colptr[i - 1];
                if (!lazy && !all) move_nbhd(z, &z->g2idx[z->g2ptr[i - 1]], z->g2ptr[i] - z->g2ptr[i - 1], tp); /* :85 */
                if (lazy) { z->xf[i - 1] = pos_at(z, i, tp); z->tf[i - 1] = tp; }
                z->th[i - 1] = rf->sigma[i - 1] * (draw(z, i) < 0.5 ? -1.0 : 1.0);            /* :100-101 */
                h_set(&Q, qkey, tp - zz_log(ctr ? draw(z, i) : xoro_rand(&z->rng)) / lam1);  /* :107 */
                for (int64_t q = 0; q < nnb; ++q) {                                          /* :109-113 */
                    int64_t j = nb[q];
                    double tj = lazy ? tp : z->t[j - 1];
                    z->t_old[j - 1] = tj;
                    ab_zigzag(z, j, tp);
                    h_set(&Q, j, tj + o_poisson_time(z->ba[j - 1], z->bb[j - 1], draw(z, j)));
                }
                push_event(r, lazy ? tp : z->t[i - 1], i, lazy ? z->xf[i - 1] : z->x[i - 1], z->th[i - 1]); /* :143 */
                break;
            }
            if (lbnd && renew[i - 1]) { /* local.jl:34-41: the bound expired, renew it */
                double Delta = ab_local(z, i, tp);
                double tr = lazy ? tp : z->t[i - 1];
                z->t_old[i - 1] = tr;
                h_set(&Q, i, next_time(tr, z->ba[i - 1], z->bb[i - 1], Delta, draw(z, i), &renew[i - 1]));
                continue;
            }
            double gi;
            if (lg) gi = logit_grad(z, lg, i, tp);          /* :118 through the SelfMoving closure (sfact.jl:68) */
            else {
                gi = idot_x(z, &z->tg, i, tp);              /* :118, user closure = idot(Gamma,i,x) */
                if (h) gi = gi - h[i - 1];
            }
            double ti = lazy ? tp : z->t[i - 1];
            double l = zz_pos(gi * z->th[i - 1]);                                   /* :119, fact_samplers.jl:28-30 */
            double lb = zz_pos(z->ba[i - 1] + z->bb[i - 1] * (ti - z->t_old[i - 1])); /* sfact.jl:70 */
            num += 1;                                                               /* :120 */
            if (draw(z, i) * lb < l) {                                              /* :121 */
                r->acc[i - 1] += 1;
                if (l >= lb) {                                                      /* :123-128 */
                    if (!adapt) {
                        r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb;
                        break;
                    }
                    z->c[i - 1] *= factor;
                }
                if (!lazy && !all)                                                  /* :129 */
                    move_nbhd(z, &z->g2idx[z->g2ptr[i - 1]], z->g2ptr[i] - z->g2ptr[i - 1], tp);
                if (lazy) { z->xf[i - 1] = pos_at(z, i, tp); z->tf[i - 1] = tp; }
                z->th[i - 1] = -z->th[i - 1];                                       /* :130, dynamics.jl:46-49 */
                for (int64_t q = 0; q < nnb; ++q) {                                 /* :131-135, local.jl:60-66 */
                    int64_t j = nb[q];
                    double tj = lazy ? tp : z->t[j - 1];
                    z->t_old[j - 1] = tj;
                    if (lbnd) {
                        double Delta = ab_local(z, j, tp);
                        h_set(&Q, j, next_time(tj, z->ba[j - 1], z->bb[j - 1], Delta, draw(z, j), &renew[j - 1]));
                    } else {
                        ab_zigzag(z, j, tp);
                        h_set(&Q, j, tj + o_poisson_time(z->ba[j - 1], z->bb[j - 1], draw(z, j)));
                    }
                }
                push_event(r, ti, i, lazy ? z->xf[i - 1] : z->x[i - 1], z->th[i - 1]); /* :143, :50-52 */
                break;
            } else {                                                                /* :137-140, local.jl:68-73 */
                z->t_old[i - 1] = ti;
                if (lbnd) {
                    double Delta = ab_local(z, i, tp);
                    h_set(&Q, i, next_time(ti, z->ba[i - 1], z->bb[i - 1], Delta, draw(z, i), &renew[i - 1]));
                } else {
                    ab_zigzag(z, i, tp);
                    h_set(&Q, i, ti + o_poisson_time(z->ba[i - 1], z->bb[i - 1], draw(z, i)));
                }
            }
        }
    }
    r->num = num;
    r->loop_seconds = now_s() - tl0;
    r->t = lazy ? z->tf : z->t; r->x = lazy ? z->xf : z->x; r->th = z->th; r->c = z->c;
    if (lazy) { free(z->t); free(z->x); } else { free(z->tf); free(z->xf); }
    free(z->t_old); free(z->ba); free(z->bb); free(z->kctr);
    free(Q.key); free(Q.val); free(Q.index);
    free(z->g2ptr); free(z->g2idx); free(renew);
    return r;
}

zzo_run *zzo_spdmp(int64_t d,
                   const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                   const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                   double t0, const double *x0, const double *th0, double T, const double *c_in,
                   const uint64_t *seed, int adapt, double factor, int mode)
{
    return spdmp_impl(d, tg_colptr, tg_rowval, tg_nzval, h, bd_colptr, bd_rowval, bd_nzval, mu, t0, x0, th0, T, c_in,
                      seed, adapt, factor, mode, NULL, NULL);
}

/* spdmp with Z = ZigZag(Gamma, mu, sigma; lambdaref > 0): zzo_spdmp plus the refreshment clock (see `refr` above) */
zzo_run *zzo_spdmp_refresh(int64_t d,
                   const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                   const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                   const double *sigma, double lambdaref,
                   double t0, const double *x0, const double *th0, double T, const double *c_in,
                   const uint64_t *seed, int adapt, double factor, int mode)
{
    refr rf = { lambdaref, sigma };
    if ((mode & ZZO_LOCAL_BOUND) || !(lambdaref > 0.0)) {
        zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run)); r->d = 0; r->status = ZZO_E_ARG; return r;
    }
    return spdmp_impl(d, tg_colptr, tg_rowval, tg_nzval, h, bd_colptr, bd_rowval, bd_nzval, mu, t0, x0, th0, T, c_in,
                      seed, adapt, factor, mode, NULL, &rf);
}

/* spdmp(grad_phi_moving, t0, x0, th0, T, c, Zdrop, SelfMoving(), A, At, mu, y, ny, k; adapt, factor), scripts/logistic.jl:167.
 * (bd_*, bd_mu) = Zdrop.Gamma, Zdrop.mu; A is n x d, At its transpose (both Julia-layout CSC); mu_cv the control-variate point. */
zzo_run *zzo_spdmp_logistic(int64_t d, int64_t n,
                   const int64_t *a_colptr, const int64_t *a_rowval, const double *a_nzval,
                   const int64_t *at_colptr, const int64_t *at_rowval, const double *at_nzval,
                   const double *y, const double *ny, const double *mu_cv, double gamma0, int64_t k,
                   const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *bd_mu,
                   double t0, const double *x0, const double *th0, double T, const double *c_in,
                   const uint64_t *seed, int adapt, double factor, int mode)
{
    logit L; memset(&L, 0, sizeof L);
    L.n = n; L.A.colptr = a_colptr; L.A.rowval = a_rowval; L.A.nzval = a_nzval;
    L.At.colptr = at_colptr; L.At.rowval = at_rowval; L.At.nzval = at_nzval;
    L.y = y; L.ny = ny; L.mu = mu_cv; L.gamma0 = gamma0; L.k = k;
    L.grng.x = seed[0] ^ 0x9E3779B97F4A7C15ULL; L.grng.y = seed[1] ^ 0xD1342543DE82EF95ULL;
    if ((mode & (ZZO_LOCAL_BOUND | ZZO_GRAPH_ALL)) || k < 1) {
        zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run)); r->d = 0; r->status = ZZO_E_ARG; return r;
    }
    for (int64_t j = 0; j < d; ++j)
        if (a_colptr[j + 1] == a_colptr[j]) { /* rand(sampler) on an empty range throws upstream */
            zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run)); r->d = 0; r->status = ZZO_E_ARG; return r;
        }
    /* the target matrix arguments are unused with a logistic target: pass the bound matrix */
    return spdmp_impl(d, bd_colptr, bd_rowval, bd_nzval, NULL, bd_colptr, bd_rowval, bd_nzval, bd_mu, t0, x0, th0, T, c_in,
                      seed, adapt, factor, mode, &L, NULL);
}
double zzo_exp(double x) { return zz_exp(x); }

/* =====================================================================================================
 * Sticky ZigZag, sspdmp (src/ss_fact.jl:10-217): coordinates freeze when they hit 0 and thaw after an Exp(kappa_i) time.
 *   freezing_time            ss_fact.jl:10-16      ssmove_forward!   :25-45
 *   queue_time!              ss_fact.jl:54-66      sspdmp_inner!     :78-157     sspdmp  :159-217
 * Options restated: reversible = false, strong_upperbounds = false; `adapt` is refused (its reset of the global
 * counters, :134, depends on the global event order).  Every draw of the reference comes from Julia's GLOBAL RNG
 * (:96,112,130,179 and queue_time!'s default rng): mode seq uses one xoroshiro stream in that order, mode ctr the
 * per-coordinate streams (a coordinate draws when IT is queued, proposes, or freezes -- the thaw clock).
 * Returned acc[] counts accepted reflections per coordinate (the reference keeps only their sum).
 * ===================================================================================================== */
static double freezing_time(double x, double th)
{ /* ss_fact.jl:10-16 */
    if (th * x >= 0) return INFINITY;
    return -x / th;
}

typedef struct { ctx *z; heapq *Q; char *f; } sctx;

static void s_queue_time(sctx *S, int64_t j, double tj, double xj)
{ /* queue_time!(rng, Q, t, x, th, i, b, f, Z), ss_fact.jl:54-66 */
    ctx *z = S->z;
    double trefl = (z->mode & ZZO_STICKYZZ) ? zzo_poisson_time3(z->ba[j - 1], z->bb[j - 1], 0.01, draw(z, j))
                                            : o_poisson_time(z->ba[j - 1], z->bb[j - 1], draw(z, j));
    double tfreeze = freezing_time(xj, z->th[j - 1]);
    if (tfreeze <= trefl) { S->f[j - 1] = 1; h_set(S->Q, j, tj + tfreeze); }
    else { S->f[j - 1] = 0; h_set(S->Q, j, tj + trefl); }
}

static void s_move_nbhd(ctx *z, const int64_t *idx, int64_t n, double tp)
{ /* ssmove_forward!(G, i, ...), ss_fact.jl:38-45: frozen coordinates are not moved */
    for (int64_t q = 0; q < n; ++q) {
        int64_t k = idx[q] - 1;
        if (z->th[k] != 0.0) { z->x[k] = z->x[k] + z->th[k] * (tp - z->t[k]); z->t[k] = tp; }
    }
}

/* sspdmp(...; adapt = true, factor) (ss_fact.jl:132-136): an accepted proposal with l > lb multiplies c[i] by `factor` instead of
 * raising the error.  The reference also resets its two diagnostic counters (acc = num = 0, :134) at that moment, so what it
 * returns counts the proposals since the LAST adaptation in global event order; the contract returns the totals (the dynamics do
 * not depend on the counters). */
static zzo_run *sspdmp_impl(int64_t d,
                    const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                    const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                    double t0, const double *x0, const double *th0, double T, const double *c_in, const double *kappa,
                    const uint64_t *seed, int mode, int adapt, double factor);
zzo_run *zzo_sspdmp(int64_t d,
                    const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                    const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                    double t0, const double *x0, const double *th0, double T, const double *c_in, const double *kappa,
                    const uint64_t *seed, int mode)
{
    return sspdmp_impl(d, tg_colptr, tg_rowval, tg_nzval, h, bd_colptr, bd_rowval, bd_nzval, mu, t0, x0, th0, T, c_in, kappa, seed, mode, 0, 1.0);
}
zzo_run *zzo_sspdmp_adapt(int64_t d,
                    const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                    const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                    double t0, const double *x0, const double *th0, double T, const double *c_in, const double *kappa,
                    const uint64_t *seed, int mode, int adapt, double factor)
{
    return sspdmp_impl(d, tg_colptr, tg_rowval, tg_nzval, h, bd_colptr, bd_rowval, bd_nzval, mu, t0, x0, th0, T, c_in, kappa, seed, mode, adapt, factor);
}
static zzo_run *sspdmp_impl(int64_t d,
                    const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                    const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                    double t0, const double *x0, const double *th0, double T, const double *c_in, const double *kappa,
                    const uint64_t *seed, int mode, int adapt, double factor)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    ctx zs; ctx *z = &zs; memset(z, 0, sizeof(ctx));
    r->d = d; r->mode = mode; r->t0 = t0;
    z->d = d; z->mode = mode;
    z->tg.colptr = tg_colptr; z->tg.rowval = tg_rowval; z->tg.nzval = tg_nzval;
    z->bd.colptr = bd_colptr; z->bd.rowval = bd_rowval; z->bd.nzval = bd_nzval;
    z->h = h; z->mu = mu;
    size_t nb = (size_t)d * sizeof(double);
    z->t = (double *)malloc(nb); z->x = (double *)malloc(nb); z->th = (double *)malloc(nb);
    z->t_old = (double *)malloc(nb); z->ba = (double *)malloc(nb); z->bb = (double *)malloc(nb);
    z->c = (double *)malloc(nb); z->tf = (double *)malloc(nb); z->xf = (double *)malloc(nb);
    z->kctr = (uint32_t *)calloc((size_t)d, sizeof(uint32_t));
    double *thf = (double *)calloc((size_t)d, sizeof(double));
    char *f = (char *)calloc((size_t)d, 1);
    r->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    z->s0 = seed[0]; z->s1 = seed[1]; z->rng.x = seed[0]; z->rng.y = seed[1];
    const int lazy = (mode & ZZO_ARITH_LAZY) != 0;
    double tp = t0;
    for (int64_t k = 0; k < d; ++k) {
        z->t[k] = t0; z->t_old[k] = t0; z->x[k] = x0[k]; z->th[k] = th0[k]; z->c[k] = c_in[k];
        z->tf[k] = t0; z->xf[k] = x0[k];
    }
    if (!lazy) build_g2(z);
    heapq Q; Q.n = 0; Q.lex = (mode & ZZO_RNG_CTR) != 0;
    Q.key = (int64_t *)malloc(((size_t)d + 2) * sizeof(int64_t));
    Q.val = (double *)malloc(((size_t)d + 2) * sizeof(double));
    Q.index = (int64_t *)malloc(((size_t)d + 2) * sizeof(int64_t));
    sctx S = { z, &Q, f };
    /* ss_fact.jl:177-188 */
    if (mode & ZZO_STICKYZZ)   /* stickyzz.jl:198-206: a coordinate that starts at 0 starts frozen and keeps its speed for the thaw */
        for (int64_t k = 0; k < d; ++k) if (x0[k] == 0.0) { thf[k] = z->th[k]; z->th[k] = 0.0; }
    for (int64_t i = 1; i <= d; ++i) ab_zigzag(z, i, t0);
    for (int64_t i = 1; i <= d; ++i) {
        if ((mode & ZZO_STICKYZZ) && x0[i - 1] == 0.0) { f[i - 1] = 0; h_enqueue(&Q, i, t0 - zz_log(draw(z, i)) / kappa[i - 1]); continue; }
        double trefl = (mode & ZZO_STICKYZZ) ? zzo_poisson_time3(z->ba[i - 1], z->bb[i - 1], 0.01, draw(z, i))
                                             : o_poisson_time(z->ba[i - 1], z->bb[i - 1], draw(z, i));
        double tfreez = freezing_time(x0[i - 1], th0[i - 1]);
        if (trefl > tfreez) { f[i - 1] = 1; h_enqueue(&Q, i, t0 + tfreez); }
        else { f[i - 1] = 0; h_enqueue(&Q, i, t0 + trefl); }
    }
    int64_t num = 0;
    while (tp < T && r->status == ZZO_OK) { /* ss_fact.jl:202 */
        for (;;) {                           /* sspdmp_inner!, :82 */
            int64_t i = Q.key[1]; tp = Q.val[1];
            const int64_t *nbv = &z->bd.rowval[z->bd.colptr[i - 1] - 1];
            int64_t nnb = z->bd.colptr[i] - z->bd.colptr[i - 1];
            const int64_t *g2 = lazy ? NULL : &z->g2idx[z->g2ptr[i - 1]];
            int64_t ng2 = lazy ? 0 : z->g2ptr[i] - z->g2ptr[i - 1];
            double xi_now = lazy ? pos_at(z, i, tp) : 0.0;
            int is_event = 1;
            if (f[i - 1]) {                                         /* case 1: freeze, :87-107 */
                if (!lazy) { z->x[i - 1] = z->x[i - 1] + z->th[i - 1] * (tp - z->t[i - 1]); z->t[i - 1] = tp; xi_now = z->x[i - 1]; }
                if (fabs(xi_now) > 1e-8) { r->status = 9; break; }  /* error("x[i] = ... !~ 0"), :89-91 */
                double xz = -0.0 * z->th[i - 1];                      /* x[i] = -0*theta[i], :92 */
                if (lazy) { z->xf[i - 1] = xz; z->tf[i - 1] = tp; } else z->x[i - 1] = xz;
                thf[i - 1] = z->th[i - 1]; z->th[i - 1] = 0.0;      /* :93 */
                z->t_old[i - 1] = tp; f[i - 1] = 0;
                h_set(&Q, i, tp - zz_log(draw(z, i)) / kappa[i - 1]); /* :96 */
                if (!(mode & ZZO_STICKY_STRONG_UB)) {               /* :97 if !strong_upperbounds */
                if (!lazy) { s_move_nbhd(z, nbv, nnb, tp); s_move_nbhd(z, g2, ng2, tp); }
                for (int64_t q = 0; q < nnb; ++q) {                 /* :100-106 */
                    int64_t j = nbv[q];
                    if (z->th[j - 1] != 0) {
                        ab_zigzag(z, j, tp);
                        double tj = lazy ? tp : z->t[j - 1];
                        z->t_old[j - 1] = tj;
                        s_queue_time(&S, j, tj, lazy ? pos_at(z, j, tp) : z->x[j - 1]);
                    }
                }
                }
            } else if ((lazy ? z->xf[i - 1] : z->x[i - 1]) == 0 && z->th[i - 1] == 0) { /* case 2: thaw, :108-123 */
                if (lazy) z->tf[i - 1] = tp; else z->t[i - 1] = tp;
                z->th[i - 1] = thf[i - 1]; thf[i - 1] = 0.0;
                if (mode & ZZO_STICKY_REVERSIBLE) z->th[i - 1] *= (draw(z, i) < 0.5 ? -1.0 : 1.0);   /* :111-113 theta[i] *= rand((-1,1)) */
                z->t_old[i - 1] = tp;
                if (!lazy) { s_move_nbhd(z, nbv, nnb, tp); s_move_nbhd(z, g2, ng2, tp); }
                for (int64_t q = 0; q < nnb; ++q) {
                    int64_t j = nbv[q];
                    if (z->th[j - 1] != 0) {
                        ab_zigzag(z, j, tp);
                        double tj = lazy ? tp : z->t[j - 1];
                        z->t_old[j - 1] = tj;
                        s_queue_time(&S, j, tj, lazy ? pos_at(z, j, tp) : z->x[j - 1]);
                    }
                }
            } else {                                                /* case 3: proposal, :124-152 */
                if (!lazy) s_move_nbhd(z, nbv, nnb, tp);
                double gi = idot_x(z, &z->tg, i, tp);
                if (h) gi = gi - h[i - 1];
                double ti = lazy ? tp : z->t[i - 1];
                double l = zz_pos(gi * z->th[i - 1]);
                double lb = zz_pos(z->ba[i - 1] + z->bb[i - 1] * (ti - z->t_old[i - 1]));
                num += 1;
                if (draw(z, i) * lb < l) {
                    r->acc[i - 1] += 1;
                    if (l > lb) {                                           /* :132-136 */
                        if (!adapt) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                        z->c[i - 1] *= factor;
                    }
                    if (!lazy) s_move_nbhd(z, g2, ng2, tp);
                    if (lazy) { z->xf[i - 1] = pos_at(z, i, tp); z->tf[i - 1] = tp; }
                    z->th[i - 1] = -z->th[i - 1];
                    for (int64_t q = 0; q < nnb; ++q) {
                        int64_t j = nbv[q];
                        if (z->th[j - 1] != 0) {
                            ab_zigzag(z, j, tp);
                            double tj = lazy ? tp : z->t[j - 1];
                            z->t_old[j - 1] = tj;
                            s_queue_time(&S, j, tj, lazy ? pos_at(z, j, tp) : z->x[j - 1]);
                        }
                    }
                } else {
                    ab_zigzag(z, i, tp);
                    z->t_old[i - 1] = ti;
                    s_queue_time(&S, i, ti, lazy ? pos_at(z, i, tp) : z->x[i - 1]);
                    is_event = 0;
                }
            }
            if (is_event) {                                         /* :154 */
                push_event(r, tp, i, lazy ? z->xf[i - 1] : z->x[i - 1], z->th[i - 1]);
                break;
            }
        }
    }
    r->num = num;
    r->t = lazy ? z->tf : z->t; r->x = lazy ? z->xf : z->x; r->th = z->th; r->c = z->c;
    if (lazy) { free(z->t); free(z->x); } else { free(z->tf); free(z->xf); }
    free(z->t_old); free(z->ba); free(z->bb); free(z->kctr); free(thf); free(f);
    free(Q.key); free(Q.val); free(Q.index);
    free(z->g2ptr); free(z->g2idx);
    return r;
}

/* =====================================================================================================
 * Sparse sticky ZigZag, sparsestickyzz / sspdmp3 (src/sparsestickyzz.jl:3-41,119-172,192-276,280-401,405-422) -- config 4
 * as the reference runs it.  Restated faithfully for the CPU (mode seq only: one xoroshiro stream where the reference uses
 * `rng`, a second one where it uses Julia's global RNG, :11,296,312): there is no separate device kernel for it -- the
 * device serves config 4 with the sticky kernel of src/ss_fact.jl (zzo_sspdmp above is its contract), whose process has the
 * same law.  tests/test_sparse_sticky.py checks that claim statistically against this restatement.
 *   SparseState (:3-41): only unfrozen coordinates carry (t, x, theta); here dense arrays + an `active` flag.
 *   ab (:136-142): CONSTANT "strong" bound a = c + grad_i theta_i, b = 0, valid until t + s/c (s = 1 - rand^2 when adapt);
 *     a reflection reschedules ONLY the reflecting coordinate -- neighbours keep their bound until it expires (`renew`).
 *   queue_time! (:144-172): tau = min(expiry, t + poisson_time((a, 0, 0.01), r), hitting time of 0) (poissontime.jl:86-92,39-65).
 *   lambda (:129-134): (pos(grad_i theta_i), pos(a + b (t_i - t_ref))) -- the 0.01 floor is not in lb.
 *   thaw clock Q0 (:218,325,362,368): ONE clock of rate kappa * #frozen; a thaw picks a frozen coordinate uniformly (:295-298)
 *     and re-enters with velocity -1 + 2 p[i] (rule :sticky, :316) or a random sign (rule :reversible, :312).
 *   events (:249): accepted thaw (:326), hit (:363; the record of a deleted coordinate is (t', 0, 0), :20-26) and accepted
 *     reflection (:392).  clusteralpha = 1 (no cluster moves, :299-308,348-357).
 * The target is grad_i(u) = idot(G, i, u) - h_i over the sparse state (:42-51; frozen coordinates contribute x = 0).
 * ===================================================================================================== */
/* peek(q::PriorityQueues), src/morepriorityqueues.jl:33-42, with a LinearQueue head (:12-19: findmin = first minimum) and
 * a heap tail: the head wins only with a STRICTLY smaller time.  Returns the key; head keys are head_first + position. */
static int64_t pqs_peek(const double *head_vals, int64_t nhead, int64_t head_first, const heapq *tail, double *t)
{
    int64_t i1 = 0; double t1 = head_vals[0];
    for (int64_t k = 1; k < nhead; ++k) if (head_vals[k] < t1) { t1 = head_vals[k]; i1 = k; }
    if (tail->n == 0 || t1 < tail->val[1]) { *t = t1; return head_first + i1; }
    *t = tail->val[1];
    return tail->key[1];
}

/* Test hook (tests/test_oracle.py replays test/priority.jl:11-31 and random scripts through it): build the heap by
 * enqueueing (keys0, vals0) in order, then apply the updates `key -> val` (setindex!, priorityqueue.jl:95-105; a key of the
 * head range updates the LinearQueue) and record peek after each one (slot 0 = before any update). */
void zzo_queue_script(const double *head_vals_in, int64_t nhead, int64_t head_first, const int64_t *keys0, const double *vals0,
                      int64_t n0, int64_t maxkey, const int64_t *op_keys, const double *op_vals, int64_t nops,
                      int64_t *peek_keys, double *peek_vals)
{
    heapq Q; Q.n = 0; Q.lex = 0;
    Q.key = (int64_t *)malloc(((size_t)maxkey + 2) * sizeof(int64_t));
    Q.val = (double *)malloc(((size_t)maxkey + 2) * sizeof(double));
    Q.index = (int64_t *)calloc((size_t)maxkey + 2, sizeof(int64_t));
    double *head = (double *)malloc((size_t)(nhead > 0 ? nhead : 1) * sizeof(double));
    memcpy(head, head_vals_in, (size_t)nhead * sizeof(double));
    for (int64_t k = 0; k < n0; ++k) h_enqueue(&Q, keys0[k], vals0[k]);
    peek_keys[0] = pqs_peek(head, nhead, head_first, &Q, &peek_vals[0]);
    for (int64_t k = 0; k < nops; ++k) {
        if (op_keys[k] >= head_first && op_keys[k] < head_first + nhead) head[op_keys[k] - head_first] = op_vals[k];
        else h_set(&Q, op_keys[k], op_vals[k]);
        peek_keys[k + 1] = pqs_peek(head, nhead, head_first, &Q, &peek_vals[k + 1]);
    }
    free(Q.key); free(Q.val); free(Q.index); free(head);
}

static void hq_delete(heapq *q, int64_t key)
{ /* delete!(Q, i) of DataStructures.PriorityQueue: move the last entry into the hole and restore the heap */
    int64_t i = q->index[key];
    int64_t lk = q->key[q->n]; double lv = q->val[q->n];
    q->n -= 1;
    q->index[key] = 0;
    if (i > q->n) return;
    q->key[i] = lk; q->val[i] = lv; q->index[lk] = i;
    if (i > 1 && h_lt(q, lv, lk, q->val[i / 2], q->key[i / 2])) h_up(q, i); else h_down(q, i);
}

typedef struct {
    int64_t d; csc G; const double *h;
    char *active; double *t, *x, *th; char *p; double tmax;     /* SparseState: u, t', p */
    char *action; double *bt, *ba, *bb, *bexp;                   /* clocks[i] = (action, (t_ref, a, b, t_expire)) */
    heapq Q; double q0;                                          /* PriorityQueues(Q0, Q) */
    double c; int adapt; double mult, kappa; int64_t nactive;
    xoro rng, grng;
    int64_t acc, num;                                            /* AcceptanceDiagnostics */
} ssp;
enum { A_HIT = 0, A_REFLECT = 1, A_RENEW = 2 };

static double ss_grad(const ssp *S, int64_t i)
{ /* idot(A, j, u::SparseState), :42-51 */
    double s = 0.0;
    for (int64_t q = S->G.colptr[i - 1]; q < S->G.colptr[i]; ++q) {
        int64_t k = S->G.rowval[q - 1];
        s += S->G.nzval[q - 1] * (S->active[k - 1] ? S->x[k - 1] : 0.0);
    }
    return S->h ? s - S->h[i - 1] : s;
}
static void ss_move(ssp *S, int64_t j, double tp)
{ /* move_forward!(G, j, u, t', ::StickyFlow), :259-270: active neighbours only */
    for (int64_t q = S->G.colptr[j - 1]; q < S->G.colptr[j]; ++q) {
        int64_t i = S->G.rowval[q - 1];
        if (!S->active[i - 1]) continue;
        S->x[i - 1] = S->x[i - 1] + S->th[i - 1] * (tp - S->t[i - 1]);
        S->t[i - 1] = tp;
    }
    if (S->tmax < tp) S->tmax = tp;
}
static void ss_ab(ssp *S, int64_t i, double gi)
{ /* ab(rng, su, i, u, grad_i, flow), :136-142 */
    double a = S->c + gi * S->th[i - 1];
    double s = 1.0;
    if (S->adapt) { double r = xoro_rand(&S->rng); s = 1.0 - r * r; }
    S->bt[i - 1] = S->t[i - 1]; S->ba[i - 1] = a; S->bb[i - 1] = 0.0; S->bexp[i - 1] = S->t[i - 1] + s / S->c;
}
static void ss_queue_time(ssp *S, int64_t i, int enqueue)
{ /* queue_time!, :144-172 */
    double t = S->t[i - 1], x = S->x[i - 1], v = S->th[i - 1];
    double trefresh = S->bexp[i - 1];
    double dt = t - S->bt[i - 1];                                 /* poisson_time(t, b::Tuple, r), poissontime.jl:86-92 */
    double trefl = t + zzo_poisson_time3(S->ba[i - 1] + dt * S->bb[i - 1], S->bb[i - 1], 0.01, xoro_rand(&S->rng));
    double thit = (v * (x - 0.0) >= 0) ? INFINITY : t - (x - 0.0) / v;   /* hitting_time, :119-126 */
    double tau = trefresh < trefl ? trefresh : trefl;
    if (thit < tau) tau = thit;
    if (enqueue) h_enqueue(&S->Q, i, tau); else h_set(&S->Q, i, tau);
    S->action[i - 1] = (thit == tau) ? A_HIT : (trefl == tau) ? A_REFLECT : A_RENEW;
}
static double ss_randexp(xoro *r) { double u = xoro_rand(r); return -zz_log(u > 0 ? u : 0x1p-53); }

/* rule: 0 = :sticky, 1 = :reversible.  th0 gives the initial velocities of the coordinates with x0 != 0 (the reference draws
 * them from the global RNG, :11).  Returns events; acc[0] = accepted reflections, num = proposals (AcceptanceDiagnostics). */
zzo_run *zzo_sparsestickyzz(int64_t d, const int64_t *g_colptr, const int64_t *g_rowval, const double *g_nzval, const double *h,
                            const double *x0, const double *th0, double T, double c, double kappa, int rule, int adapt,
                            double multiplier, const uint64_t *seed)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    ssp Ss; ssp *S = &Ss; memset(S, 0, sizeof Ss);
    r->d = d; S->d = d; S->G.colptr = g_colptr; S->G.rowval = g_rowval; S->G.nzval = g_nzval; S->h = h;
    size_t nb = (size_t)d * sizeof(double);
    S->active = (char *)calloc((size_t)d, 1); S->p = (char *)malloc((size_t)d); S->action = (char *)calloc((size_t)d, 1);
    S->t = (double *)calloc((size_t)d, 8); S->x = (double *)calloc((size_t)d, 8); S->th = (double *)calloc((size_t)d, 8);
    S->bt = (double *)calloc((size_t)d, 8); S->ba = (double *)calloc((size_t)d, 8); S->bb = (double *)calloc((size_t)d, 8);
    S->bexp = (double *)calloc((size_t)d, 8);
    r->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    r->c = (double *)malloc(nb);
    S->c = c; S->adapt = adapt; S->mult = multiplier; S->kappa = kappa;
    S->rng.x = seed[0]; S->rng.y = seed[1];
    S->grng.x = seed[0] ^ 0x9E3779B97F4A7C15ULL; S->grng.y = seed[1] ^ 0xD1342543DE82EF95ULL;
    S->Q.n = 0; S->Q.lex = 0;
    S->Q.key = (int64_t *)malloc(((size_t)d + 2) * sizeof(int64_t));
    S->Q.val = (double *)malloc(((size_t)d + 2) * sizeof(double));
    S->Q.index = (int64_t *)calloc((size_t)d + 2, sizeof(int64_t));
    double tp = 0.0;                                             /* u.t' = 0 (:11) */
    for (int64_t k = 0; k < d; ++k) {                            /* sparsestickystate (:10-12), u.p (:196-201) */
        S->p[k] = 1;
        if (x0[k] != 0) { S->active[k] = 1; S->x[k] = x0[k]; S->th[k] = th0[k]; S->p[k] = th0[k] > 0; S->nactive++; }
    }
    S->q0 = tp + ss_randexp(&S->rng) / (kappa * (double)(d - S->nactive));   /* :218 */
    for (int64_t i = 1; i <= d; ++i) {                           /* :224-228 */
        if (!S->active[i - 1]) continue;
        ss_ab(S, i, ss_grad(S, i));
        ss_queue_time(S, i, 1);
    }
    const double tl0 = now_s();
    while (tp < T && r->status == ZZO_OK) {                      /* sparsesticky_main, :246 */
        int64_t ev_i = 0;
        for (;;) {                                               /* sparsestickyzz_inner!, :280-401 */
            int64_t i = pqs_peek(&S->q0, 1, 0, &S->Q, &tp); /* peek(Qs), morepriorityqueues.jl:33-42; key 0 = thaw clock */
            if (i == 0) {                                        /* thaw, :291-329 */
                if (!(tp < INFINITY)) { r->status = 9; break; }
                do { i = 1 + (int64_t)(xoro_rand(&S->grng) * (double)d); if (i > d) i = d; } while (S->active[i - 1]);
                double vi = rule == 1 ? (xoro_rand(&S->grng) < 0.5 ? -1.0 : 1.0) : -1.0 + 2.0 * (double)S->p[i - 1];
                S->active[i - 1] = 1; S->nactive++;              /* insert!(u, i, (t', barriers.x, vi)) */
                S->t[i - 1] = tp; S->x[i - 1] = 0.0; S->th[i - 1] = vi;
                if (S->tmax < tp) S->tmax = tp;
                ss_move(S, i, tp);
                ss_ab(S, i, ss_grad(S, i));
                ss_queue_time(S, i, 1);
                S->q0 = tp + ss_randexp(&S->rng) / (kappa * (double)(d - S->nactive));
                ev_i = i;
                break;
            }
            if (S->action[i - 1] == A_RENEW) {                   /* :330-340 */
                ss_move(S, i, tp);
                double gi = ss_grad(S, i);
                double l = zz_pos(gi * S->th[i - 1]);
                double lb = zz_pos(S->ba[i - 1] + S->bb[i - 1] * (S->t[i - 1] - S->bt[i - 1]));
                if (l > lb) {
                    if (!adapt) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                    S->acc = S->num = 0; S->c *= S->mult;
                }
                ss_ab(S, i, gi);
                ss_queue_time(S, i, 0);
                continue;
            }
            if (S->action[i - 1] == A_HIT) {                     /* :341-371 */
                ss_move(S, i, tp);
                if (fabs(S->x[i - 1]) > 1e-7) { r->status = 9; break; }
                S->active[i - 1] = 0; S->nactive--;              /* delete!(u.u, i); delete!(clocks, i); delete!(Q, i) */
                S->x[i - 1] = 0.0; S->th[i - 1] = 0.0;
                hq_delete(&S->Q, i);
                S->q0 = tp + ss_randexp(&S->rng) / (kappa * (double)(d - S->nactive));
                ev_i = i;
                break;
            }
            ss_move(S, i, tp);                                   /* reflection time or event of the bound, :372-399 */
            double gi = ss_grad(S, i);
            double l = zz_pos(gi * S->th[i - 1]);
            double lb = zz_pos(S->ba[i - 1] + S->bb[i - 1] * (S->t[i - 1] - S->bt[i - 1]));
            if (xoro_rand(&S->rng) * lb < l) {
                S->acc++; S->num++; r->acc[i - 1]++;
                if (l > lb) {
                    if (!adapt) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                    S->acc = S->num = 0; S->c *= S->mult;
                }
                S->th[i - 1] = -S->th[i - 1];                    /* reflect!, :271-275 */
                if (rule == 0) S->p[i - 1] = S->th[i - 1] > 0;
                ss_ab(S, i, ss_grad(S, i));
                ss_queue_time(S, i, 0);
                ev_i = i;
                break;
            }
            S->num++;
            ss_ab(S, i, gi);
            ss_queue_time(S, i, 0);
        }
        if (r->status != ZZO_OK) break;
        /* push!(Xi, event(i, u, flow)), :249: (t_i, i, x_i, theta_i), (t', 0, 0) for a coordinate that was just deleted */
        if (S->active[ev_i - 1]) push_event(r, S->t[ev_i - 1], ev_i, S->x[ev_i - 1], S->th[ev_i - 1]);
        else push_event(r, S->tmax, ev_i, 0.0, 0.0);
    }
    r->loop_seconds = now_s() - tl0;
    r->num = S->num;
    for (int64_t k = 0; k < d; ++k) { r->c[k] = S->c; if (!S->active[k]) { S->t[k] = S->tmax; S->x[k] = 0.0; S->th[k] = 0.0; } }
    r->t = S->t; r->x = S->x; r->th = S->th;
    free(S->active); free(S->p); free(S->action); free(S->bt); free(S->ba); free(S->bb); free(S->bexp);
    free(S->Q.key); free(S->Q.val); free(S->Q.index);
    return r;
}

/* ---- the same sampler in the parity arithmetic (mode ctr|lazy): the contract a device kernel for the strong-bound sticky
 * sampler reproduces bit for bit (zz_run_kernel_csr_strong, since round 2).  Differences to the
 * faithful restatement above, all equal in law:
 *   * per-coordinate counter streams u(i, k) (zz_math.h): coordinate i draws, in the order of ITS OWN events, the thaw
 *     waiting time when it freezes (and at t = 0 if it starts frozen), [rule :reversible: the sign at a thaw], the time of
 *     its next proposal whenever it is queued, and the thinning uniform of a proposal;
 *   * one Exp(kappa) thaw clock PER frozen coordinate instead of one clock of rate kappa * #frozen with a uniform pick
 *     (:218,291-298) -- the superposition of the former is the latter;
 *   * flip-anchored positions x_k(s) = xf_k + th_k (s - tf_k), rewritten only when the velocity of k changes (thaw, hit,
 *     accepted reflection); a frozen coordinate sits at 0; ties in time are broken by coordinate;
 *   * adapt is refused (the adapted bound is ONE global c, :336,383: its value depends on the global event order).
 * A reflection still reschedules nobody but the reflecting coordinate: in the windowed scheme of the device the timeline of
 * a coordinate then has OWN items only, and neighbours enter through the positions read at those items. */
/* The same contract with a bound constant c[i] and a thaw rate kappa[i] PER coordinate, a start time t0 and rule 2 = "a thawed
 * coordinate continues with the velocity it had when it froze" -- the process of asynchzz / sspdmp4 (src/asynchzz.jl:20-37,
 * 150-245: StrongUpperBounds ab :20-28 with per-coordinate c, queue_time! :42-52 = min(bound expiry, proposal at rate
 * 0.01 + a^+, hitting time of 0), thaw clock per coordinate :188, rule :sticky :206-213, accepted reflection :217-231, renewal
 * :192-203), whose law does not depend on its schedule (local minima of a PartialQueue processed by threads); only
 * adapt = false.  The reference's own end of a run depends on that schedule (it finishes the round in which the earliest time
 * passes T, :138-147); the contract ends like sspdmp3: after the first event at or after T.
 * Rules 0 / 1 with all c, all kappa equal and t0 = 0 are zzo_sparsestickyzz_ctr. */
zzo_run *zzo_strongsticky_ctr(int64_t d, const int64_t *g_colptr, const int64_t *g_rowval, const double *g_nzval, const double *h,
                              double t0, const double *x0, const double *th0, double T, const double *cv, const double *kv, int rule,
                              const uint64_t *seed)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    csc G = { g_colptr, g_rowval, g_nzval };
    size_t nb = (size_t)d * sizeof(double);
    r->d = d; r->t0 = t0;
    char *active = (char *)calloc((size_t)d, 1), *action = (char *)calloc((size_t)d, 1);
    double *psign = (double *)malloc(nb);   /* rules 0 / 1: remembered sign as +-1; rule 2: the velocity to continue with */
    double *tf = (double *)calloc((size_t)d, 8), *xf = (double *)calloc((size_t)d, 8), *th = (double *)calloc((size_t)d, 8);
    double *bt = (double *)calloc((size_t)d, 8), *ba = (double *)calloc((size_t)d, 8), *bexp = (double *)calloc((size_t)d, 8);
    uint32_t *kc = (uint32_t *)calloc((size_t)d, sizeof(uint32_t));
    r->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    r->c = (double *)malloc(nb);
    heapq Q; Q.n = 0; Q.lex = 1;
    Q.key = (int64_t *)malloc(((size_t)d + 2) * sizeof(int64_t));
    Q.val = (double *)malloc(((size_t)d + 2) * sizeof(double));
    Q.index = (int64_t *)calloc((size_t)d + 2, sizeof(int64_t));
    enum { C_HIT = 0, C_REFLECT = 1, C_RENEW = 2, C_THAW = 3 };
#define SC_U(i) zz_u01(seed[0], seed[1], (uint64_t)((i) - 1), kc[(i) - 1]++)
#define SC_POS(k, s) (active[(k) - 1] ? xf[(k) - 1] + th[(k) - 1] * ((s) - tf[(k) - 1]) : 0.0)
#define SC_GRAD(i, s, out) do { double acc_ = 0.0; \
        for (int64_t q_ = G.colptr[(i) - 1]; q_ < G.colptr[(i)]; ++q_) { int64_t k_ = G.rowval[q_ - 1]; acc_ += G.nzval[q_ - 1] * SC_POS(k_, (s)); } \
        (out) = h ? acc_ - h[(i) - 1] : acc_; } while (0)
    /* ab (:136-142) at time s, then queue_time! (:144-172) */
#define SC_QUEUE(i, s, gi) do { \
        ba[(i) - 1] = cv[(i) - 1] + (gi) * th[(i) - 1]; bt[(i) - 1] = (s); bexp[(i) - 1] = (s) + 1.0 / cv[(i) - 1]; \
        double xs_ = SC_POS((i), (s)); \
        double trefl_ = (s) + zzo_poisson_time3(ba[(i) - 1], 0.0, 0.01, SC_U(i)); \
        double thit_ = (th[(i) - 1] * xs_ >= 0) ? INFINITY : (s) - xs_ / th[(i) - 1]; \
        double tau_ = bexp[(i) - 1] < trefl_ ? bexp[(i) - 1] : trefl_; \
        if (thit_ < tau_) tau_ = thit_; \
        action[(i) - 1] = (thit_ == tau_) ? C_HIT : (trefl_ == tau_) ? C_REFLECT : C_RENEW; \
        if (Q.index[(i)]) h_set(&Q, (i), tau_); else h_enqueue(&Q, (i), tau_); } while (0)
    for (int64_t k = 0; k < d; ++k) {
        psign[k] = rule == 2 ? th0[k] : 1.0;
        tf[k] = t0;
        if (x0[k] != 0) { active[k] = 1; xf[k] = x0[k]; th[k] = th0[k]; psign[k] = rule == 2 ? th0[k] : (th0[k] > 0 ? 1.0 : -1.0); }
    }
    for (int64_t i = 1; i <= d; ++i) {
        if (active[i - 1]) { double gi; SC_GRAD(i, t0, gi); SC_QUEUE(i, t0, gi); }
        else { action[i - 1] = C_THAW; h_enqueue(&Q, i, t0 - zz_log(SC_U(i)) / kv[i - 1]); }
    }
    double tp = t0; int64_t num = 0;
    const double tl0 = now_s();
    while (tp < T && r->status == ZZO_OK) {
        for (;;) {
            int64_t i = Q.key[1]; tp = Q.val[1];
            double gi;
            if (action[i - 1] == C_THAW) {
                double vi = rule == 1 ? (SC_U(i) < 0.5 ? -1.0 : 1.0) : psign[i - 1];
                active[i - 1] = 1; tf[i - 1] = tp; th[i - 1] = vi;   /* xf stays the (signed) zero of the hit / of x0 */
                SC_GRAD(i, tp, gi); SC_QUEUE(i, tp, gi);
                push_event(r, tp, i, xf[i - 1], vi);
                break;
            }
            if (action[i - 1] == C_HIT) {
                if (fabs(SC_POS(i, tp)) > 1e-7) { r->status = 9; break; }
                /* the frozen record is x = -0*theta, theta = 0 -- the signed zero the sticky kernels commit (ss_fact.jl:92); the
                   reference's record of a deleted coordinate is (t', 0.0, 0.0) (:20-26): equal up to the sign of zero */
                active[i - 1] = 0; tf[i - 1] = tp; xf[i - 1] = -0.0 * th[i - 1]; th[i - 1] = 0.0;
                action[i - 1] = C_THAW;
                h_set(&Q, i, tp - zz_log(SC_U(i)) / kv[i - 1]);
                push_event(r, tp, i, xf[i - 1], 0.0);
                break;
            }
            SC_GRAD(i, tp, gi);
            double l = zz_pos(gi * th[i - 1]), lb = zz_pos(ba[i - 1]);
            if (action[i - 1] == C_RENEW) {
                if (l > lb) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                SC_QUEUE(i, tp, gi);
                continue;
            }
            num++;
            if (SC_U(i) * lb < l) {
                r->acc[i - 1]++;
                if (l > lb) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                xf[i - 1] = SC_POS(i, tp); tf[i - 1] = tp; th[i - 1] = -th[i - 1];
                if (rule == 0) psign[i - 1] = th[i - 1] > 0 ? 1.0 : -1.0;
                else if (rule == 2) psign[i - 1] = th[i - 1];
                SC_GRAD(i, tp, gi); SC_QUEUE(i, tp, gi);
                push_event(r, tp, i, xf[i - 1], th[i - 1]);
                break;
            }
            SC_QUEUE(i, tp, gi);
        }
    }
#undef SC_U
#undef SC_POS
#undef SC_GRAD
#undef SC_QUEUE
    r->loop_seconds = now_s() - tl0;
    r->num = num;
    for (int64_t k = 0; k < d; ++k) r->c[k] = cv[k];
    r->t = tf; r->x = xf; r->th = th;
    free(active); free(psign); free(action); free(bt); free(ba); free(bexp); free(kc);
    free(Q.key); free(Q.val); free(Q.index);
    return r;
}

zzo_run *zzo_sparsestickyzz_ctr(int64_t d, const int64_t *g_colptr, const int64_t *g_rowval, const double *g_nzval, const double *h,
                                const double *x0, const double *th0, double T, double c, double kappa, int rule,
                                const uint64_t *seed)
{
    double *cv = (double *)malloc((size_t)d * 8), *kv = (double *)malloc((size_t)d * 8);
    for (int64_t k = 0; k < d; ++k) { cv[k] = c; kv[k] = kappa; }
    zzo_run *r = zzo_strongsticky_ctr(d, g_colptr, g_rowval, g_nzval, h, 0.0, x0, th0, T, cv, kv, rule, seed);
    free(cv); free(kv);
    return r;
}

/* =====================================================================================================
 * Factorised Boomerang in spdmp / pdmp (src/sfact.jl:29-48,73-145,162-212 with F::FactBoomerang):
 *   flow                     sfact.jl:29-38 / dynamics.jl:29-36 (rotation of (x - mu, theta), `sincos` -> zz_sincos)
 *   lambda                   fact_samplers.jl:37-39     pos((grad_i - (x_i - mu_i) Gamma_ii) theta_i)
 *   ab                       fact_samplers.jl:58-65     a = c_i sqrt(x_i^2 + th_i^2) z + (x_i^2 + th_i^2) Gamma_ii, b = 0
 *   refreshment              sfact.jl:78-114, hasrefresh fact_samplers.jl:18, waiting_time_ref dynamics.jl:100-101
 * RNG seq  : ONE stream in the reference's call order; the reference draws the refreshed coordinate TWICE from Julia's
 *            global RNG (sfact.jl:80,84: the neighbourhood of the first pick is moved, the velocity of the second is
 *            refreshed at ITS stale local time) -- reproduced literally with arithmetic inplace.
 * RNG ctr  : (GPU parity contract) the single clock of rate lambda_ref picking a uniform coordinate is replaced by the
 *            equal-in-law superposition of d independent clocks of rate lambda_ref/d, one per coordinate, each drawing
 *            from its coordinate's counter stream; the refreshment is applied at the event time.  Draw roles per
 *            coordinate: first proposal (k=0), first refreshment time (k=1); a proposal consumes (thinning, reschedule);
 *            a neighbour's event consumes (reschedule); a refreshment consumes (2 for the normal, next refreshment time)
 *            and then the reschedule of the coordinate itself.
 * randn is Box-Muller on two uniforms (Julia: ziggurat; equal in law).
 * ===================================================================================================== */
typedef struct { const double *mu, *sigma; double lref, rho, rhobar; double *diag; } boomp;

static void b_at(const ctx *z, const boomp *B, int64_t k, double s, double *x, double *th)
{ /* lazy: anchor (tf, xf, th) of coordinate k (1-based) rotated to time s */
    zz_boom_at(z->tf[k - 1], z->xf[k - 1], z->th[k - 1], B->mu[k - 1], s, x, th);
}
static void b_move(ctx *z, const boomp *B, const int64_t *idx, int64_t n, double tp)
{ /* smove_forward!(G, i, ..., B), sfact.jl:29-38 (inplace) */
    for (int64_t q = 0; q < n; ++q) {
        int64_t k = idx[q] - 1;
        double sn, cs; zz_sincos(tp - z->t[k], &sn, &cs);
        double xm = z->x[k] - B->mu[k];
        double xn = xm * cs + z->th[k] * sn + B->mu[k];
        double tn = -xm * sn + z->th[k] * cs;
        z->t[k] = tp; z->x[k] = xn; z->th[k] = tn;
    }
}
static void b_move_all(ctx *z, const boomp *B, double tp)
{ /* sfact.jl:40-48 */
    for (int64_t k = 1; k <= z->d; ++k) b_move(z, B, &k, 1, tp);
}
/* ab(G,i,x,th,c,Z::FactBoomerang), fact_samplers.jl:58-65; positions/velocities at time s (lazy) or as stored (inplace) */
static void ab_boom(ctx *z, const boomp *B, int64_t i, double s)
{
    const int lazy = (z->mode & ZZO_ARITH_LAZY) != 0;
    double sum = 0.0, xi = 0.0, thi = 0.0; int first = 1, found = 0;
    for (int64_t p = z->bd.colptr[i - 1]; p < z->bd.colptr[i]; ++p) {   /* sum(... for j in nhd): left fold in storage order */
        int64_t j = z->bd.rowval[p - 1];
        double xj, thj;
        if (lazy) b_at(z, B, j, s, &xj, &thj); else { xj = z->x[j - 1]; thj = z->th[j - 1]; }
        double term = (xj - B->mu[j - 1]) * (xj - B->mu[j - 1]) + thj * thj;
        sum = first ? term : sum + term; first = 0;
        if (j == i) { xi = xj; thi = thj; found = 1; }
    }
    if (!found) { if (lazy) b_at(z, B, i, s, &xi, &thi); else { xi = z->x[i - 1]; thi = z->th[i - 1]; } }
    double zz = sqrt(sum);
    double z2 = xi * xi + thi * thi;
    z->ba[i - 1] = z->c[i - 1] * sqrt(z2) * zz + z2 * B->diag[i - 1];
    z->bb[i - 1] = 0.0;
}
static double b_grad(ctx *z, const boomp *B, int64_t i, double s, const double *h)
{ /* the user closure idot(Gamma, i, x) [- h_i], common.jl:16-24 */
    double acc = 0.0;
    if (!(z->mode & ZZO_ARITH_LAZY)) acc = idot(&z->tg, i, z->x);
    else for (int64_t p = z->tg.colptr[i - 1]; p < z->tg.colptr[i]; ++p) {
        double xj, thj; b_at(z, B, z->tg.rowval[p - 1], s, &xj, &thj);
        acc += z->tg.nzval[p - 1] * xj;
    }
    if (h) acc = acc - h[i - 1];
    return acc;
}

zzo_run *zzo_spdmp_boom(int64_t d,
                        const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                        const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                        const double *sigma, double lambdaref, double rho,
                        double t0, const double *x0, const double *th0, double T, const double *c_in,
                        const uint64_t *seed, int adapt, double factor, int mode)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    ctx zs; ctx *z = &zs; memset(z, 0, sizeof(ctx));
    r->d = d; r->mode = mode; r->t0 = t0;
    z->d = d; z->mode = mode;
    z->tg.colptr = tg_colptr; z->tg.rowval = tg_rowval; z->tg.nzval = tg_nzval;
    z->bd.colptr = bd_colptr; z->bd.rowval = bd_rowval; z->bd.nzval = bd_nzval;
    z->h = h; z->mu = mu;
    size_t nb = (size_t)d * sizeof(double);
    z->t = (double *)malloc(nb); z->x = (double *)malloc(nb); z->th = (double *)malloc(nb);
    z->t_old = (double *)malloc(nb); z->ba = (double *)malloc(nb); z->bb = (double *)malloc(nb);
    z->c = (double *)malloc(nb); z->tf = (double *)malloc(nb); z->xf = (double *)malloc(nb);
    z->kctr = (uint32_t *)calloc((size_t)d, sizeof(uint32_t));
    r->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    z->s0 = seed[0]; z->s1 = seed[1]; z->rng.x = seed[0]; z->rng.y = seed[1];
    const int lazy = (mode & ZZO_ARITH_LAZY) != 0, all = (mode & ZZO_GRAPH_ALL) != 0, ctr = (mode & ZZO_RNG_CTR) != 0;
    boomp B; B.mu = mu; B.sigma = sigma; B.lref = lambdaref; B.rho = rho; B.rhobar = sqrt(1 - rho * rho);
    B.diag = (double *)calloc((size_t)d, sizeof(double));
    for (int64_t i = 1; i <= d; ++i)  /* Z.Gamma[i,i] (sparse getindex: 0 when not stored) */
        for (int64_t p = bd_colptr[i - 1]; p < bd_colptr[i]; ++p)
            if (bd_rowval[p - 1] == i) B.diag[i - 1] = bd_nzval[p - 1];
    double tp = t0;
    for (int64_t k = 0; k < d; ++k) {
        z->t[k] = t0; z->t_old[k] = t0; z->x[k] = x0[k]; z->th[k] = th0[k]; z->c[k] = c_in[k];
        z->tf[k] = t0; z->xf[k] = x0[k];
    }
    if (!all && !lazy) build_g2(z);
    /* keys 1..d: reflections; seq: key d+1 = the refreshment clock (sfact.jl:188-190); ctr: keys d+1..2d = one clock per coordinate */
    heapq Q; Q.n = 0; Q.lex = ctr;
    Q.key = (int64_t *)malloc((2 * (size_t)d + 2) * sizeof(int64_t));
    Q.val = (double *)malloc((2 * (size_t)d + 2) * sizeof(double));
    Q.index = (int64_t *)malloc((2 * (size_t)d + 2) * sizeof(int64_t));
    const double lam1 = ctr ? lambdaref / (double)d : lambdaref;
    for (int64_t i = 1; i <= d; ++i) ab_boom(z, &B, i, t0);
    for (int64_t i = 1; i <= d; ++i) h_enqueue(&Q, i, o_poisson_time(z->ba[i - 1], z->bb[i - 1], draw(z, i))); /* sfact.jl:186, no + t0 */
    if (ctr) { for (int64_t i = 1; i <= d; ++i) h_enqueue(&Q, d + i, -zz_log(draw(z, i)) / lam1); }
    else h_enqueue(&Q, d + 1, -zz_log(xoro_rand(&z->rng)) / lam1);   /* waiting_time_ref(rng, F), no + t0 either */

    int64_t num = 0;
    while (tp < T && r->status == ZZO_OK) {
        for (;;) {
            int64_t i = Q.key[1]; tp = Q.val[1];
            const int refresh = i > d;
            int64_t qkey = i;
            if (refresh) i = ctr ? i - d : (int64_t)(xoro_rand(&z->rng) * (double)d) + 1;          /* :80 rand(1:n) */
            const int64_t *nbv = &z->bd.rowval[z->bd.colptr[i - 1] - 1];
            int64_t nnb = z->bd.colptr[i] - z->bd.colptr[i - 1];
            if (!lazy) { if (all) b_move_all(z, &B, tp); else b_move(z, &B, nbv, nnb, tp); }       /* :82 */
            if (refresh) {
                if (!ctr) i = (int64_t)(xoro_rand(&z->rng) * (double)d) + 1;                       /* :84 second pick */
                nbv = &z->bd.rowval[z->bd.colptr[i - 1] - 1]; nnb = z->bd.colptr[i] - z->bd.colptr[i - 1];
                if (!lazy && !all) b_move(z, &B, &z->g2idx[z->g2ptr[i - 1]], z->g2ptr[i] - z->g2ptr[i - 1], tp); /* :85 */
                double u1 = draw(z, i), u2 = draw(z, i);
                double xi, thi;
                if (lazy) { b_at(z, &B, i, tp, &xi, &thi); z->xf[i - 1] = xi; z->tf[i - 1] = tp; }
                else { xi = z->x[i - 1]; thi = z->th[i - 1]; }
                z->th[i - 1] = B.rho * thi + B.rhobar * B.sigma[i - 1] * zz_randn(u1, u2);         /* :102 */
                h_set(&Q, qkey, tp - zz_log(ctr ? draw(z, i) : xoro_rand(&z->rng)) / lam1);        /* :108 */
                for (int64_t q = 0; q < nnb; ++q) {                                                /* :110-114 */
                    int64_t j = nbv[q];
                    ab_boom(z, &B, j, tp);
                    double tj = lazy ? tp : z->t[j - 1];
                    z->t_old[j - 1] = tj;
                    h_set(&Q, j, tj + o_poisson_time(z->ba[j - 1], z->bb[j - 1], draw(z, j)));
                }
                push_event(r, lazy ? tp : z->t[i - 1], i, lazy ? z->xf[i - 1] : z->x[i - 1], z->th[i - 1]);
                break;
            }
            double xi, thi;
            if (lazy) b_at(z, &B, i, tp, &xi, &thi); else { xi = z->x[i - 1]; thi = z->th[i - 1]; }
            double gi = b_grad(z, &B, i, tp, h);                                                   /* :118 */
            double ti = lazy ? tp : z->t[i - 1];
            double l = zz_pos((gi - (xi - B.mu[i - 1]) * B.diag[i - 1]) * thi);                    /* fact_samplers.jl:37-39 */
            double lb = zz_pos(z->ba[i - 1] + z->bb[i - 1] * (ti - z->t_old[i - 1]));
            num += 1;
            if (draw(z, i) * lb < l) {
                r->acc[i - 1] += 1;
                if (l >= lb) {
                    if (!adapt) { r->status = ZZO_E_BOUND; r->err_i = i; r->err_t = tp; r->err_l = l; r->err_lb = lb; break; }
                    z->c[i - 1] *= factor;
                }
                if (!lazy && !all) b_move(z, &B, &z->g2idx[z->g2ptr[i - 1]], z->g2ptr[i] - z->g2ptr[i - 1], tp);
                if (lazy) { z->xf[i - 1] = xi; z->tf[i - 1] = tp; }
                z->th[i - 1] = -thi;                                                               /* dynamics.jl:46-49 */
                for (int64_t q = 0; q < nnb; ++q) {
                    int64_t j = nbv[q];
                    ab_boom(z, &B, j, tp);
                    double tj = lazy ? tp : z->t[j - 1];
                    z->t_old[j - 1] = tj;
                    h_set(&Q, j, tj + o_poisson_time(z->ba[j - 1], z->bb[j - 1], draw(z, j)));
                }
                push_event(r, ti, i, lazy ? z->xf[i - 1] : z->x[i - 1], z->th[i - 1]);
                break;
            } else {
                ab_boom(z, &B, i, tp);
                z->t_old[i - 1] = ti;
                h_set(&Q, i, ti + o_poisson_time(z->ba[i - 1], z->bb[i - 1], draw(z, i)));
            }
        }
    }
    r->num = num;
    r->t = lazy ? z->tf : z->t; r->x = lazy ? z->xf : z->x; r->th = z->th; r->c = z->c;
    if (lazy) { free(z->t); free(z->x); } else { free(z->tf); free(z->xf); }
    free(z->t_old); free(z->ba); free(z->bb); free(z->kctr); free(B.diag);
    free(Q.key); free(Q.val); free(Q.index);
    free(z->g2ptr); free(z->g2idx);
    return r;
}

void zzo_sincos(double x, double *s, double *c) { zz_sincos(x, s, c); }
double zzo_randn(double u1, double u2) { return zz_randn(u1, u2); }

/* =====================================================================================================
 * parallel_spdmp (src/parallel.jl:5-253): the reference's experimental MULTITHREADED local ZigZag -- the CPU baseline
 * "with all the host threads it can use" of bench.py.  Restated with pthreads:
 *   Partition                 parallel.jl:5-30       contiguous chunks of d / K coordinates, one heap per chunk
 *   parallel_innermost!       parallel.jl:34-61      one proposal (in-place moves, thinning, reschedule inside the chunk)
 *   parallel_spdmp_inner!     parallel.jl:63-104     worker: runs the events of its chunk while the next one is an INNER
 *                                                    coordinate (whole neighbourhood in the chunk) and not later than
 *                                                    t_next; otherwise reports (i, t') and sleeps
 *   parallel_spdmp_outer!     parallel.jl:182-253    when every worker sleeps: absorb their events, process the reported
 *                                                    events in time order if the neighbouring chunks have reached that
 *                                                    time, wake the workers that are not waiting for a neighbour
 *   parallel_spdmp            parallel.jl:106-151    setup; requires that the BOUND matrix Z.Gamma has no entries across
 *                                                    chunks ("Upper bounds may not depend across chunks", :126-129) while
 *                                                    the target's gradient and the neighbourhoods G use the full matrix
 * Differences: Julia tasks + Threads.Condition become pthreads + spin/yield flags; every thread draws from its own
 * xoroshiro128+ stream seeded from `seed` (upstream: Rng() with fresh entropy per task).  Like upstream, the result is not
 * reproducible run to run in its draws; the law is that of spdmp with Z = ZigZag(Gamma2, mu).
 * ===================================================================================================== */
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>

typedef struct pctx_s pctx;
typedef struct {
    pctx *P; int ti; pthread_t th; xoro rng;
    heapq Q;
    zzo_event *ev; int64_t nev, cap;
    int64_t res_i; double res_t; int64_t res_acc, res_num;   /* ret[] = (i, t', acc, num), parallel.jl:74 */
    atomic_int wake;                                          /* 0 sleeping, 1 run, 2 done */
    char pad[64];
} pworker;

struct pctx_s {
    int64_t d, K, csz; double Delta, T, t0;
    csc G, G1, tg; const double *h, *mu;
    int64_t *g2ptr, *g2idx;
    double *t, *x, *th, *t_old, *ba, *bb, *c;
    char *inner; int64_t *acc;
    int adapt; double factor;
    atomic_int active; atomic_int failed;
    int64_t err_i; double err_t, err_l, err_lb;
    pworker *W;
};

static inline void p_move(pctx *P, const int64_t *idx, int64_t n, double tp)
{ /* smove_forward!(G, i, t, x, th, t', Z::ZigZag), sfact.jl:6-12 */
    for (int64_t q = 0; q < n; ++q) {
        int64_t k = idx[q] - 1;
        P->x[k] = P->x[k] + P->th[k] * (tp - P->t[k]);
        P->t[k] = tp;
    }
}
static inline void p_ab(pctx *P, int64_t i)
{ /* ab(G1, i, x, th, c, Z::ZigZag), fact_samplers.jl:50-54 with Z.Gamma = the block-diagonal bound matrix */
    P->ba[i - 1] = P->c[i - 1] + (idot(&P->G1, i, P->x) - idot(&P->G1, i, P->mu)) * P->th[i - 1];
    P->bb[i - 1] = P->c[i - 1] / 100 + P->th[i - 1] * idot(&P->G1, i, P->th);
}
static inline void p_resched(pctx *P, xoro *rng, int64_t j)
{ /* Q[q1][q2] = t[j] + poisson_time(b[j], rand(rng)), parallel.jl:50-51,57-58 */
    int64_t q1 = (j - 1) / P->csz, q2 = (j - 1) % P->csz + 1;
    h_set(&P->W[q1].Q, q2, P->t[j - 1] + o_poisson_time(P->ba[j - 1], P->bb[j - 1], xoro_rand(rng)));
}
/* parallel_innermost!, parallel.jl:34-61; returns 1 when the proposal was accepted */
static int p_innermost(pctx *P, xoro *rng, int64_t i, double tp)
{
    const int64_t *nbG = &P->G.rowval[P->G.colptr[i - 1] - 1]; int64_t nG = P->G.colptr[i] - P->G.colptr[i - 1];
    p_move(P, nbG, nG, tp);
    double gi = idot(&P->tg, i, P->x);
    if (P->h) gi = gi - P->h[i - 1];
    double l = zz_pos(gi * P->th[i - 1]);
    double lb = zz_pos(P->ba[i - 1] + P->bb[i - 1] * (P->t[i - 1] - P->t_old[i - 1]));
    if (xoro_rand(rng) * lb < l) {
        if (l >= lb) {
            if (!P->adapt) {
                if (!atomic_exchange(&P->failed, 1)) { P->err_i = i; P->err_t = tp; P->err_l = l; P->err_lb = lb; }
                return 0;
            }
            P->c[i - 1] *= P->factor;
        }
        p_move(P, &P->g2idx[P->g2ptr[i - 1]], P->g2ptr[i] - P->g2ptr[i - 1], tp);
        P->th[i - 1] = -P->th[i - 1];
        for (int64_t p = P->G1.colptr[i - 1]; p < P->G1.colptr[i]; ++p) {
            int64_t j = P->G1.rowval[p - 1];
            p_ab(P, j);
            P->t_old[j - 1] = P->t[j - 1];
            p_resched(P, rng, j);
        }
        P->acc[i - 1] += 1;
        return 1;
    }
    p_ab(P, i);
    P->t_old[i - 1] = P->t[i - 1];
    p_resched(P, rng, i);
    return 0;
}
static void p_push(pworker *w, double t, int64_t i, double x, double th)
{
    if (w->nev == w->cap) { w->cap = w->cap ? 2 * w->cap : 4096; w->ev = (zzo_event *)realloc(w->ev, (size_t)w->cap * sizeof(zzo_event)); }
    zzo_event e = { t, i, x, th };
    w->ev[w->nev++] = e;
}
static inline void p_spin(int *n) { if (++*n < 200) __builtin_ia32_pause(); else { sched_yield(); *n = 0; } }

/* parallel_spdmp_inner!, parallel.jl:63-104 */
static void *p_worker(void *arg)
{
    pworker *w = (pworker *)arg; pctx *P = w->P;
    int64_t acc = 0, num = 0;
    double tnext = P->t0 + P->Delta;
    for (;;) {
        int spin = 0;
        while (atomic_load(&w->wake) == 0) p_spin(&spin);
        if (atomic_load(&w->wake) == 2) return NULL;
        for (;;) {
            num += 1;
            int64_t ii = w->Q.key[1]; double tp = w->Q.val[1];
            int64_t i = (int64_t)w->ti * P->csz + ii;
            if (!P->inner[i - 1] || tp > tnext || atomic_load_explicit(&P->failed, memory_order_relaxed)) {
                tnext = tp + P->Delta;
                w->res_i = i; w->res_t = tp; w->res_acc = acc; w->res_num = num;
                acc = num = 0;
                atomic_store(&w->wake, 0);
                atomic_fetch_sub(&P->active, 1);   /* "last one turns the light off" */
                break;
            }
            if (p_innermost(P, &w->rng, i, tp)) { acc += 1; p_push(w, P->t[i - 1], i, P->x[i - 1], P->th[i - 1]); }
        }
    }
}

static int ev_cmp(const void *a, const void *b)
{
    double ta = ((const zzo_event *)a)->t, tb = ((const zzo_event *)b)->t;
    return (ta > tb) - (ta < tb);
}

zzo_run *zzo_parallel_spdmp(int64_t d,
                            const int64_t *tg_colptr, const int64_t *tg_rowval, const double *tg_nzval, const double *h,
                            const int64_t *bd_colptr, const int64_t *bd_rowval, const double *bd_nzval, const double *mu,
                            double t0, const double *x0, const double *th0, double T, const double *c_in,
                            const uint64_t *seed, int adapt, double factor, int64_t K, double Delta)
{
    zzo_run *r = (zzo_run *)calloc(1, sizeof(zzo_run));
    r->d = d; r->t0 = t0;
    if (K < 1 || d % K != 0) { r->status = ZZO_E_ARG; return r; }   /* Partition(nt, n) = Partition{div(n, nt)}, parallel.jl:27 */
    pctx Ps; pctx *P = &Ps; memset(P, 0, sizeof Ps);
    P->d = d; P->K = K; P->csz = d / K; P->Delta = Delta; P->T = T; P->t0 = t0; P->adapt = adapt; P->factor = factor;
    P->tg.colptr = tg_colptr; P->tg.rowval = tg_rowval; P->tg.nzval = tg_nzval; P->G = P->tg;   /* G = pattern of the target */
    P->G1.colptr = bd_colptr; P->G1.rowval = bd_rowval; P->G1.nzval = bd_nzval;
    P->h = h; P->mu = mu;
    size_t nb = (size_t)d * sizeof(double);
    P->t = (double *)malloc(nb); P->x = (double *)malloc(nb); P->th = (double *)malloc(nb); P->t_old = (double *)malloc(nb);
    P->ba = (double *)malloc(nb); P->bb = (double *)malloc(nb); P->c = (double *)malloc(nb);
    P->inner = (char *)malloc((size_t)d); P->acc = (int64_t *)calloc((size_t)d, sizeof(int64_t));
    r->x0 = (double *)malloc(nb); memcpy(r->x0, x0, nb);
    for (int64_t k = 0; k < d; ++k) { P->t[k] = t0; P->t_old[k] = t0; P->x[k] = x0[k]; P->th[k] = th0[k]; P->c[k] = c_in[k]; }
    /* inner (parallel.jl:116), the subset assertion (:120) and the chunk-locality of G1 / G2 (:124-129) */
    int bad = 0;
    for (int64_t i = 1; i <= d && !bad; ++i) {
        int64_t q = (i - 1) / P->csz; char in = 1;
        for (int64_t p = tg_colptr[i - 1]; p < tg_colptr[i]; ++p) if ((tg_rowval[p - 1] - 1) / P->csz != q) in = 0;
        P->inner[i - 1] = in;
        int64_t pt = tg_colptr[i - 1];
        for (int64_t p = bd_colptr[i - 1]; p < bd_colptr[i]; ++p) {
            if ((bd_rowval[p - 1] - 1) / P->csz != q) bad = 1;                      /* bound crosses chunks */
            while (pt < tg_colptr[i] && tg_rowval[pt - 1] < bd_rowval[p - 1]) ++pt;
            if (pt >= tg_colptr[i] || tg_rowval[pt - 1] != bd_rowval[p - 1]) bad = 2; /* G[i] must contain G1[i] */
        }
    }
    if (bad) { r->status = ZZO_E_GRAPH; goto cleanup0; }
    {   /* G2[i] = setdiff(union(G1[j] for j in G1[i]), G[i]), parallel.jl:122 */
        P->g2ptr = (int64_t *)calloc((size_t)d + 1, sizeof(int64_t));
        int64_t *mark = (int64_t *)calloc((size_t)d + 1, sizeof(int64_t));
        int64_t cap = 16, n = 0;
        P->g2idx = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
        for (int64_t i = 1; i <= d; ++i) {
            P->g2ptr[i - 1] = n;
            for (int64_t p = tg_colptr[i - 1]; p < tg_colptr[i]; ++p) mark[tg_rowval[p - 1]] = -i;
            for (int64_t p = bd_colptr[i - 1]; p < bd_colptr[i]; ++p) {
                int64_t j = bd_rowval[p - 1];
                for (int64_t q = bd_colptr[j - 1]; q < bd_colptr[j]; ++q) {
                    int64_t k = bd_rowval[q - 1];
                    if (mark[k] == -i || mark[k] == i) continue;
                    mark[k] = i;
                    if (n == cap) { cap *= 2; P->g2idx = (int64_t *)realloc(P->g2idx, (size_t)cap * sizeof(int64_t)); }
                    P->g2idx[n++] = k;
                }
            }
        }
        P->g2ptr[d] = n;
        free(mark);
    }
    P->W = (pworker *)calloc((size_t)K, sizeof(pworker));
    xoro seeder = { seed[0], seed[1] };
    for (int64_t i = 1; i <= d; ++i) p_ab(P, i);
    for (int64_t q = 0; q < K; ++q) {
        pworker *w = &P->W[q];
        w->P = P; w->ti = (int)q; w->rng.x = xoro_next(&seeder) | 1ULL; w->rng.y = xoro_next(&seeder);
        w->Q.n = 0; w->Q.lex = 0;
        w->Q.key = (int64_t *)malloc(((size_t)P->csz + 2) * sizeof(int64_t));
        w->Q.val = (double *)malloc(((size_t)P->csz + 2) * sizeof(double));
        w->Q.index = (int64_t *)malloc(((size_t)P->csz + 2) * sizeof(int64_t));
        for (int64_t q2 = 1; q2 <= P->csz; ++q2) {   /* parallel.jl:130-133 (global rand(); no + t0) */
            int64_t i = q * P->csz + q2;
            h_enqueue(&w->Q, q2, o_poisson_time(P->ba[i - 1], P->bb[i - 1], xoro_rand(&seeder)));
        }
        atomic_store(&w->wake, 0);
    }
    {
        xoro orng = { xoro_next(&seeder) | 1ULL, xoro_next(&seeder) };
        double *tpv = (double *)malloc((size_t)K * sizeof(double));
        double *evtime = (double *)calloc((size_t)K, sizeof(double));
        int64_t *perm = (int64_t *)malloc((size_t)K * sizeof(int64_t));
        int64_t *waitfor = (int64_t *)calloc((size_t)K, sizeof(int64_t));
        pworker outer; memset(&outer, 0, sizeof outer);   /* event buffer of the outer task */
        for (int64_t q = 0; q < K; ++q) { tpv[q] = t0; perm[q] = q; }
        double tmin = t0;
        int64_t acc = 0, num = 0;
        atomic_store(&P->active, (int)K);
        const double tl0 = now_s();
        for (int64_t q = 0; q < K; ++q) { atomic_store(&P->W[q].wake, 1); pthread_create(&P->W[q].th, NULL, p_worker, &P->W[q]); }
        /* parallel_spdmp_outer!, parallel.jl:182-253 */
        while (tmin < T) {
            int spin = 0;
            while (atomic_load(&P->active) != 0) p_spin(&spin);
            if (atomic_load(&P->failed)) break;
            for (int64_t q = 0; q < K; ++q)
                if (waitfor[q] == 0) {
                    pworker *w = &P->W[q];
                    for (int64_t e = 0; e < w->nev; ++e) p_push(&outer, w->ev[e].t, w->ev[e].i, w->ev[e].x, w->ev[e].th);
                    w->nev = 0;
                    evtime[q] = w->res_t;
                }
            for (int64_t a = 1; a < K; ++a) {   /* sortperm!(perm, evtime, alg = InsertionSort) */
                int64_t v = perm[a], b2 = a;
                while (b2 > 0 && evtime[perm[b2 - 1]] > evtime[v]) { perm[b2] = perm[b2 - 1]; --b2; }
                perm[b2] = v;
            }
            for (int64_t a = 0; a < K; ++a) {
                int64_t q = perm[a];
                pworker *w = &P->W[q];
                int64_t i = w->res_i; double tq = w->res_t;
                if (waitfor[q] == 0) { num += w->res_num; acc += w->res_acc; tpv[q] = tq; }
                waitfor[q] = 0;
                for (int64_t p = tg_colptr[i - 1]; p < tg_colptr[i]; ++p) {
                    int64_t j = tg_rowval[p - 1];
                    if (j == i) continue;
                    if (tpv[(j - 1) / P->csz] < tq) waitfor[q] = i;
                }
                if (waitfor[q] != 0) continue;
                if (p_innermost(P, &orng, i, tq)) { acc += 1; p_push(&outer, P->t[i - 1], i, P->x[i - 1], P->th[i - 1]); }
            }
            tmin = tpv[0];
            for (int64_t q = 1; q < K; ++q) if (tpv[q] < tmin) tmin = tpv[q];
            if (tmin >= T || atomic_load(&P->failed)) break;
            int nwake = 0;
            for (int64_t q = 0; q < K; ++q) if (waitfor[q] == 0) nwake++;
            atomic_store(&P->active, nwake);
            for (int64_t q = 0; q < K; ++q) if (waitfor[q] == 0) atomic_store(&P->W[q].wake, 1);
        }
        for (int64_t q = 0; q < K; ++q) atomic_store(&P->W[q].wake, 2);
        for (int64_t q = 0; q < K; ++q) pthread_join(P->W[q].th, NULL);
        r->loop_seconds = now_s() - tl0;
        if (atomic_load(&P->failed)) { r->status = ZZO_E_BOUND; r->err_i = P->err_i; r->err_t = P->err_t; r->err_l = P->err_l; r->err_lb = P->err_lb; }
        qsort(outer.ev, (size_t)outer.nev, sizeof(zzo_event), ev_cmp);   /* sort!(Xi.events, by = ev -> ev[1]), parallel.jl:173 */
        r->ev = outer.ev; r->nev = outer.nev; r->cap = outer.cap;
        r->num = num; (void)acc;
        free(tpv); free(evtime); free(perm); free(waitfor);
    }
    for (int64_t q = 0; q < K; ++q) { free(P->W[q].Q.key); free(P->W[q].Q.val); free(P->W[q].Q.index); free(P->W[q].ev); }
    free(P->W); free(P->g2ptr); free(P->g2idx);
cleanup0:
    r->acc = P->acc; r->t = P->t; r->x = P->x; r->th = P->th; r->c = P->c;
    free(P->t_old); free(P->ba); free(P->bb); free(P->inner);
    return r;
}

int zzo_status(const zzo_run *r) { return r->status; }
double zzo_loop_seconds(const zzo_run *r) { return r->loop_seconds; }
void zzo_error_info(const zzo_run *r, int64_t *i, double *t, double *l, double *lb)
{ *i = r->err_i; *t = r->err_t; *l = r->err_l; *lb = r->err_lb; }
int64_t zzo_trace_len(const zzo_run *r) { return r->nev; }
void zzo_trace_copy(const zzo_run *r, zzo_event *dst, int64_t first, int64_t count)
{ memcpy(dst, r->ev + first, (size_t)count * sizeof(zzo_event)); }
void zzo_counts(const zzo_run *r, int64_t *acc, int64_t *num)
{ memcpy(acc, r->acc, (size_t)r->d * sizeof(int64_t)); *num = r->num; }
void zzo_final_state(const zzo_run *r, double *t, double *x, double *th, double *c)
{
    size_t nb = (size_t)r->d * sizeof(double);
    memcpy(t, r->t, nb); memcpy(x, r->x, nb); memcpy(th, r->th, nb); memcpy(c, r->c, nb);
}

/* First moment exactly as Statistics.mean(::Trace), src/trace.jl:182-200 (m1), and the matching
 * exact second moment of the piecewise-linear path (m2; ours -- the reference has no event-based
 * second moment, its tests discretise).  s1/s2 are the unscaled per-coordinate sums the GPU
 * accumulates: s1 = sum (x+xi)(t2-t), s2 = sum (t2-t)(x*x + x*xi + xi*xi). */
void zzo_moments(const zzo_run *r, double *m1, double *m2, double *s1, double *s2)
{
    int64_t d = r->d;
    double *x = (double *)malloc((size_t)d * sizeof(double));
    double *t = (double *)malloc((size_t)d * sizeof(double));
    for (int64_t k = 0; k < d; ++k) { x[k] = r->x0[k]; t[k] = r->t0; m1[k] = 0.0; m2[k] = 0.0; s1[k] = 0.0; s2[k] = 0.0; }
    if (r->nev == 0) { free(x); free(t); return; }
    double T = r->ev[r->nev - 1].t;
    double scale = 1 / (2 * T);
    for (int64_t k = 0; k < r->nev; ++k) {
        double t2 = r->ev[k].t, xi = r->ev[k].x; int64_t i = r->ev[k].i - 1;
        m1[i] += (x[i] + xi) * (t2 - t[i]) * scale;
        s1[i] += (x[i] + xi) * (t2 - t[i]);
        s2[i] += (t2 - t[i]) * (x[i] * x[i] + x[i] * xi + xi * xi);
        t[i] = t2; x[i] = xi;
    }
    for (int64_t k = 0; k < d; ++k) m2[k] = s2[k] / (3 * T);
    free(x); free(t);
}

void zzo_free(zzo_run *r)
{
    if (!r) return;
    free(r->ev); free(r->acc); free(r->t); free(r->x); free(r->th); free(r->c); free(r->x0);
    free(r);
}
