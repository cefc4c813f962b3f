/* zzb200.h -- C-ABI of libzzb200.so: the B200-native drop-in for the factorised local-ZigZag event loop
 * of mschauer/ZigZagBoomerang.jl (reference @ 691afe2).  All entry points are extern "C", take plain
 * pointers and sizes, return an int32 status (0 = ok) and never let an exception cross the boundary.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference):
 *   zzb_problem_create_gaussian   the (grad-phi closure, Z::ZigZag) pair a caller hands to spdmp: the closure
 *                                 `(x,i,G) -> idot(G,i,x)` (src/common.jl:16-24, scripts/gaussianrandomfield.jl:25)
 *                                 becomes the CSC arrays of the target precision (+ optional linear term h);
 *                                 `ZigZag(G, mu)` (src/types.jl:19-27) becomes the CSC arrays of Z.Gamma and Z.mu;
 *                                 also performs the neighbourhood setup of src/sfact.jl:170-178
 *   zzb_problem_create_logistic   the (grad_phi_moving closure, SelfMoving(), A, At, mu, y, ny, k) arguments of the subsampled
 *                                 logistic regression, scripts/logistic.jl:78-107,167 (src/sfact.jl:64-68)
 *   zzb_spdmp_run                 spdmp(grad, t0, x0, th0, T, c, Z, args...; factor, adapt, seed)  src/sfact.jl:162-214
 *                                 and pdmp(...) src/sfact.jl:236 (same event law; the All() graph only changes which
 *                                 coordinates the CPU code moves eagerly)
 *   zzb_run_counts                the returned (acc, num)                                        src/sfact.jl:181-182,211
 *   zzb_run_final_state           the returned (t, x, theta) and the adapted c                   src/sfact.jl:211
 *   zzb_trace_len / _copy         the returned FactTrace's `events` vector                       src/trace.jl:7-13,38
 *   zzb_run_set("max_windows") + zzb_run_execute(T = Inf) + zzb_trace_copy / _clear
 *                                 the pull-style iterator FactSampler / iterate / trace(FS, T)   src/sfactiter.jl:5-79
 *   zzb_run_discretize / zzb_run_grid   collect(discretize(trace, dt)) (src/trace.jl:94-125) produced on the device while the
 *                                 windows are committed: x(t0 + k dt) for every coordinate, without handing the trace back
 *   zzb_trace_moments             Statistics.mean(::Trace) (src/trace.jl:182-200) + matching exact second moment
 *   zzb_sspdmp_run                sspdmp(...) src/ss_fact.jl:159-217 with sspdmp_inner! :78-157, queue_time! :54-66,
 *                                 freezing_time :10-16
 *   zzb_spdmp_boomerang_run       spdmp / pdmp with F::FactBoomerang: flow src/sfact.jl:29-48, rate src/fact_samplers.jl:37-39,
 *                                 constant bound :58-65, velocity refreshment src/sfact.jl:78-114 (types.jl:62-79)
 *   flag ZZB_FLAG_LOCAL_BOUND     the LocalBound methods: ab src/local.jl:2-6, spdmp_inner! :10-78, spdmp/pdmp :95-149,
 *                                 next_time src/not_fact_samplers.jl:43-50
 *   status ZZB_E_BOUND            error("Tuning parameter `c` too small.")                       src/sfact.jl:124
 *
 * Ownership: the caller owns every host array and keeps it alive for the duration of the call; the library owns all
 * device memory behind the opaque handles.  Threading: one host thread at a time per handle; calls block until the
 * work is complete.  Indices are Julia's: 1-based Int64 `colptr` / `rowval`, rows ascending inside a column, so
 * `pointer(G.colptr)`, `pointer(G.rowval)`, `pointer(G.nzval)` can be passed as they are.
 * There is NO CPU fallback: every compute entry point fails with ZZB_E_CUDA when no CUDA driver / device / cubin is found.
 */
#ifndef ZZB200_H
#define ZZB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZZB_OK 0
#define ZZB_E_ARG 1       /* bad argument / handle */
#define ZZB_E_CUDA 2      /* driver, device or cubin problem; see zzb_last_error */
#define ZZB_E_BOUND 3     /* "Tuning parameter `c` too small." (accepted with l >= lb and adapt == 0) */
#define ZZB_E_GRAPH 4     /* malformed sparse matrix */
#define ZZB_E_NOMEM 5
#define ZZB_E_TRACE 6     /* trace buffer too small for a single window */
#define ZZB_E_INTERNAL 9

/* flags of zzb_spdmp_run / zzb_run_create */
#define ZZB_FLAG_NO_TRACE 1u   /* do not record events; counters, final state and moment sums are still produced */
#define ZZB_FLAG_LOCAL_BOUND 2u /* spdmp(..., C::LocalBound, ...) of src/local.jl:95-149: bounds from the target's own first and
                                   second directional derivatives, valid for 2/c/|theta|, then renewed (problem: bnd_* = NULL) */

#define ZZB_FLAG_STICKY 4u      /* sticky ZigZag sspdmp (src/ss_fact.jl): coordinates freeze at 0, thaw after Exp(kappa_i) */
#define ZZB_FLAG_STICKY_REVERSIBLE 16u /* sspdmp(...; reversible = true): a thawing coordinate re-enters with a random sign, ss_fact.jl:111-113 */
#define ZZB_FLAG_STICKY_STRONG_UB 32u  /* sspdmp(...; strong_upperbounds = true): a freeze reschedules nobody, ss_fact.jl:97-107 */
#define ZZB_FLAG_STICKY_ZZ 128u  /* with ZZB_FLAG_STICKY: the dense sticky sampler stickyzz / sspdmp2 (src/stickyzz.jl:176-338) -- the sspdmp loop with
                                   proposal times at rate 0.01 + (a + b t)^+ (queue_time! :144-165, poissontime.jl:93-99) and coordinates that
                                   start at 0 starting frozen (:198-206) */
#define ZZB_FLAG_REFRESH 64u    /* ZigZag with velocity refreshments (Z.lambdaref > 0): src/sfact.jl:78-114,188-190; zzb_run_upload_refresh */
#define ZZB_FLAG_BOOMERANG 8u   /* factorised Boomerang (F::FactBoomerang): rotation around Z.mu, velocity refreshments */

typedef struct zzb_problem_s* zzb_problem_t;
typedef struct zzb_run_s* zzb_run_t;

/* Trace record: the memory layout of Julia's Tuple{Float64,Int64,Float64,Float64} (src/trace.jl:38):
 * (event time, 1-based coordinate, position of that coordinate, its velocity AFTER the flip). */
typedef struct { double t; int64_t i; double x; double theta; } zzb_event;

/* Library / device lifetime.  dev_ids may be NULL (device 0).  cubin_path may be NULL: the kernels are then loaded from
 * $ZZB200_CUBIN or from zzb200_kernels.cubin next to the shared library. */
int32_t zzb_init(int32_t ndev, const int32_t* dev_ids, const char* cubin_path);
int32_t zzb_shutdown(void);
int32_t zzb_last_error(char* buf, int64_t len);
int32_t zzb_device_info(int32_t* sm_count, int64_t* total_mem, char* name, int64_t name_len);
/* CUDA events on the library's launching stream (which = 0 start, 1 stop) for device-side timing of a region. */
int32_t zzb_event_record(int32_t which);
int32_t zzb_event_elapsed_ms(float* ms);

/* Problem = target potential + sampler matrices (two CSC matrices of order d; hvec / bnd_mu may be NULL = zeros).
 * grad phi_i(x) = sum_k tgt[k,i] x_k - hvec[i];  bound uses bnd (Z.Gamma) and bnd_mu (Z.mu), fact_samplers.jl:50-54. */
int32_t zzb_problem_create_gaussian(zzb_problem_t* out, int64_t d,
                                    const int64_t* colptr, const int64_t* rowval, const double* nzval, const double* hvec,
                                    const int64_t* bnd_colptr, const int64_t* bnd_rowval, const double* bnd_nzval,
                                    const double* bnd_mu);
/* Problem with the subsampled logistic-regression target of scripts/logistic.jl (BASELINE config 3): replaces the closure
 * `grad_phi_moving(t, x, theta, i, t', F, A, At, mu, y, ny, k) = gamma0*x[i] - fdot_moving(...)` (scripts/logistic.jl:78-107)
 * that the reference passes to spdmp together with `SelfMoving(), A, At, mu, y, ny, k` (:167; ExtendedForm dispatch,
 * src/sfact.jl:64-68, src/types.jl:103-119).  A is the n x d design matrix and At = SparseMatrixCSC(A') (both Julia-layout
 * CSC, passed as they are), y / ny the successes / failures per design row as Float64, mu_cv the control-variate point
 * (the mode), gamma0 the prior precision, k the number of design rows drawn per evaluation.  bnd_* / bnd_mu are Z.Gamma /
 * Z.mu of the sampler (the script's Zdrop).  The k row indices are drawn from the proposing coordinate's counter stream
 * (the reference uses Julia's global RNG, :83,86).  Run with zzb_spdmp_run / the staged calls; plain ZigZag only. */
int32_t zzb_problem_create_logistic(zzb_problem_t* out, int64_t d, int64_t n,
                                    const int64_t* a_colptr, const int64_t* a_rowval, const double* a_nzval,
                                    const int64_t* at_colptr, const int64_t* at_rowval, const double* at_nzval,
                                    const double* y, const double* ny, const double* mu_cv, double gamma0, int64_t k,
                                    const int64_t* bnd_colptr, const int64_t* bnd_rowval, const double* bnd_nzval,
                                    const double* bnd_mu);
int32_t zzb_problem_free(zzb_problem_t p);

/* One-call form, host buffers in and out: runs the sampler from t0 until the first accepted event at or after T.
 * c is in/out (adapted when adapt != 0).  seed[2] keys the counter-based uniform streams. */
int32_t zzb_spdmp_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                      const uint64_t* seed, int32_t adapt, double factor, uint32_t flags, zzb_run_t* out);

/* Sticky ZigZag, sspdmp(grad, t0, x0, th0, T, c, Z, kappa) of src/ss_fact.jl:159-217 (reversible = false,
 * strong_upperbounds = false, adapt = false).  The trace also holds the freeze events (x = -0*theta, theta = 0) and the
 * thaw events; zzb_run_counts returns the accepted reflections. */
int32_t zzb_sspdmp_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, const double* c,
                       const double* kappa, const uint64_t* seed, uint32_t flags, zzb_run_t* out);

/* Factorised Boomerang, spdmp(grad, t0, x0, th0, T, c, FactBoomerang(Gamma, mu, lambdaref, sigma; rho), ...) (src/sfact.jl:162-214
 * with the FactBoomerang methods; the problem's bnd_* / bnd_mu are Z.Gamma / Z.mu).  sigma[d] scales the refreshed
 * velocities (types.jl:79: diag(Gamma)^-1/2), lambdaref > 0 is the total refreshment rate, rho the autoregression
 * coefficient of the refreshment.  The trace holds reflections and refreshments; zzb_run_counts the reflections.
 * c is in/out like zzb_spdmp_run. */
int32_t zzb_spdmp_boomerang_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                                const double* sigma, double lambdaref, double rho, const uint64_t* seed, int32_t adapt,
                                double factor, uint32_t flags, zzb_run_t* out);

/* sspdmp(...; adapt = true, factor = 1.5) (src/ss_fact.jl:132-136,159): as zzb_sspdmp_run, but an accepted proposal with l > lb
 * multiplies c[i] by factor instead of ending the run with ZZB_E_BOUND; c is in/out (the adapted bounds).  The reference also resets
 * its diagnostic counters (acc, num) at every adaptation (:134); zzb_run_counts returns the totals. */
int32_t zzb_sspdmp_adapt_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                             const double* kappa, const uint64_t* seed, int32_t adapt, double factor, uint32_t flags, zzb_run_t* out);

/* sspdmp3 / sparsestickyzz (src/sparsestickyzz.jl:405-422,192-257): the strong-bound sparse sticky ZigZag of BASELINE config 4.
 * One bound constant c (SparseStickyUpperBounds :127-142, adapt = false), one thaw rate kappa (StickyBarriers), rule 0 = :sticky,
 * 1 = :reversible.  Coordinates with x0 == 0 start frozen (sparsestickystate :10-12); theta0 = velocities of the others; time
 * starts at 0.  Trace records as for zzb_sspdmp_run (theta == 0: the coordinate froze). */
int32_t zzb_sspdmp3_run(zzb_problem_t p, const double* x0, const double* theta0, double T, double c, double kappa, int32_t rule,
                        const uint64_t* seed, uint32_t flags, zzb_run_t* out);

/* sspdmp4 / asynchzz (src/asynchzz.jl:250-265,80-147): the same strong-bound sticky process with a bound constant c[i] and a thaw
 * rate kappa[i] PER coordinate (StrongUpperBounds :2-7,20-28; StickyBarriers :258), start time t0; coordinates with x0 == 0 start
 * frozen and continue, once thawed, with theta0 (rule :sticky of :206-213).  adapt = false only.  The reference's parallel schedule
 * (local minima of a PartialQueue, :150-245) is replaced by the device's own; the run ends after the first event at or after T. */
int32_t zzb_sspdmp4_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, const double* c,
                        const double* kappa, const uint64_t* seed, uint32_t flags, zzb_run_t* out);

/* spdmp / pdmp with Z = ZigZag(Gamma, mu, sigma; lambdaref > 0): velocity refreshments theta_i <- sigma_i * (+-1) at total rate
 * lambdaref (hasrefresh src/fact_samplers.jl:19; refresh branch src/sfact.jl:78-114, clock :188-190).  Refreshments are trace events
 * but not acceptances.  Per-coordinate clocks of rate lambdaref / d (superposition of the reference's single clock). */
int32_t zzb_spdmp_refresh_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                              const double* sigma, double lambdaref, const uint64_t* seed, int32_t adapt, double factor,
                              uint32_t flags, zzb_run_t* out);

/* Staged form (what zzb_spdmp_run is made of); lets a caller keep inputs resident in HBM and time the kernel alone. */
int32_t zzb_run_create(zzb_problem_t p, uint32_t flags, int64_t trace_capacity_events, zzb_run_t* out);
int32_t zzb_run_upload(zzb_run_t r, double t0, const double* x0, const double* theta0, const double* c,
                       const uint64_t* seed, int32_t adapt, double factor);
/* Coordinate sharding over the GPUs of one node (one process per GPU): rank r owns a contiguous block of coordinates;
 * halo records are read, and remote coordinates queued, through CUDA-IPC peer mappings over NVLink.  Order: zzb_run_create,
 * zzb_run_shard, exchange zzb_run_ipc_export blobs (e.g. torch.distributed all_gather), zzb_run_ipc_import for every peer,
 * zzb_run_upload on every rank, a host barrier, zzb_run_execute on every rank.  Results cover the owned range.  Plain ZigZag,
 * LocalBound, sticky (sspdmp / sspdmp2) and Boomerang runs shard; refreshments, strong bounds and the logistic target do not. */
int32_t zzb_run_shard(zzb_run_t r, int32_t rank, int32_t nranks);
int32_t zzb_run_ipc_export(zzb_run_t r, void* buf, int64_t cap, int64_t* len);
int32_t zzb_run_ipc_import(zzb_run_t r, int32_t peer_rank, const void* buf, int64_t len);
int32_t zzb_run_range(zzb_run_t r, int64_t* lo, int64_t* hi);       /* owned coordinates [lo, hi), 0-based */
int32_t zzb_run_upload_kappa(zzb_run_t r, const double* kappa);      /* sticky runs: before zzb_run_upload */
int32_t zzb_run_upload_boomerang(zzb_run_t r, const double* sigma, double lambdaref, double rho);  /* Boomerang runs: before zzb_run_upload */
int32_t zzb_run_upload_refresh(zzb_run_t r, const double* sigma, double lambdaref);   /* ZZB_FLAG_REFRESH runs: before zzb_run_upload */
int32_t zzb_run_reset(zzb_run_t r);                                  /* re-initialise from the inputs resident in HBM */
int32_t zzb_run_execute(zzb_run_t r, double T, float* device_ms);   /* device_ms: CUDA-event time of the kernel(s) */
int32_t zzb_run_set(zzb_run_t r, const char* key, double value);    /* "delta0", "target_frac" / "target_flip_frac" (proposals / accepted flips per window over d), "tag_limit", "max_windows", "grid", "host_sort", "strong_c" / "strong_rule" (sparsestickyzz), "seq_warps" (sequential chains: 1, 2, 4 or 8 warps per chain thinning speculatively; 0 = automatic), "schedule": -1 automatic (default: sequential chains for the logistic target and for chains of at most 64 coordinates or densely coupled ones, else 1), 2 sequential chains -- one warp per connected component runs spdmp_inner! (src/sfact.jl:73-145) as written; plain ZigZag, components of at most 2700 coordinates --, 1 windowed asynchronous relaxation, 0 its pass-synchronous predecessor */
int32_t zzb_run_stats(zzb_run_t r, int64_t* out, int32_t n);        /* windows, retries, passes, node evaluations, rebases,
                                                                       kernel launches, grid size, block size */

/* one-shot read-back straight into caller buffers (any may be NULL; pinned buffers copy at PCIe speed) */
int32_t zzb_run_fetch(zzb_run_t r, double* t, double* x, double* theta, double* c, int64_t* acc, double* s1, double* s2,
                      int64_t* num, int64_t* nacc);
int32_t zzb_run_counts(zzb_run_t r, int64_t* acc, int64_t* num);
int32_t zzb_run_final_state(zzb_run_t r, double* t, double* x, double* theta, double* c);
int32_t zzb_trace_len(zzb_run_t r, int64_t* n);
int32_t zzb_trace_copy(zzb_run_t r, zzb_event* dst, int64_t first, int64_t count);
int32_t zzb_trace_clear(zzb_run_t r);                                 /* streaming: drop the events already copied out */
int32_t zzb_trace_moments(zzb_run_t r, double* m1, double* m2);     /* time averages of x and x^2 over [t0, last event] */
int32_t zzb_trace_sums(zzb_run_t r, double* s1, double* s2);        /* the unscaled device accumulators */
/* subtrace at the source (src/trace.jl:275-290): only the events of the coordinates J (1-based, strictly ascending) are recorded,
 * renumbered to their position in J; nJ = 0 lifts the filter.  Before zzb_run_upload. */
int32_t zzb_run_trace_filter(zzb_run_t r, const int64_t* J, int64_t nJ);
/* cummean(trace) (src/trace.jl:203-225): per coordinate the running time average after each of its events, CSR-shaped -- the values
 * of coordinate k (0-based) are entries offsets[k] .. offsets[k+1]-1 of times[] / values[] (sizes d + 1, zzb_trace_len, zzb_trace_len);
 * the reference's leading pair (t0, x0_k) is left to the caller.  Computed on the device while the trace is still in HBM. */
int32_t zzb_trace_cummean(zzb_run_t r, int64_t* offsets, double* times, double* values);
/* inclusion_prob(trace) (src/trace.jl:161-178) of a sticky run from a device accumulator (no trace needed) */
int32_t zzb_trace_inclusion(zzb_run_t r, double* p);
/* Device-side discretisation.  zzb_run_discretize (before zzb_run_upload) asks for the n_rows grid times t0 + k dt,
 * k = 0 .. n_rows-1; zzb_run_grid (after zzb_run_execute) copies rows [first_row, first_row + n) of the row-major
 * n_rows x d array (row k = x(t0 + k dt), evaluated from the anchor of the segment containing the grid time) and reports in
 * *valid_rows how many leading rows lie at or before the simulated frontier (later rows are NaN). */
int32_t zzb_run_discretize(zzb_run_t r, double dt, int64_t n_rows);
int32_t zzb_run_grid(zzb_run_t r, double* xs, int64_t first_row, int64_t n, int64_t* valid_rows);
int32_t zzb_run_error_info(zzb_run_t r, int64_t* i, double* t, double* l, double* lb);  /* after ZZB_E_BOUND */
/* Test probe: the DEVICE build of the scalar primitives the kernels and the oracle share (csrc/zz_math.h) on host-supplied
 * arguments -- kind 0 log(x), 1 exp(x), 2 sincos(x) -> (o1, o2), 3 poisson_time(a = x, b = y, u = z) (src/poissontime.jl:8-30),
 * 4 the counter-based uniforms u(seed = (x[0], y[0]) bit patterns, coordinate k, counter k ^ 0x5bd1), k = 0 .. n-1. */
int32_t zzb_math_probe(int32_t kind, int64_t n, const double* x, const double* y, const double* z, double* o1, double* o2);
int32_t zzb_run_free(zzb_run_t r);

#ifdef __cplusplus
}
#endif
#endif /* ZZB200_H */
