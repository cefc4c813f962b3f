// zz_dev.h -- launch-parameter and device control-block layouts shared by zz_kernels.cu (device) and
// zzb200.cpp (host, driver API).  Plain PODs; identical layout under nvcc and g++ on x86-64.
#ifndef ZZ_DEV_H
#define ZZ_DEV_H

#include "zz_core.h"
#include "zz_ctl.h"
#include "zz_logit.h"
#include "zz_strong.h"

struct ZzEvent {  // memory layout of Tuple{Float64,Int64,Float64,Float64}, src/trace.jl:38
    double t; long long i; double x; double theta;
};

// Message one GPU leaves in a peer's mailbox at every pass boundary (all-reduce by all-to-all stores over NVLink).
struct ZzMsg {
    unsigned long long sum;      // work-list appends issued / proposals committed (summed over ranks)
    unsigned long long minkey;   // earliest-flip key (min over ranks)
    unsigned int flags;          // ZZ_X_* (or-ed over ranks)
    unsigned int pad;
    unsigned long long epoch;    // written last; the receiver polls it
};
#define ZZ_DBG_REC 96
#define ZZ_X_OVERFLOW 1u
#define ZZ_X_STOP 2u

// Counters and persistent controller state; lives in device memory, zeroed by the host before the first launch.
struct ZzDevCtl {
    unsigned long long bar;          // grid barrier arrival counter (monotone; zeroed by the host before each launch)
    unsigned int wl_cnt[3];          // relaxation work lists, rotated per pass; top bit = "a coordinate overflowed"
    unsigned int viol;               // bound violation seen (adapt == false), sfact.jl:124
    unsigned int touched_cnt[3];     // per window attempt (attempt number mod 3)
    int viol_i;
    unsigned long long smin_key[3];  // order-preserving key of the earliest flip (phase B)
    unsigned long long nprop_win[3]; // proposals | accepted flips << 32 committed by the window of that attempt slot
    unsigned long long num;          // total proposals (sfact.jl:120)
    unsigned long long nacc;         // total accepted flips
    unsigned long long trace_len;    // records appended to the trace buffer (events + window-end markers)
    unsigned long long f0_key;       // min over initial proposal times (init kernel)
    double viol_t, viol_l, viol_lb;
    unsigned int trace_full;         // (defensive) an event did not fit; cannot happen with the drain protocol
    unsigned int need_drain;         // kernel returned because the next window might not fit into the trace buffer
    unsigned int started;            // controller below is valid (resumed launch)
    unsigned int cur;                // last iteration tag used
    unsigned int itg;                // index of the work list consumed last
    unsigned int wattempt;           // window attempts so far
    unsigned int tail_li, tail_cur;  // pass counters published by CTA 0 after a run of block-local tail passes
    // persisted across launches
    ZzCtl ctl;
    // statistics
    unsigned long long windows, retries, iters, node_evals, rebases;
    // wall time (ns, %globaltimer) CTA 0 spent in: 0 scan pass work, 1 relaxation pass work, 2 tail passes, 3 commit work,
    // 4 waiting at grid barriers, 5 phase-B scans, 6 number of grid barriers, 7 number of tail passes
    unsigned long long tprof[8];
    unsigned long long dbg[8];       // development counters (ZZ_PROF_NODE builds)
    // multi-GPU pass-boundary exchange
    unsigned long long issued[3];    // appends issued by this rank into ANY rank's next work list, per list slot
    unsigned long long xrelease;     // bumped by CTA 0 once the exchange of the current boundary is complete
    ZzMsg xres[2];                   // reduced result of the exchange, by boundary parity
    unsigned long long tail_xep;     // sharded tail passes (CTA 0 only): exchange counter and last reduced result, for the other CTAs
    ZzMsg tail_xres;
    ZzMsg mbox[2][ZZ_MAXRANKS];      // incoming messages, by boundary parity and sender
    // asynchronous tile-local relaxation (zz_run_body_async), per window attempt (attempt number mod 3):
    // Quiescence detection with two MONOTONE counters per GPU: created = evaluations queued for tiles of this GPU (by anybody;
    // + one token per CTA until its scan is done), done = evaluations finished (or dropped as duplicates) by this GPU.  The window
    // has converged when the sums over all GPUs agree -- read `done` first, `created` second: created >= done at all times, so
    // equal sums prove that nothing was queued or in flight anywhere at an instant in between.
    unsigned long long created[3], done[3];
    unsigned int abortf[3];          // some evaluation overflowed (flips / pool / items / tags / inbox): retry the window shorter
                                     // (sharded: raised in EVERY rank's copy, read locally)
    unsigned int doneflag[3];        // sharded: CTA 0 of this GPU saw the node-wide sums agree
};

struct ZzParams {
    ZzGraph g;
    ZzView v;
    const int32_t* dptr;     // dependents: who reads coordinate j
    const int32_t* didx;
    ZzSpec* spec;
    double* viol_info;       // [d][3] (t, l, lb) of a recorded violation
    unsigned int* dstamp;    // iteration tag for which the coordinate is (already) queued
    unsigned int* acc;       // accepted flips per coordinate (sfact.jl:122)
    double* s1;              // sum (x_prev + x_new)(t_new - t_prev)               (trace.jl:194, unscaled)
    double* s2;              // sum (t_new - t_prev)(x_prev^2 + x_prev x_new + x_new^2)
    int32_t* wl[3];
    int32_t* touched[1];
    ZzEvent* trace;
    unsigned long long trace_cap;
    ZzDevCtl* ctl;
    double t0, T, delta0, target, target_flips;
    unsigned int tag_limit;
    unsigned int max_windows;   // return to the host after this many committed windows (0 = run to the end)
    int32_t record_trace;
    int32_t pad;
    // device-side discretize (src/trace.jl:94-125 produced on the device): row k of `grid` holds x(t0 + k grid_dt) for every
    // coordinate, written at commit for the grid times inside each committed segment; grid_n == 0: off
    double* grid;
    double grid_dt;
    long long grid_n;
    // peer mappings (nranks > 1): who owns coordinate k also owns its stamp, its work-list slots and its counters
    unsigned int* dstamp_peer[ZZ_MAXRANKS];
    int32_t* wl_peer[3][ZZ_MAXRANKS];
    int32_t* touched_peer[ZZ_MAXRANKS];
    ZzDevCtl* ctl_peer[ZZ_MAXRANKS];
    // asynchronous relaxation: marks that cross a tile boundary are pushed into the owning CTA's inbox.  Entry =
    // (window attempt << 32) | coordinate, so stale entries of earlier attempts read as "not written yet"; the counters rotate
    // with the attempt number mod 3 and are reset by their owner one attempt ahead.
    unsigned long long* inbox;      // [grid][inbox_cap]
    unsigned int* inbox_cnt;        // [3][grid]
    unsigned int inbox_cap;
    unsigned int flag_words;        // 32-bit words per per-tile bit array (dynamic shared memory = 2 arrays)
    unsigned long long* inbox_peer[ZZ_MAXRANKS];   // sharded runs: the inboxes / counters of every rank (own entries = the plain pointers)
    unsigned int* inbox_cnt_peer[ZZ_MAXRANKS];
    int32_t tile_per, tile_pad;     // coordinates per tile (a multiple of 32), identical on every rank
    int32_t setup_lo, setup_hi;     // records written by zz_setup_kernel (sharded lattice: slab + halo columns; otherwise [0, d))
    int32_t init_lo, init_hi;       // coordinates initialised by zz_init_kernel (sharded lattice: the owned slab; otherwise [0, d))
    // development: per-CTA log of one window (records of 4 x u64: kind, count, globaltimer, clock64), ZZB200_DBG_WINDOW
    unsigned long long* dbgbuf;     // [grid][ZZ_DBG_REC][4] or null
    unsigned int dbg_window;
    unsigned int eval_threads;      // development: threads of a CTA that take queue entries (0 = all)
    // subtrace at the source (src/trace.jl:275-290): when set, only events of coordinates with trace_map[j] != 0 are recorded,
    // renumbered to trace_map[j] (the 1-based position of j in the caller's sorted list J)
    const int32_t* trace_map;
    // sticky samplers: time spent away from 0 per coordinate, sum of [x_prev != 0 or x_new != 0] (t_new - t_prev) over its
    // events (inclusion_prob of src/trace.jl:161-178, unscaled); null = not wanted
    double* s3;
    // subsampled logistic target (zz_logit.h; only read by zz_run_kernel_csr_logit)
    ZzLogit lg;
    ZzStrong st;
};

// Sequential chains (zz_seq.cuh): one warp runs the exact event loop of one connected component of the dependency graph.
// Compact 0-based CSC copies of the sampler matrix Z.Gamma (bound and reschedule lists, G1[i] = rows of column i) and -- for a
// Gaussian target that is not the same object -- of the target precision; comp[] holds the ncomp + 1 component boundaries
// (components are contiguous index ranges).  phase: 0 = every item before T; 1 = probe (no writes) for the first accepted flip
// at or after T, its time min-reduced into ctl->smin_key[0]; 2 = every item up to and including that time.
struct __attribute__((aligned(32))) ZzSeqEnt {   // one stored entry of the design matrix A and where its row starts: one 32-byte sector
    int32_t row, q0, len, pad;       // row of A; offset / length of that row in (rcol, rval)
    double val, pad2;
};
struct __attribute__((aligned(32))) ZzSeqRow {   // per design row: y, m - y, sigmoidn(u0), nsigmoid(u0) with u0 = idot(At, row, mu)
    double y, ny, sn0, ns0;          // (scripts/logistic.jl:89-91: the control-variate terms do not depend on the state)
};
struct ZzSeq {
    const int32_t* bcp; const int32_t* brow; const double* bval;
    const int32_t* tcp; const int32_t* trow; const double* tval;
    const int32_t* comp;
    const int32_t* orig;             // chain-order id -> original id: the matrices above and comp[] use chain-order ids (every chain a
                                     // contiguous range); draw streams, trace ids and all per-coordinate arrays keep the original ids
    const ZzSeqEnt* ent;             // logistic target: entries of A, column by column (same order as ZzLogit::arow)
    const ZzSeqRow* rowrec;
    const int32_t* rcoln;            // ZzLogit::rcol in chain-order ids
    int32_t ncomp, phase, ncmax;
    int32_t colmax;                  // scratch entries per column-product array: max(32, longest column), even
};

// order-preserving map double -> uint64 (so atomicMin works for any sign)
ZZ_HD unsigned long long zz_key(double x)
{
    unsigned long long u = zz_d2u(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
ZZ_HD double zz_unkey(unsigned long long k)
{
    return zz_u2d((k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k);
}

#endif  // ZZ_DEV_H
