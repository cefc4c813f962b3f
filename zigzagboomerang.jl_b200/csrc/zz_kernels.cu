// zz_kernels.cu -- sm_100a kernels of the windowed local-ZigZag event loop (compiled to a .cubin and
// loaded by zzb200.cpp through the driver API).
//
//   zz_setup_kernel   pack (x0, theta0, c) into the per-coordinate records
//   zz_init_kernel    initial bounds and first proposal times   (src/sfact.jl:184-187)
//   zz_run_kernel     PERSISTENT COOPERATIVE kernel: the whole event loop (sfact.jl:199-208 with
//                     spdmp_inner! :73-145 inside), one grid barrier per relaxation pass
//   zz_export_kernel  unpack the final (t, x, theta, c) for the host
//
//   zz_run_kernel_csr_logit: the same event loop with the subsampled logistic-regression target of scripts/logistic.jl
//                     (zz_logit.h) instead of a Gaussian one
//
// Schedule: time is cut into windows [F, H).  Pass 1 of a window scans the proposal times,
// compacts the coordinates with tau < H per warp and evaluates their timelines (zz_process_node); every
// coordinate whose list of accepted flips changed marks the coordinates that read it; the following passes
// re-evaluate exactly those, until no list changes.  The fixed point is the sequential event history, so
// the converged window is committed: frontier state, trace records, counters, moment sums.
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include <cooperative_groups/reduce.h>

#include "zz_dev.h"

namespace cg = cooperative_groups;

#ifndef ZZ_BLOCK
#define ZZ_BLOCK 256   // threads per CTA of the small kernels
#endif
// threads per CTA of the persistent event-loop kernels (one CTA per SM; the host reads the two constants below).
// Lattice kernels: 384 threads at 168 registers beat 256 at 219 (measured: -7 % per step); the general-sparse kernels
// keep 256 threads because they need the full register budget.
#ifndef ZZ_RUN_BLOCK_GRID
#define ZZ_RUN_BLOCK_GRID 384
#endif
#ifndef ZZ_RUN_BLOCK_CSR
#define ZZ_RUN_BLOCK_CSR 256
#endif
#define ZZ_RUN_BLOCK_MAX (ZZ_RUN_BLOCK_GRID > ZZ_RUN_BLOCK_CSR ? ZZ_RUN_BLOCK_GRID : ZZ_RUN_BLOCK_CSR)
extern "C" __device__ const int zz_run_block_grid = ZZ_RUN_BLOCK_GRID;
extern "C" __device__ const int zz_run_block_csr = ZZ_RUN_BLOCK_CSR;
#define ZZ_SCAN_U 8
#ifndef ZZ_MINB
#define ZZ_MINB 1
#endif
#define ZZ_OVF_BIT 0x80000000u
#define ZZ_TAIL 256u   // work lists this short are relaxed by one CTA with block-level barriers

__device__ __forceinline__ unsigned long long zz_ld_acq(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long zz_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All CTAs are co-resident (cooperative launch): arrive on a monotone counter, spin until everyone has.
// `prof` (CTA 0, thread 0 only) accumulates the time spent waiting and the number of barriers.
__device__ __forceinline__ void zz_grid_barrier(ZzDevCtl* C, unsigned long long& epoch, unsigned long long* prof)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t0 = prof ? zz_now() : 0ULL;
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(&C->bar, 1ULL);
        while (zz_ld_acq(&C->bar) < epoch) { }
        __threadfence();
        if (prof) { prof[4] += zz_now() - t0; prof[6] += 1; }
    }
    __syncthreads();
}

// The same barrier in two halves: everything this CTA wrote before zz_grid_arrive is visible to every CTA that has returned
// from zz_grid_wait; work that needs nothing from the other CTAs can run in between.
__device__ __forceinline__ void zz_grid_arrive(ZzDevCtl* C, unsigned long long& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(&C->bar, 1ULL);
    }
}
__device__ __forceinline__ void zz_grid_wait(ZzDevCtl* C, unsigned long long epoch, unsigned long long* prof)
{
    if (threadIdx.x == 0) {
        const unsigned long long t0 = prof ? zz_now() : 0ULL;
        while (zz_ld_acq(&C->bar) < epoch) { }
        __threadfence();
        if (prof) { prof[4] += zz_now() - t0; prof[6] += 1; }
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long zz_ld_acq_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void zz_st_rel_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct ZzXres { unsigned long long sum, minkey; unsigned int flags; };

// All-reduce of (sum, min, or) across the GPUs of the node by WARP 0 of CTA 0: lane p writes this rank's contribution into
// peer p's mailbox (plain stores over NVLink, double-buffered by the parity of the exchange counter), one system-scope fence,
// then the epoch word; lane p then waits for peer p's message in the own mailbox and the warp reduces.  All peers are served
// in parallel: the cost is one NVLink round trip, not one per peer.  The result is also left in the control block for the other
// CTAs of this GPU.  Must be called by all 32 lanes of the warp.
__device__ __forceinline__ ZzXres zz_exchange(const ZzParams& P, unsigned long long xep, unsigned long long ms,
                                              unsigned long long mk, unsigned int mf)
{
    ZzDevCtl* C = P.ctl;
    const int par = (int)(xep & 1ULL);
    const int lane = threadIdx.x & 31;
    const bool have = lane < P.v.nranks;
    if (have) {
        ZzMsg* dst = &P.ctl_peer[lane]->mbox[par][P.v.rank];
        dst->sum = ms; dst->minkey = mk; dst->flags = mf;
    }
    __threadfence_system();
    if (have) *(volatile unsigned long long*)&P.ctl_peer[lane]->mbox[par][P.v.rank].epoch = xep;
    unsigned long long s = 0ULL, k = ~0ULL; unsigned int f = 0u;
    if (have) {
        const ZzMsg* src = &C->mbox[par][lane];
        while (zz_ld_acq_sys(&src->epoch) < xep) { }
        s = *(volatile const unsigned long long*)&src->sum;
        k = *(volatile const unsigned long long*)&src->minkey;
        f = *(volatile const unsigned int*)&src->flags;
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
    ZzXres r;
    r.sum = cg::reduce(w, s, cg::plus<unsigned long long>());
    r.minkey = cg::reduce(w, k, cg::less<unsigned long long>());
    r.flags = cg::reduce(w, f, cg::bit_or<unsigned int>());
    if (lane == 0) {
        C->xres[par].sum = r.sum; C->xres[par].minkey = r.minkey; C->xres[par].flags = r.flags;
        __threadfence();
        atomicExch(&C->xrelease, xep);
    }
    return r;
}

// Pass boundary: local grid barrier, then (MULTI) an all-reduce of (sum, min, or) across the GPUs of the node done by
// CTA 0 with plain stores into the peers' mailboxes over NVLink; the other CTAs are released once the result is known.
// The caller passes this rank's contributions as addresses inside the control block; they are read by CTA 0 after
// every local CTA has arrived (so they are final).
template <bool MULTI>
__device__ __forceinline__ ZzXres zz_boundary(const ZzParams& P, unsigned long long& epoch, unsigned long long& xep,
                                              unsigned long long* prof, const unsigned long long* psum,
                                              const unsigned long long* pmin, const unsigned int* pflag_word,
                                              unsigned int flag_mask, unsigned int flag_value, bool arrived = false)
{
    ZzDevCtl* C = P.ctl;
    if (arrived) zz_grid_wait(C, epoch, prof);   // (the caller has called zz_grid_arrive for this boundary already)
    else zz_grid_barrier(C, epoch, prof);
    ZzXres r;
    if (!MULTI) {
        r.sum = psum ? __ldcg(psum) : 0ULL;
        r.minkey = pmin ? __ldcg(pmin) : ~0ULL;
        r.flags = (pflag_word && (__ldcg(pflag_word) & flag_mask)) ? flag_value : 0u;
        return r;
    }
    xep += 1;
    const int par = (int)(xep & 1ULL);
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const unsigned long long t0 = prof ? zz_now() : 0ULL;
        const unsigned long long ms = psum ? __ldcg(psum) : 0ULL;
        const unsigned long long mk = pmin ? __ldcg(pmin) : ~0ULL;
        const unsigned int mf = (pflag_word && (__ldcg(pflag_word) & flag_mask)) ? flag_value : 0u;
        zz_exchange(P, xep, ms, mk, mf);
        if (prof) prof[5] += zz_now() - t0;
    }
    if (threadIdx.x == 0) {
        while (zz_ld_acq(&C->xrelease) < xep) { }
        __threadfence();
    }
    __syncthreads();
    r.sum = __ldcg(&C->xres[par].sum); r.minkey = __ldcg(&C->xres[par].minkey); r.flags = __ldcg(&C->xres[par].flags);
    return r;
}

// Append `val` to a global list; the active lanes of the warp share one atomic.
template <bool MULTI>
__device__ __forceinline__ void zz_append(int32_t* list, unsigned int* cnt, int32_t val)
{
    cg::coalesced_group cgp = cg::coalesced_threads();
    unsigned int base = 0;
    if (cgp.thread_rank() == 0) base = MULTI ? atomicAdd_system(cnt, cgp.size()) : atomicAdd(cnt, cgp.size());
    base = cgp.shfl(base, 0);
    list[(base & ~ZZ_OVF_BIT) + cgp.thread_rank()] = val;
}

// Work list of the single-CTA tail passes, in shared memory (CTA 0 only): appends are shared-memory atomics and the next
// pass reads its entries without a trip to L2.
struct ZzTailList {
    int32_t buf[2][4 * ZZ_TAIL];   // a pass of n <= ZZ_TAIL coordinates marks at most 4 n readers (lattice) / is cut off otherwise
    unsigned int n[2];
    unsigned int ovf;              // a coordinate overflowed, or the list did
};

// Exclusive prefix sum and total of a small per-lane count (< 8) over the currently active lanes: three ballots instead
// of a shuffle scan.
__device__ __forceinline__ unsigned int zz_prefix3(unsigned int mask, unsigned int v, unsigned int& total)
{
    const unsigned int lt = (1u << (threadIdx.x & 31)) - 1u;
    const unsigned int b0 = __ballot_sync(mask, v & 1u), b1 = __ballot_sync(mask, v & 2u), b2 = __ballot_sync(mask, v & 4u);
    total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
    return __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
}

// Up to four candidates per lane: those whose stamp was older than `tagn` go to the next work list, those older
// than `w0` (first touch in this window) also to the touched list.  One atomicAdd per list for the active lanes.
template <bool MULTI>
__device__ __forceinline__ void zz_append4(int32_t* wl, unsigned int* wl_cnt, int32_t* tl, unsigned int* tl_cnt,
                                           const int32_t (&kk)[4], const uint32_t (&old)[4], uint32_t tagn, uint32_t w0)
{
    unsigned int na = 0, nt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { na += (wl && old[q] < tagn) ? 1u : 0u; nt += (tl && old[q] < w0) ? 1u : 0u; }
    const unsigned int mask = __activemask();
    unsigned int ta, tt;
    const unsigned int pa = zz_prefix3(mask, na, ta);
    const unsigned int pt = zz_prefix3(mask, nt, tt);
    const int leadl = __ffs(mask) - 1;
    unsigned int ba = 0, bt = 0;
    if ((threadIdx.x & 31) == leadl) {
        if (ta) ba = MULTI ? atomicAdd_system(wl_cnt, ta) : atomicAdd(wl_cnt, ta);
        if (tt) bt = MULTI ? atomicAdd_system(tl_cnt, tt) : atomicAdd(tl_cnt, tt);
    }
    ba = (__shfl_sync(mask, ba, leadl) & ~ZZ_OVF_BIT) + pa;
    bt = __shfl_sync(mask, bt, leadl) + pt;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (wl && old[q] < tagn) wl[ba++] = kk[q];
        if (tl && old[q] < w0) tl[bt++] = kk[q];
    }
}

__device__ __forceinline__ void zz_store_spec(ZzSpec* p, const ZzNodeOut& o)
{
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(o.a, o.b);
    q[1] = make_double2(o.told, o.tau);
    unsigned long long w = (unsigned long long)o.k | ((unsigned long long)(o.nprop & 0xffffu) << 32) |
                           ((unsigned long long)(o.nflip & 0xffu) << 48) | ((unsigned long long)(o.flags & 0xffu) << 56);
    q[2] = make_double2(o.c, __longlong_as_double((long long)w));
}

struct ZzSpecR { double a, b, told, tau, c; unsigned int k, nprop, nflip, flags; };
__device__ __forceinline__ ZzSpecR zz_load_spec(const ZzSpec* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 u0 = __ldcg(q), u1 = __ldcg(q + 1), u2 = __ldcg(q + 2);
    unsigned long long w = (unsigned long long)__double_as_longlong(u2.y);
    ZzSpecR r;
    r.a = u0.x; r.b = u0.y; r.told = u1.x; r.tau = u1.y; r.c = u2.x;
    r.k = (unsigned int)w; r.nprop = (unsigned int)(w >> 32) & 0xffffu;
    r.nflip = (unsigned int)(w >> 48) & 0xffu; r.flags = (unsigned int)(w >> 56) & 0xffu;
    return r;
}

// Publish the result of one timeline evaluation: if the list of accepted flips differs from the one the
// readers of this pass see, write it into the other slot and queue every coordinate that reads j.
// Stamp up to four readers of a changed coordinate and queue those that were not queued yet.  MULTI: a reader owned
// by another GPU is stamped and queued in its owner's memory with system-scope atomics over NVLink.
template <bool MULTI>
__device__ __forceinline__ void zz_mark4(const ZzParams& P, const int32_t (&kk)[4], uint32_t tagn, uint32_t w0, int nxt, int ws,
                                         ZzTailList* tl, int tslot)
{
    ZzDevCtl* C = P.ctl;
    uint32_t old[4];
    if (!MULTI) {
#pragma unroll
        for (int q = 0; q < 4; ++q) old[q] = (kk[q] >= 0) ? atomicMax(P.dstamp + kk[q], tagn) : 0xffffffffu;
        if (tl) {   // tail pass: the next list lives in shared memory; only a first touch goes to global memory
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (old[q] < tagn) {
                    const unsigned int pos = atomicAdd(&tl->n[tslot], 1u);
                    if (pos < 4u * ZZ_TAIL) tl->buf[tslot][pos] = kk[q];
                    else tl->ovf = 1u;
                }
            }
            bool any = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) any = any || (old[q] < w0);
            if (any) zz_append4<false>(nullptr, nullptr, P.touched[0], &C->touched_cnt[ws], kk, old, tagn, w0);
            return;
        }
        zz_append4<false>(P.wl[nxt], &C->wl_cnt[nxt], P.touched[0], &C->touched_cnt[ws], kk, old, tagn, w0);
        return;
    }
    int32_t kl[4]; uint32_t ol[4];
    unsigned int issued = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        kl[q] = -1; ol[q] = 0xffffffffu;
        if (kk[q] < 0) continue;
        const int o = kk[q] / P.v.shard;
        if (o == P.v.rank) {
            kl[q] = kk[q];
            ol[q] = atomicMax_system(P.dstamp + kk[q], tagn);
            issued += (ol[q] < tagn) ? 1u : 0u;
        } else {
            const uint32_t od = atomicMax_system(P.dstamp_peer[o] + kk[q], tagn);
            if (od < tagn) {
                const unsigned int pos = atomicAdd_system(&P.ctl_peer[o]->wl_cnt[nxt], 1u) & ~ZZ_OVF_BIT;
                P.wl_peer[nxt][o][pos] = kk[q];
                issued += 1u;
                if (od < w0) {
                    const unsigned int pt = atomicAdd_system(&P.ctl_peer[o]->touched_cnt[ws], 1u);
                    P.touched_peer[o][pt] = kk[q];
                }
            }
        }
    }
    zz_append4<true>(P.wl[nxt], &C->wl_cnt[nxt], P.touched[0], &C->touched_cnt[ws], kl, ol, tagn, w0);
    if (issued) atomicAdd(&C->issued[nxt], (unsigned long long)issued);
}

template <int KIND, bool MULTI, int MODE>
__device__ __forceinline__ void zz_publish(const ZzParams& P, int32_t j, const ZzNodeOut& o, uint32_t w0,
                                           uint32_t cur, int nxt, int ws, ZzTailList* tl, int tslot)
{
    ZzDevCtl* C = P.ctl;
    int slot;
    const uint32_t cnt = zz_pick_slot(o.hdr0, o.hdr1, w0, cur, slot);
    bool same = (cnt == o.nflip);
    if (same && cnt) {
        const double* fl = P.v.flips + ((size_t)j * 2 + slot) * ZZ_MAXFLIP;
#pragma unroll
        for (int m = 0; m < ZZ_MAXFLIP; ++m)
            if (m < (int)cnt) same = same && (zz_d2u(__ldcg(fl + m)) == zz_d2u(o.fl[m]));
        if (ZZ_MODE_HAS_VEL(MODE)) {
            const double* ft = P.v.fth + ((size_t)j * 2 + slot) * ZZ_MAXFLIP;
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m)
                if (m < (int)cnt) same = same && (zz_d2u(__ldcg(ft + m)) == zz_d2u(o.fth[m]));
        }
    }
    if (!same) {
        const int wsl = (slot == 0) ? 1 : 0;
        double* fl = P.v.flips + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
        for (int m = 0; m < ZZ_MAXFLIP; ++m)
            if (m < (int)o.nflip) fl[m] = o.fl[m];
        if (ZZ_MODE_HAS_VEL(MODE)) {
            double* ft = P.v.fth + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m)
                if (m < (int)o.nflip) ft[m] = o.fth[m];
        }
        reinterpret_cast<uint32_t*>(P.v.kin + j)[6 + wsl] = (cur << 4) | o.nflip;
        const uint32_t tagn = cur + 1;
        // queue the readers of j: all stamp updates are issued back to back (independent atomics in flight), then
        // the appends of the whole warp share one atomicAdd per list (zz_append4)
        if (KIND == ZZ_KIND_GRID) {
            const int32_t M = P.g.grid_m, N = P.g.grid_n;
            int32_t kk[4];
            if (o.interior) {   // all four readers exist
                kk[0] = j - M; kk[1] = j - 1; kk[2] = j + 1; kk[3] = j + M;
            } else {
                const int32_t col = zz_grid_col(P.g, j), row = j - col * M;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool ok = (q == 0) ? (col > 0) : (q == 1) ? (row > 0) : (q == 2) ? (row < M - 1) : (col < N - 1);
                    kk[q] = ok ? j + ((q == 0) ? -M : (q == 1) ? -1 : (q == 2) ? 1 : M) : -1;
                }
            }
            zz_mark4<MULTI>(P, kk, tagn, w0, nxt, ws, tl, tslot);
        } else {
            const int32_t q1 = P.dptr[j + 1];
            for (int32_t q0 = P.dptr[j]; q0 < q1; q0 += 4) {
                int32_t kk[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) kk[q] = (q0 + q < q1) ? P.didx[q0 + q] : -1;
                zz_mark4<MULTI>(P, kk, tagn, w0, nxt, ws, tl, tslot);
            }
        }
    }
    zz_store_spec(P.spec + j, o);
    if (o.flags & ZZ_F_VIOL) {
        double* vi = P.viol_info + (size_t)j * 3;
        vi[0] = o.viol_t; vi[1] = o.viol_l; vi[2] = o.viol_lb;
    }
    if (o.flags & ZZ_F_OVERFLOW) { atomicOr(&C->wl_cnt[nxt], ZZ_OVF_BIT); if (tl) tl->ovf = 1u; }
}

// One timeline evaluation + publication; kept out of line so the three call sites (scan pass, relaxation pass,
// tail pass) share one copy of the code and its register allocation.
template <int KIND, bool MULTI, int MODE>
__device__ __noinline__ void zz_eval_publish(const ZzParams& P, int32_t j, double H, int incl, uint32_t w0,
                                             uint32_t cur, bool first, int nxt, int ws, ZzTailList* tl = nullptr, int tslot = 0)
{
    ZzNodeOut o;
#ifdef ZZ_PROF_NODE
    const long long c0 = clock64();
    zz_process_node_k<KIND, MODE, MULTI>(P.g, P.v, j, H, incl, w0, cur, first, o);
    const long long c1 = clock64();
    zz_publish<KIND, MULTI, MODE>(P, j, o, w0, cur, nxt, ws, tl, tslot);
    const long long c2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0 && zz_dbg_on) {
        P.ctl->dbg[0] += (unsigned long long)(c2 - c1);   // cycles in publication
        P.ctl->dbg[7] += 1ULL;
        P.ctl->dbg[3] += (unsigned long long)(c1 - c0);   // cycles in the evaluation
    }
#else
    if constexpr (MODE == ZZ_MODE_LOGIT) zz_process_node_logit(P.g, P.v, P.lg, j, H, incl, w0, cur, first, o);
    else if constexpr (MODE == ZZ_MODE_STRONG) zz_process_node_strong(P.g, P.v, P.st, j, H, incl, w0, cur, first, o);
    else zz_process_node_k<KIND, MODE, MULTI>(P.g, P.v, j, H, incl, w0, cur, first, o);
    zz_publish<KIND, MULTI, MODE>(P, j, o, w0, cur, nxt, ws, tl, tslot);
#endif
}

// Device-side discretize: write x_j(t_k) for the grid times t_k = t0 + k dt in (tf, fs] of the segment that starts at the
// anchor (tf, xf, th) and ends at fs (the position is evaluated from the anchor, x = xf + th (t_k - tf), resp. the rotation
// of the Boomerang flow).  Row 0 (t_k = t0) is written by zz_setup_kernel.
template <bool BOOM>
__device__ __forceinline__ void zz_grid_fill(const ZzParams& P, int32_t j, double tf, double xf, double th, double fs, double muj)
{
    const double dt = P.grid_dt;
    long long k = (long long)floor((tf - P.t0) / dt);
    if (k < 0) k = 0;
    while (P.t0 + (double)k * dt <= tf) ++k;
    while (k > 0 && P.t0 + (double)(k - 1) * dt > tf) --k;   // first k with t_k > tf
    for (; k < P.grid_n; ++k) {
        const double tk = P.t0 + (double)k * dt;
        if (tk > fs) break;
        double x;
        if (BOOM) { double tho; zz_boom_at(tf, xf, th, muj, tk, &x, &tho); }
        else x = xf + th * (tk - tf);
        P.grid[(size_t)k * (size_t)P.v.d + (size_t)j] = x;
    }
}

template <int KIND> __device__ __forceinline__ unsigned int zz_reader_ranks(const ZzParams& P, int32_t j);

// Fold the converged end-of-window state of coordinate j into the frontier.  Two halves: the private record (every coordinate
// with a proposal inside the window) and the kinematic record / moment sums / trace (only coordinates with events).
template <int MODE>
__device__ __forceinline__ void zz_commit_light(const ZzParams& P, int32_t j, const ZzSpecR& s, unsigned int& nprop_acc)
{
    ZzDevCtl* C = P.ctl;
    if (s.flags & ZZ_F_VIOL) {
        if (atomicExch(&C->viol, 1u) == 0u) {
            const double* vi = P.viol_info + (size_t)j * 3;
            C->viol_i = j + 1; C->viol_t = __ldcg(vi); C->viol_l = __ldcg(vi + 1); C->viol_lb = __ldcg(vi + 2);
        }
    }
    double2* pq = reinterpret_cast<double2*>(P.v.priv + j);
    pq[0] = make_double2(s.a, s.b);
    pq[1] = make_double2(s.told, s.c);
    P.v.tau[j] = s.tau;
    P.v.kctr[j] = s.k;
    if (MODE == ZZ_MODE_REFRESH) {   // next proposal / refreshment time (tau is the earlier of the two)
        const double2 rs = __ldcg(reinterpret_cast<const double2*>(P.v.rspec) + j);
        reinterpret_cast<double2*>(P.v.rst)[j] = rs;
    }
    nprop_acc += s.nprop;
    if (s.flags & ZZ_F_STICKY_ERR) atomicExch(&C->viol, 2u);      // error("x[i] !~ 0"), ss_fact.jl:89-91
}
template <int MODE, int KIND = ZZ_KIND_CSR, bool MULTI = false>
__device__ __forceinline__ void zz_commit_heavy(const ZzParams& P, int32_t j, const ZzSpecR& s, uint32_t w0, uint32_t cur,
                                                unsigned int& nflip_acc)
{
    ZzDevCtl* C = P.ctl;
    if (s.nflip) {
        nflip_acc += s.nflip;
        double th, tf, xf; uint32_t h0, h1;
        zz_ld_kin(P.v.kin + j, th, tf, xf, h0, h1);
        int slot;
        zz_pick_slot(h0, h1, w0, cur, slot);
        const double* fl = P.v.flips + ((size_t)j * 2 + slot) * ZZ_MAXFLIP;
        unsigned long long pos = 0;
        const int32_t tid = P.trace_map ? __ldg(P.trace_map + j) : j + 1;   // id in the trace (0: filtered out, src/trace.jl:275-290)
        const bool rec = P.record_trace && tid != 0;
        if (P.record_trace) {
            const unsigned int mask = __activemask();
            unsigned int tot;
            const unsigned int pre = zz_prefix3(mask, rec ? s.nflip : 0u, tot);   // nflip <= ZZ_MAXFLIP = 6
            const int leadl = __ffs(mask) - 1;
            unsigned long long base = 0;
            if ((threadIdx.x & 31) == leadl) base = atomicAdd(&C->trace_len, (unsigned long long)tot);
            base = __shfl_sync(mask, base, leadl);
            pos = base + pre;
        }
        double a1 = __ldcg(P.s1 + j), a2 = __ldcg(P.s2 + j);
        const double* ft = ZZ_MODE_HAS_VEL(MODE) ? P.v.fth + ((size_t)j * 2 + slot) * ZZ_MAXFLIP : nullptr;
        double a3 = (ft && P.s3) ? __ldcg(P.s3 + j) : 0.0;
        const bool boom = (MODE == ZZ_MODE_BOOM);
        const double muj = boom ? P.v.bmu[j] : 0.0;
        const bool counted = boom || MODE == ZZ_MODE_REFRESH;     // reflections counted by the timeline (refreshments are events too)
        unsigned int nrefl = counted ? ((s.flags >> 3) & 7u) : 0u;
        for (unsigned int m = 0; m < s.nflip; ++m) {
            const double fs = __ldcg(fl + m);
            double xs, thn;
            if (boom) {   // reflection or refreshment: rotate to the event, then the recorded velocity (sfact.jl:29-38,100-102,130)
                double tho;
                zz_boom_at(tf, xf, th, muj, fs, &xs, &tho);
                thn = __ldcg(ft + m);
            } else if (ft) {   // sticky: flip / freeze (velocity after = 0, x = -0*theta) / thaw (x stays 0), ss_fact.jl:87-123
                thn = __ldcg(ft + m);
                if (thn == 0.0) xs = -0.0 * th;
                else if (th == 0.0) xs = xf;
                else { xs = xf + th * (fs - tf); if (!counted) nrefl++; }
            } else {
                xs = xf + th * (fs - tf); thn = -th; nrefl++;
            }
            if (P.grid_n) zz_grid_fill<MODE == ZZ_MODE_BOOM>(P, j, tf, xf, th, fs, muj);
            if (!boom) {   // moment sums of the piecewise LINEAR path only
                a1 += (xf + xs) * (fs - tf);                      // trace.jl:194 (scaled by 1/(2T) on the host)
                a2 += (fs - tf) * (xf * xf + xf * xs + xs * xs);
            }
            if (ft && !boom && P.s3) a3 += ((xf != 0.0 || xs != 0.0) ? 1.0 : 0.0) * (fs - tf);   // trace.jl:170-172
            th = thn; tf = fs; xf = xs;
            if (rec) {
                if (pos + m < P.trace_cap) {
                    double2* e = reinterpret_cast<double2*>(P.trace + pos + m);   // sfact.jl:50-52
                    e[0] = make_double2(fs, __longlong_as_double((long long)tid));
                    e[1] = make_double2(xs, th);
                } else {
                    C->trace_full = 1u;
                }
            }
        }
        P.s1[j] = a1; P.s2[j] = a2;
        if (ft && P.s3) P.s3[j] = a3;
        P.acc[j] = __ldcg(P.acc + j) + nrefl;            // accepted reflections (freezes and thaws are events, not acceptances)
        double2* kq = reinterpret_cast<double2*>(P.v.kin + j);
        kq[0] = make_double2(th, tf);
        reinterpret_cast<double*>(P.v.kin + j)[2] = xf;
        if (MULTI) {   // sharded: the new anchor also goes into the replicas of the ranks that read j
            for (unsigned int rm = (KIND == ZZ_KIND_GRID) ? zz_reader_ranks<ZZ_KIND_GRID>(P, j) : zz_reader_ranks<ZZ_KIND_CSR>(P, j); rm; rm &= rm - 1u) {
                const int rr = __ffs((int)rm) - 1;
                double2* kr = reinterpret_cast<double2*>(P.v.kin_peer[rr] + j);
                kr[0] = make_double2(th, tf);
                reinterpret_cast<double*>(P.v.kin_peer[rr] + j)[2] = xf;
                __threadfence_system();   // (ordered before the node-wide boundary that ends the commit)
            }
        }
    }
}
template <int MODE, int KIND = ZZ_KIND_CSR, bool MULTI = false>
__device__ __forceinline__ void zz_commit_node(const ZzParams& P, int32_t j, const ZzSpecR& s, uint32_t w0, uint32_t cur,
                                               unsigned int& nprop_acc, unsigned int& nflip_acc)
{
    zz_commit_light<MODE>(P, j, s, nprop_acc);
    zz_commit_heavy<MODE, KIND, MULTI>(P, j, s, w0, cur, nflip_acc);
}

extern "C" __global__ void __launch_bounds__(ZZ_BLOCK)
zz_setup_kernel(const ZzParams P, const double* __restrict__ x0, const double* __restrict__ th0,
                const double* __restrict__ c0)
{
    for (int32_t j = P.setup_lo + blockIdx.x * blockDim.x + threadIdx.x; j < P.setup_hi; j += gridDim.x * blockDim.x) {
        ZzKin k; k.theta = th0[j]; k.tf = P.t0; k.xf = x0[j]; k.hdr[0] = 0; k.hdr[1] = 0;
        ZzPriv p; p.a = 0.0; p.b = 0.0; p.told = P.t0; p.c = c0[j];
        // strong-bound sampler, rule 2 (asynchzz, src/asynchzz.jl:112-116): a coordinate that starts at 0 starts frozen and
        // remembers the velocity it will continue with
        if (P.st.c > 0.0 && P.st.rule == 2 && k.xf == 0.0) { p.told = k.theta; k.theta = 0.0; }
        // stickyzz / sspdmp2 (src/stickyzz.jl:198-206): the same, the velocity travels in the `a` slot like that of every frozen coordinate
        if (P.v.sticky && (P.v.sticky & ZZ_STICKY_ZZ) && k.xf == 0.0) { p.a = k.theta; k.theta = 0.0; }
        P.v.kin[j] = k;
        P.v.priv[j] = p;
        P.dstamp[j] = 0; P.acc[j] = 0; P.s1[j] = 0.0; P.s2[j] = 0.0;
        if (P.s3) P.s3[j] = 0.0;
        if (P.grid_n) P.grid[j] = x0[j];   // row 0: x(t0)
    }
}

// Rows of the discretisation grid between every coordinate's last committed event and the frontier `tend` (all events
// before the frontier are final, so the current segment is valid up to it).  Idempotent; run before the grid is read.
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_grid_tail_kernel(const ZzParams P, double tend)
{
    for (int32_t j = P.v.lo + blockIdx.x * blockDim.x + threadIdx.x; j < P.v.hi; j += gridDim.x * blockDim.x) {
        const ZzKin k = P.v.kin[j];
        if (P.v.boom) zz_grid_fill<true>(P, j, k.tf, k.xf, k.theta, tend, P.v.bmu[j]);
        else zz_grid_fill<false>(P, j, k.tf, k.xf, k.theta, tend, 0.0);
    }
}

template <bool BOOM>
__device__ __forceinline__ void zz_init_body(const ZzParams& P)
{
    ZzView vloc = P.v; vloc.nranks = 1;   // from this rank's own copies (sharded lattice: the owned slab, whose halo is set up too)
    unsigned long long kmin = ~0ULL;
    for (int32_t j = P.init_lo + blockIdx.x * blockDim.x + threadIdx.x; j < P.init_hi; j += gridDim.x * blockDim.x) {
        if (BOOM) zz_init_node_boom(P.g, vloc, j, P.t0);
        else zz_init_node(P.g, vloc, j, P.t0);
        const unsigned long long k = zz_key(vloc.tau[j]);
        kmin = k < kmin ? k : kmin;
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
    kmin = cg::reduce(w, kmin, cg::less<unsigned long long>());
    if (w.thread_rank() == 0 && kmin != ~0ULL) atomicMin(&P.ctl->f0_key, kmin);
}
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_init_kernel(const ZzParams P) { zz_init_body<false>(P); }
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_init_kernel_boom(const ZzParams P) { zz_init_body<true>(P); }
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_init_kernel_strong(const ZzParams P)
{
    unsigned long long kmin = ~0ULL;
    for (int32_t j = P.init_lo + blockIdx.x * blockDim.x + threadIdx.x; j < P.init_hi; j += gridDim.x * blockDim.x) {
        zz_init_node_strong(P.g, P.v, P.st, j, P.t0);
        const unsigned long long k = zz_key(P.v.tau[j]);
        kmin = k < kmin ? k : kmin;
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
    kmin = cg::reduce(w, kmin, cg::less<unsigned long long>());
    if (w.thread_rank() == 0 && kmin != ~0ULL) atomicMin(&P.ctl->f0_key, kmin);
}

extern "C" __global__ void __launch_bounds__(ZZ_BLOCK)
zz_export_kernel(const ZzParams P, double* __restrict__ t, double* __restrict__ x, double* __restrict__ th,
                 double* __restrict__ c, long long* __restrict__ acc)
{
    for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < P.v.d; j += gridDim.x * blockDim.x) {
        const ZzKin k = P.v.kin[j];
        t[j] = k.tf; x[j] = k.xf; th[j] = k.theta; c[j] = P.v.priv[j].c; acc[j] = (long long)P.acc[j];
    }
}

// Test probe: the scalar primitives of zz_math.h evaluated ON THE DEVICE for host-supplied arguments, so that the tests can pin
// them against libm and against the host build of the same header (oracle and kernels share these routines: an error in one of
// them would cancel in every device == oracle comparison).  kind 0: log(x); 1: exp(x); 2: sincos(x) -> (o1, o2);
// 3: poisson_time(a = x, b = y, u = z); 4: u01(seed = (x, y) bit patterns, coordinate = k, counter = k ^ 0x5bd1)
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK)
zz_math_probe_kernel(int kind, long long n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                     double* __restrict__ o1, double* __restrict__ o2)
{
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        if (kind == 0) o1[k] = zz_log(x[k]);
        else if (kind == 1) o1[k] = zz_exp(x[k]);
        else if (kind == 2) { double sn, cs; zz_sincos(x[k], &sn, &cs); o1[k] = sn; o2[k] = cs; }
        else if (kind == 3) o1[k] = zz_poisson_time(x[k], y[k], z[k]);
        else o1[k] = zz_u01(zz_d2u(x[0]), zz_d2u(y[0]), (uint64_t)k, (uint64_t)(k ^ 0x5bd1));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Device-side ordering of the trace (src/trace.jl:38 is a time-ordered vector; upstream's own parallel samplers sort too,
// src/asynchzz.jl:131, src/parallel.jl:168).  The commit appends the events of a window in arbitrary order; before the host
// reads them they are ordered by (time, coordinate) here: a bucket sort on the time axis -- the bucket of an event is a
// monotone function of its time, so buckets are ordered among themselves -- followed by a bitonic sort of every bucket in
// shared memory.  Window-end markers (i == 0) are dropped.  A bucket that does not fit raises `ovf` and the host falls back to
// its own sort.
#define ZZ_TSORT_CAP 1024u     // events per bucket the in-shared-memory sort handles
struct ZzTsort {
    const ZzEvent* in; ZzEvent* out;
    unsigned long long n;       // records in `in` (events + markers)
    double tmin, scale;         // bucket = min(nb - 1, (t - tmin) * scale)
    unsigned int nb;
    unsigned int mode;          // 0: buckets on the time axis; 1: bucket = coordinate (i - 1), nb = d (grouping for cummean)
    unsigned int* cnt;          // [nb]     events per bucket
    unsigned int* base;         // [nb + 1] exclusive prefix sum
    unsigned int* fill;         // [nb]     scatter cursors
    unsigned int* ovf;          // [1]
};
__device__ __forceinline__ unsigned int zz_tsort_bucket(const ZzTsort& Q, double t, long long i)
{
    if (Q.mode == 1u) return (unsigned int)(i - 1);
    const double u = (t - Q.tmin) * Q.scale;
    if (!(u > 0.0)) return 0u;
    return u >= (double)(Q.nb - 1u) ? Q.nb - 1u : (unsigned int)u;
}
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_tsort_hist_kernel(const ZzTsort Q)
{
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < Q.n; e += (unsigned long long)gridDim.x * blockDim.x)
        if (Q.in[e].i != 0) atomicAdd(Q.cnt + zz_tsort_bucket(Q, Q.in[e].t, Q.in[e].i), 1u);
}
extern "C" __global__ void __launch_bounds__(1024) zz_tsort_scan_kernel(const ZzTsort Q)
{   // one CTA: exclusive prefix sum of the bucket counts (chunks of 1024 with a running carry)
    __shared__ unsigned int wsum[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (unsigned int b0 = 0; b0 < Q.nb; b0 += blockDim.x) {
        const unsigned int b = b0 + threadIdx.x;
        const unsigned int v = b < Q.nb ? Q.cnt[b] : 0u;
        if (v > ZZ_TSORT_CAP) *Q.ovf = 1u;
        unsigned int inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const unsigned int up = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += up; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        unsigned int pre = 0, tot = 0;
        for (unsigned int q = 0; q < (blockDim.x >> 5); ++q) { const unsigned int x = wsum[q]; if (q < warp) pre += x; tot += x; }
        if (b < Q.nb) Q.base[b] = carry + pre + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) Q.base[Q.nb] = carry;
}
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_tsort_scatter_kernel(const ZzTsort Q)
{
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < Q.n; e += (unsigned long long)gridDim.x * blockDim.x) {
        const double2* src = reinterpret_cast<const double2*>(Q.in + e);
        const double2 a = src[0], b = src[1];
        if (__double_as_longlong(a.y) == 0LL) continue;   // window-end marker
        const unsigned int bk = zz_tsort_bucket(Q, a.x, __double_as_longlong(a.y));
        const unsigned int pos = Q.base[bk] + atomicAdd(Q.fill + bk, 1u);
        double2* dst = reinterpret_cast<double2*>(Q.out + pos);
        dst[0] = a; dst[1] = b;
    }
}
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_tsort_sort_kernel(const ZzTsort Q)
{
    __shared__ unsigned long long key[ZZ_TSORT_CAP];   // order-preserving image of the time
    __shared__ unsigned int sub[ZZ_TSORT_CAP];          // coordinate (ties in time are broken by it)
    __shared__ unsigned short idx[ZZ_TSORT_CAP];        // position inside the bucket before the sort
    for (unsigned int bk = blockIdx.x; bk < Q.nb; bk += gridDim.x) {
        const unsigned int n = Q.cnt[bk];
        if (n < 2u || n > ZZ_TSORT_CAP) continue;
        ZzEvent* ev = Q.out + Q.base[bk];
        unsigned int m = 2; while (m < n) m <<= 1;
        __syncthreads();
        for (unsigned int e = threadIdx.x; e < m; e += blockDim.x) {
            if (e < n) { key[e] = zz_key(ev[e].t); sub[e] = (unsigned int)ev[e].i; idx[e] = (unsigned short)e; }
            else { key[e] = ~0ULL; sub[e] = 0xffffffffu; idx[e] = 0xffffu; }
        }
        __syncthreads();
        for (unsigned int k = 2; k <= m; k <<= 1) {
            for (unsigned int j = k >> 1; j > 0; j >>= 1) {
                for (unsigned int e = threadIdx.x; e < m; e += blockDim.x) {
                    const unsigned int p = e ^ j;
                    if (p > e) {
                        const bool up = ((e & k) == 0u);
                        const bool gt = key[e] > key[p] || (key[e] == key[p] && sub[e] > sub[p]);
                        if (gt == up) {
                            const unsigned long long tk = key[e]; key[e] = key[p]; key[p] = tk;
                            const unsigned int ts = sub[e]; sub[e] = sub[p]; sub[p] = ts;
                            const unsigned short ti = idx[e]; idx[e] = idx[p]; idx[p] = ti;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // permute the records in place: every thread first reads the (at most four) records that belong at its positions
        double2 ra[4], rb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned int e = threadIdx.x + (unsigned int)q * blockDim.x;
            if (e < n) { const double2* src = reinterpret_cast<const double2*>(ev + idx[e]); ra[q] = src[0]; rb[q] = src[1]; }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned int e = threadIdx.x + (unsigned int)q * blockDim.x;
            if (e < n) { double2* dst = reinterpret_cast<double2*>(ev + e); dst[0] = ra[q]; dst[1] = rb[q]; }
        }
    }
}

// cummean(trace) (src/trace.jl:203-225) on the device.  Input: the records grouped by coordinate (zz_tsort_hist / scan / scatter
// with mode 1; inside a group in arbitrary order).  One thread per coordinate puts its (few) records into time order and walks
// them: y += (x_prev + x)(t - t_prev), value = y / (2 t), exactly the reference's running sum.  A coordinate with more than
// ZZ_CM_MAX events raises `ovf` (the host then does the whole thing itself).
#define ZZ_CM_MAX 64u
struct ZzCummean {
    ZzEvent* ev;                // records grouped by coordinate (sorted in place)
    const unsigned int* base;   // [d + 1]
    const double* x0;
    double t0;
    int32_t d, pad;
    double* times;              // [n] out, grouped like ev
    double* values;             // [n] out
    unsigned int* ovf;
};
extern "C" __global__ void __launch_bounds__(ZZ_BLOCK) zz_cummean_kernel(const ZzCummean Q)
{
    for (int32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < Q.d; k += gridDim.x * blockDim.x) {
        const unsigned int b0 = Q.base[k], n = Q.base[k + 1] - b0;
        if (n == 0u) continue;
        if (n > ZZ_CM_MAX) { *Q.ovf = 1u; continue; }
        ZzEvent* ev = Q.ev + b0;
        for (unsigned int a = 1; a < n; ++a) {   // insertion sort by time (a coordinate has one event at a time)
            const ZzEvent x = ev[a];
            unsigned int p = a;
            while (p > 0u && ev[p - 1].t > x.t) { ev[p] = ev[p - 1]; --p; }
            ev[p] = x;
        }
        double y = 0.0, tp = Q.t0, xp = Q.x0[k];
        for (unsigned int a = 0; a < n; ++a) {
            const double t = ev[a].t, x = ev[a].x;
            y += (xp + x) * (t - tp);            // trace.jl:217
            tp = t; xp = x;
            Q.times[b0 + a] = t;
            Q.values[b0 + a] = y / (2 * t);      // trace.jl:222
        }
    }
}

// Pass 1 finds its coordinates by scanning: every CTA scans its contiguous share of the owned proposal times and compacts
// the coordinates with a proposal inside the window into a CTA-wide queue (the work list consumed last is free during this
// pass; the CTA uses the slice that mirrors its share); the CTA then works through the queue with all its threads:
// ceil(active / threads) evaluation rounds.  TOUCH (single GPU): every coordinate of the queue is new to the window, so the
// whole queue is appended to the touched list with ONE atomic per CTA and no evaluation waits for an atomic.
template <bool TOUCH>
__device__ __forceinline__ unsigned int zz_build_queue(const ZzParams& P, unsigned int* sq_cnt, int32_t lo, int32_t hi, uint32_t li,
                                                       double H, int incl, int ws, const int32_t*& qout)
{
    ZzDevCtl* C = P.ctl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t nwc = (int32_t)(blockDim.x >> 5);
    const int32_t per = (hi - lo + (int32_t)gridDim.x - 1) / (int32_t)gridDim.x;
    const int32_t c_lo = lo + (int32_t)blockIdx.x * per;
    const int32_t c_hi = (c_lo + per < hi) ? c_lo + per : hi;
    int32_t* q = P.wl[li] + c_lo;
    __syncthreads();   // the previous user of the queue counter is done
    if (threadIdx.x == 0) { sq_cnt[0] = 0u; sq_cnt[1] = 0u; }
    __syncthreads();
    for (int32_t base = c_lo + warp * (32 * ZZ_SCAN_U); base < c_hi; base += nwc * (32 * ZZ_SCAN_U)) {
        bool act[ZZ_SCAN_U];
#pragma unroll
        for (int u = 0; u < ZZ_SCAN_U; ++u) {
            const int32_t j = base + u * 32 + lane;
            act[u] = false;
            if (j < c_hi) { const double tj = __ldcg(P.v.tau + j); act[u] = (tj < H) || (incl && tj == H); }
        }
#pragma unroll
        for (int u = 0; u < ZZ_SCAN_U; ++u) {
            const unsigned int m = __ballot_sync(0xffffffffu, act[u]);
            if (m) {
                unsigned int wb = 0;
                if (lane == 0) wb = atomicAdd(sq_cnt, (unsigned int)__popc(m));
                wb = __shfl_sync(0xffffffffu, wb, 0);
                if (act[u]) q[wb + __popc(m & ((1u << lane) - 1u))] = base + u * 32 + lane;
            }
        }
    }
    __syncthreads();
    const unsigned int qn = sq_cnt[0];
    if (TOUCH && qn) {
        if (threadIdx.x == 0) sq_cnt[1] = atomicAdd(&C->touched_cnt[ws], qn);
        __syncthreads();
        const unsigned int tb = sq_cnt[1];
        for (unsigned int e = threadIdx.x; e < qn; e += blockDim.x) P.touched[0][tb + e] = __ldcg(q + e);
    }
    qout = q;
    return qn;
}

template <int KIND, bool MULTI, int MODE>
__device__ __forceinline__ void zz_queue_pass(const ZzParams& P, unsigned int* sq_cnt, int32_t lo, int32_t hi, uint32_t li,
                                              double H, int incl, uint32_t w0, uint32_t cur, int nxt, int ws,
                                              unsigned long long& st_evals)
{
    ZzDevCtl* C = P.ctl;
    const int32_t* q;
    const unsigned int qn = zz_build_queue<!MULTI>(P, sq_cnt, lo, hi, li, H, incl, ws, q);
    for (unsigned int e = threadIdx.x; e < qn; e += blockDim.x) {
        const int32_t j = __ldcg(q + e);
        if (MULTI) {
            const uint32_t old = atomicMax_system(P.dstamp + j, cur);
            if (old < w0) zz_append<MULTI>(P.touched[0], &C->touched_cnt[ws], j);
        } else {
            atomicMax(P.dstamp + j, cur);   // result unused (RED); the queue went to the touched list in bulk
        }
        zz_eval_publish<KIND, MULTI, MODE>(P, j, H, incl, w0, cur, true, nxt, ws);
        st_evals++;
    }
}

template <int KIND, bool MULTI, int MODE>
__device__ __forceinline__ void zz_run_body(const ZzParams& P)
{
    ZzDevCtl* C = P.ctl;
#ifdef ZZ_PROF_NODE
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        zz_dbg_ptr = C->dbg;
#ifdef ZZ_PROF_TAIL
        zz_dbg_on = 0;
#else
        zz_dbg_on = 1;
#endif
    }
    __syncthreads();
#endif
    __shared__ ZzTailList tail_list;
    __shared__ ZzXres tail_xr;   // sharded tail: result of CTA 0's last exchange
    __shared__ unsigned int sq_cnt2[2];   // entries of this CTA's scan queue / base of its bulk append
    unsigned int* const sq_cnt_p = sq_cnt2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nthreads = gridDim.x * blockDim.x;
    const unsigned int gwarp = gtid >> 5, nwarps = nthreads >> 5, nwc = blockDim.x >> 5;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    const int32_t lo = MULTI ? P.v.lo : 0, hi = MULTI ? P.v.hi : P.v.d;   // owned coordinates
    unsigned long long epoch = 0;
    unsigned long long xep = MULTI ? __ldcg(&C->xrelease) : 0ULL;        // cross-GPU boundary counter (persists)

    // every thread of every GPU keeps an identical copy of the controller; decisions only depend on values that are
    // stable between two boundaries (and, MULTI, reduced over all GPUs)
    ZzCtl ctl; uint32_t cur, li, wat;
    if (__ldcg(&C->started)) {
        ctl = C->ctl; cur = C->cur; li = C->itg; wat = C->wattempt;
    } else {
        double F0 = zz_unkey(__ldcg(&C->f0_key));
        zz_ctl_init(ctl, F0 < P.T ? F0 : P.T, P.T, P.delta0, P.target, P.target_flips);
        cur = 0; li = 0; wat = 0;
    }
    unsigned long long nprop_prev = (unsigned long long)P.target | ((unsigned long long)P.target_flips << 32);
    unsigned int windows_done = 0;
    unsigned long long st_iters = 0, st_retries = 0, st_evals = 0, st_rebases = 0;
    bool stop = false;
    unsigned long long profbuf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    unsigned long long* prof = leader ? profbuf : nullptr;
    unsigned long long tmark = 0;
#define ZZ_TIC() do { if (prof) tmark = zz_now(); } while (0)
#define ZZ_TOC(k) do { if (prof) prof[k] += zz_now() - tmark; } while (0)

    while (ctl.phase < ZZ_PH_DONE && !stop) {
        if (cur > P.tag_limit) {  // iteration tags are about to run out of bits: forget all of them
            for (int32_t j = lo + (int32_t)gtid; j < hi; j += nthreads) {
                reinterpret_cast<unsigned long long*>(P.v.kin + j)[3] = 0ULL;
                P.dstamp[j] = 0;
            }
            zz_boundary<MULTI>(P, epoch, xep, prof, nullptr, nullptr, nullptr, 0u, 0u);
            cur = 0; st_rebases++;
        }
        const ZzCtl saved = ctl;
        zz_ctl_begin(ctl);
        const uint32_t w0 = cur + 1;
        cur = w0;
        const int ws = (int)(wat % 3u);
        wat++;
        const double H = ctl.H; const int incl = ctl.incl;

        // ---------------- pass 1: scan + evaluate every owned coordinate with a proposal inside the window
        int nxt = (int)((li + 1) % 3u);
        ZZ_TIC();
        if (leader) {
            C->wl_cnt[(li + 2) % 3u] = 0; C->issued[(li + 2) % 3u] = 0;
            const int wz = (int)(wat % 3u);  // slot of the NEXT attempt
            C->touched_cnt[wz] = 0; C->smin_key[wz] = ~0ULL; C->nprop_win[wz] = 0;
        }
        zz_queue_pass<KIND, MULTI, MODE>(P, sq_cnt_p, lo, hi, li, H, incl, w0, cur, nxt, ws, st_evals);
        ZZ_TOC(0);
        ZzXres xr = zz_boundary<MULTI>(P, epoch, xep, prof, MULTI ? &C->issued[nxt] : nullptr, nullptr,
                                       MULTI ? &C->wl_cnt[nxt] : nullptr, ZZ_OVF_BIT, ZZ_X_OVERFLOW);
        st_iters++;

        // ---------------- relaxation passes
        bool overflow = false;
        for (;;) {
            const unsigned int cwl = __ldcg(&C->wl_cnt[nxt]);            // this GPU's share of the next list
            const unsigned int cw = cwl & ~ZZ_OVF_BIT;
            const unsigned long long total = MULTI ? xr.sum : (unsigned long long)cw;
            if (MULTI ? (xr.flags & ZZ_X_OVERFLOW) != 0u : (cwl & ZZ_OVF_BIT) != 0u) { overflow = true; li = (li + 1) % 3u; break; }
            if (total == 0) break;
            if (!MULTI && cw <= ZZ_TAIL) {
                // Few coordinates left (typically a couple of hot neighbours resolving a long causal chain one
                // event per pass): CTA 0 runs these passes alone with block barriers; the others wait once.
                if (blockIdx.x == 0) {
                    ZZ_TIC();
                    unsigned int n = cw;
                    // Lattice kernels keep the tail's lists in shared memory (a coordinate marks at most four readers, so
                    // 4 * ZZ_TAIL entries always suffice); general graphs use the global lists.
                    ZzTailList* const tl = (KIND == ZZ_KIND_GRID) ? &tail_list : nullptr;
                    int src = -1, dst = 0;          // src < 0: the first tail pass reads the global list
                    if (tl && threadIdx.x == 0) tl->ovf = 0u;
#if defined(ZZ_PROF_NODE) && defined(ZZ_PROF_TAIL)
                    if (threadIdx.x == 0) zz_dbg_on = 1;
#endif
                    for (;;) {
                        li = (li + 1) % 3u;
                        nxt = (int)((li + 1) % 3u);
                        cur++;
                        if (threadIdx.x == 0) { C->wl_cnt[(li + 2) % 3u] = 0; if (tl) tl->n[dst] = 0u; }
                        if (tl) __syncthreads();
                        const int32_t* wlt = P.wl[li];
                        // entry e goes to warp e % (#warps), lane e / (#warps): few lanes per warp, so that coordinates with
                        // different timelines do not serialise each other's branches
                        for (unsigned int e = (unsigned int)lane * nwc + (unsigned int)warp; e < n; e += blockDim.x) {
                            const int32_t j = (tl && src >= 0) ? tl->buf[src][e] : __ldcg(wlt + e);
                            zz_eval_publish<KIND, MULTI, MODE>(P, j, H, incl, w0, cur, false, nxt, ws, tl, dst);
                            st_evals++;
                        }
                        __syncthreads();   // block-wide visibility is enough inside the tail; the grid barrier below fences
                        st_iters++;
                        if (prof) prof[7] += 1;
                        if (tl) {
                            n = tl->n[dst];
                            src = dst; dst ^= 1;
                            if (n == 0 || n > ZZ_TAIL || tl->ovf) break;
                        } else {
                            n = __ldcg(&C->wl_cnt[nxt]);
                            if (n == 0 || n > ZZ_TAIL) break;  // (an overflow bit makes n > ZZ_TAIL)
                        }
                    }
#if defined(ZZ_PROF_NODE) && defined(ZZ_PROF_TAIL)
                    if (threadIdx.x == 0) zz_dbg_on = 0;
#endif
                    if (tl) {   // hand the pending entries (if any) back to the global list the grid-wide passes read
                        for (unsigned int e = threadIdx.x; e < n; e += blockDim.x) P.wl[nxt][e] = tl->buf[src][e];
                        if (threadIdx.x == 0 && n) atomicAdd(&C->wl_cnt[nxt], n);   // keeps an overflow bit set by a publication
                    }
                    if (threadIdx.x == 0) { C->tail_li = li; C->tail_cur = cur; }
                    ZZ_TOC(2);
                }
                zz_grid_barrier(C, epoch, prof);
                li = __ldcg(&C->tail_li); cur = __ldcg(&C->tail_cur);
                nxt = (int)((li + 1) % 3u);
                continue;
            }
            if (MULTI && total <= ZZ_TAIL) {
                // Sharded tail: every GPU's CTA 0 relaxes its own few coordinates and the CTAs 0 exchange the number of
                // issued marks directly -- no local grid barrier, no release broadcast per pass; the other CTAs wait once.
                if (blockIdx.x == 0) {
                    ZZ_TIC();
                    unsigned long long txep = xep;
                    for (;;) {
                        li = (li + 1) % 3u;
                        nxt = (int)((li + 1) % 3u);
                        cur++;
                        if (threadIdx.x == 0) { C->wl_cnt[(li + 2) % 3u] = 0; C->issued[(li + 2) % 3u] = 0; }
                        const unsigned int n = __ldcg(&C->wl_cnt[li]) & ~ZZ_OVF_BIT;   // this GPU's share of the list
                        const int32_t* wlt = P.wl[li];
                        for (unsigned int e = (unsigned int)lane * nwc + (unsigned int)warp; e < n; e += blockDim.x) {
                            const int32_t j = __ldcg(wlt + e);
                            zz_eval_publish<KIND, MULTI, MODE>(P, j, H, incl, w0, cur, false, nxt, ws);
                            st_evals++;
                        }
                        __syncthreads();
                        if (threadIdx.x == 0) {
                            const unsigned long long ms = __ldcg(&C->issued[nxt]);
                            const unsigned int mf = (__ldcg(&C->wl_cnt[nxt]) & ZZ_OVF_BIT) ? ZZ_X_OVERFLOW : 0u;
                            txep += 1;
                            const ZzXres t = zz_exchange(P, txep, ms, ~0ULL, mf);
                            tail_xr.sum = t.sum; tail_xr.minkey = t.minkey; tail_xr.flags = t.flags;
                        }
                        __syncthreads();
                        st_iters++;
                        if (prof) prof[7] += 1;
                        if (tail_xr.sum == 0ULL || tail_xr.sum > ZZ_TAIL || (tail_xr.flags & ZZ_X_OVERFLOW)) break;
                    }
                    if (threadIdx.x == 0) {
                        C->tail_li = li; C->tail_cur = cur; C->tail_xep = txep;
                        C->tail_xres.sum = tail_xr.sum; C->tail_xres.minkey = tail_xr.minkey; C->tail_xres.flags = tail_xr.flags;
                    }
                    ZZ_TOC(2);
                }
                zz_grid_barrier(C, epoch, prof);
                li = __ldcg(&C->tail_li); cur = __ldcg(&C->tail_cur); xep = __ldcg(&C->tail_xep);
                xr.sum = __ldcg(&C->tail_xres.sum); xr.minkey = __ldcg(&C->tail_xres.minkey); xr.flags = __ldcg(&C->tail_xres.flags);
                nxt = (int)((li + 1) % 3u);
                continue;
            }
            li = (li + 1) % 3u;
            nxt = (int)((li + 1) % 3u);
            cur++;
            ZZ_TIC();
            if (leader) { C->wl_cnt[(li + 2) % 3u] = 0; C->issued[(li + 2) % 3u] = 0; }
            const int32_t* wl = P.wl[li];
            // same spreading over all warps of the grid: a list of n entries occupies ceil(n / #warps) lanes of every warp
            for (unsigned int e = (unsigned int)lane * nwarps + gwarp; e < cw; e += nthreads) {
                const int32_t j = __ldcg(wl + e);
                zz_eval_publish<KIND, MULTI, MODE>(P, j, H, incl, w0, cur, false, nxt, ws);
                st_evals++;
            }
            ZZ_TOC(1);
            xr = zz_boundary<MULTI>(P, epoch, xep, prof, MULTI ? &C->issued[nxt] : nullptr, nullptr,
                                    MULTI ? &C->wl_cnt[nxt] : nullptr, ZZ_OVF_BIT, ZZ_X_OVERFLOW);
            st_iters++;
        }
        cur++;  // tag of the commit pass: every list written in this window is visible to it

        // ---------------- phase B: time of the earliest accepted flip in the (trial) window
        double smin = ZZ_INF;
        if (!overflow && ctl.phase == ZZ_PH_B) {
            ZZ_TIC();
            const unsigned int nt = __ldcg(&C->touched_cnt[ws]);
            unsigned long long kmin = ~0ULL;
            for (unsigned int e = gtid; e < nt; e += nthreads) {
                const int32_t j = __ldcg(P.touched[0] + e);
                const ZzSpecR s = zz_load_spec(P.spec + j);
                if (s.nflip) {
                    double th, tf, xf; uint32_t h0, h1;
                    zz_ld_kin(P.v.kin + j, th, tf, xf, h0, h1);
                    int slot;
                    zz_pick_slot(h0, h1, w0, cur, slot);
                    const unsigned long long k = zz_key(__ldcg(P.v.flips + ((size_t)j * 2 + slot) * ZZ_MAXFLIP));
                    kmin = k < kmin ? k : kmin;
                }
            }
            if (kmin != ~0ULL) atomicMin(&C->smin_key[ws], kmin);
            ZZ_TOC(5);
            const ZzXres xb = zz_boundary<MULTI>(P, epoch, xep, prof, nullptr, &C->smin_key[ws], nullptr, 0u, 0u);
            if (xb.minkey != ~0ULL) smin = zz_unkey(xb.minkey);
        }

        ZzCtl trial = ctl;
        const int act = zz_ctl_end(trial, overflow, smin, nprop_prev);
        if (!MULTI && act == ZZ_ACT_COMMIT && P.record_trace) {
            // every event of this window (plus its end marker) must fit; otherwise hand the buffer to the host first
            const unsigned long long tl = __ldcg(&C->trace_len);
            const unsigned long long need = (unsigned long long)__ldcg(&C->touched_cnt[ws]) * ZZ_MAXFLIP + 1ULL;
            if (tl + need > P.trace_cap) {
                ctl = saved;
                if (leader) C->need_drain = 1u;
                break;
            }
        }
        ctl = trial;
        if (act == ZZ_ACT_COMMIT) {
            ZZ_TIC();
            const unsigned int nt = __ldcg(&C->touched_cnt[ws]);
            unsigned int np = 0, nf = 0;
            // Single GPU: a coordinate can be listed twice (the bulk append of the pass-1 queue and a pass-1 mark of a neighbour
            // that arrived before the coordinate's own stamp); the first visitor claims it with the commit tag (larger than
            // every stamp of the window: the relaxation only ends after a pass that issued no mark).
            for (unsigned int e = gtid; e < nt; e += nthreads) {
                const int32_t j = __ldcg(P.touched[0] + e);
                const ZzSpecR sp = zz_load_spec(P.spec + j);   // (issued before the claim so that the two round trips overlap)
                if (!MULTI && atomicMax(P.dstamp + j, cur) >= cur) continue;
                zz_commit_node<MODE>(P, j, sp, w0, cur, np, nf);
            }
            cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
            np = cg::reduce(w, np, cg::plus<unsigned int>());
            nf = cg::reduce(w, nf, cg::plus<unsigned int>());
            if (lane == 0 && (np | nf)) {
                // proposals (low half) and accepted flips (high half) of this window, for the length controller
                atomicAdd(&C->nprop_win[ws], (unsigned long long)np | ((unsigned long long)nf << 32));
                atomicAdd(&C->num, (unsigned long long)np);
                atomicAdd(&C->nacc, (unsigned long long)nf);
            }
            ZZ_TOC(3);
            // stop word: bound violation or (defensive) dropped trace record on ANY GPU stops all of them
            const ZzXres xc = zz_boundary<MULTI>(P, epoch, xep, prof, &C->nprop_win[ws], nullptr, &C->viol, 0xffffffffu, ZZ_X_STOP);
            if (leader && P.record_trace) {  // window-end marker (i = 0): lets the host sort window by window
                const unsigned long long pos = atomicAdd(&C->trace_len, 1ULL);
                if (pos < P.trace_cap) {
                    double2* e = reinterpret_cast<double2*>(P.trace + pos);
                    e[0] = make_double2(H, __longlong_as_double(0LL));
                    e[1] = make_double2(0.0, 0.0);
                } else {
                    C->trace_full = 1u;
                }
            }
            nprop_prev = xc.sum;
            windows_done++;
            if (xc.flags & ZZ_X_STOP) stop = true;
            if (P.max_windows && windows_done >= P.max_windows) stop = true;
        } else {
            st_retries++;
        }
    }

    if (leader) {
        C->ctl = ctl; C->cur = cur; C->itg = li; C->wattempt = wat; C->started = 1u;
        for (int q = 0; q < 8; ++q) C->tprof[q] += profbuf[q];
        C->windows += windows_done; C->retries += st_retries; C->iters += st_iters; C->rebases += st_rebases;
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
    st_evals = cg::reduce(w, st_evals, cg::plus<unsigned long long>());
    if (lane == 0 && st_evals) atomicAdd(&C->node_evals, st_evals);
}


// =====================================================================================================================
// Asynchronous tile-local relaxation (round 2).  Same windows, same per-coordinate timelines, same fixed point -- but no
// grid barrier per relaxation pass.  Every CTA owns a contiguous tile of coordinates and relaxes it to a LOCAL fixed point
// with block barriers only: work queues and dedupe bits live in shared memory, a coordinate whose list of accepted flips
// changed queues its readers for the CTA's next round (Gauss-Seidel: readers that have not started yet simply see the new
// list).  Only marks that cross a tile boundary leave the SM: they are pushed into the owning CTA's inbox.  One counter
// (`pending` = evaluations queued or in flight anywhere) detects quiescence of the whole window; two grid barriers per
// WINDOW (agree on the outcome, publish the committed frontier) replace one per pass.
//
// Why any order of evaluation is exact: the converged state is the unique fixed point "every published list equals the
// timeline of its coordinate under the published lists of its neighbours" (DESIGN.md section 3); a reader is re-queued
// AFTER the list it reads is complete (publisher: data, fence, mark; consumer: clear the mark, fence, read), so an
// evaluation that raced with a publication -- even one that saw a half-written list -- is always followed by one that
// did not, and `pending` cannot reach zero before that one has finished.
#define ZZ_TAG_STRIDE 256u   // list tags one window attempt may use (a coordinate publishes at most once per evaluation)
#define ZZ_SQCAP 2048u       // queue entries kept in shared memory; longer queues continue in the CTA's slice of P.wl[]

#define ZZ_SCAN_B 16         // 32-coordinate words per warp and scan batch

__device__ __forceinline__ void zz_prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

struct ZzAsyncSh {
    int32_t sq[2][ZZ_SQCAP];
    unsigned int n[2];        // entries of the two queues
    unsigned int tcount;      // coordinates of this tile evaluated at least once in the window (commit list)
    unsigned int head[ZZ_MAXRANKS];   // inbox entries consumed, per sending rank
    unsigned int dups;        // inbox entries dropped because the coordinate was queued already
    unsigned int state;       // decision of the polling thread
    unsigned int aborted;
    unsigned int wsum[32];    // per-warp totals of the scan's prefix sum
};
#define ZZ_ST_WORK 1u
#define ZZ_ST_DONE 2u
#define ZZ_ST_ABORT 3u

__device__ __forceinline__ unsigned int zz_ld_acq32(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned int zz_ld_acq32_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// the abort flag of a sharded run is raised in every rank's copy (rare), so that everybody polls local memory only
template <bool MULTI> __device__ __forceinline__ void zz_abort_raise(const ZzParams& P, int ws)
{
    if (MULTI) { for (int r = 0; r < P.v.nranks; ++r) atomicExch_system(&P.ctl_peer[r]->abortf[ws], 1u); }
    else atomicExch(&P.ctl->abortf[ws], 1u);
}

struct ZzTile {
    int32_t lo, per, c_lo, c_hi;
    int32_t* gq[2];       // global continuation of the two queues (this CTA's slice)
    int32_t* tlist;       // this CTA's slice of the touched list
    unsigned int* dirty;  // bit per coordinate of the tile: queued for a re-evaluation that has not started
    unsigned int* tbits;  // bit per coordinate: already on the commit list
};

__device__ __forceinline__ void zz_q_put(ZzAsyncSh& S, const ZzTile& t, int buf, unsigned int pos, int32_t k)
{
    if (pos < ZZ_SQCAP) S.sq[buf][pos] = k; else t.gq[buf][pos] = k;
}
__device__ __forceinline__ int32_t zz_q_get(const ZzAsyncSh& S, const ZzTile& t, int buf, unsigned int pos)
{
    return pos < ZZ_SQCAP ? S.sq[buf][pos] : __ldcg(t.gq[buf] + pos);
}

// Queue coordinate k of THIS tile into queue `buf` unless it is queued already; returns false for a duplicate.
__device__ __forceinline__ bool zz_mark_local(ZzAsyncSh& S, const ZzTile& t, int32_t k, int buf)
{
    const unsigned int li = (unsigned int)(k - t.c_lo), w = li >> 5, b = 1u << (li & 31u);
    const unsigned int old = atomicOr(&t.dirty[w], b);
    if (old & b) return false;
    const unsigned int pos = atomicAdd(&S.n[buf], 1u);
    zz_q_put(S, t, buf, pos, k);
    const unsigned int ot = atomicOr(&t.tbits[w], b);
    if (!(ot & b)) t.tlist[atomicAdd(&S.tcount, 1u)] = k;
    return true;
}

// Marks of one changed coordinate, in two steps around ONE fence (zz_publish_async): count the readers that live in other
// tiles (possibly on another GPU) -- they are added to `pending` before the fence, because the receiver may finish the
// evaluation and subtract it as soon as it sees the entry -- then queue the readers: those of this tile into the CTA's other
// queue, the others into their owners' inboxes.  An inbox is split by SENDING rank and the slot counter of a sender lives in the
// sender's own memory, so a push is one local atomic plus one store (over NVLink when the owner is another GPU): nothing on
// this path waits for a round trip to a peer.
template <int NK, bool MULTI>
__device__ __forceinline__ void zz_mark_count(const ZzParams& P, const ZzTile& t, const int32_t (&kk)[NK], int ws, unsigned int& nrem, bool& offgpu)
{
#pragma unroll
    for (int q = 0; q < NK; ++q) {
        if (kk[q] < 0 || (unsigned int)(kk[q] - t.c_lo) < (unsigned int)(t.c_hi - t.c_lo)) continue;
        if (MULTI && (kk[q] < P.v.lo || kk[q] >= P.v.hi)) {   // queued for a tile of another GPU: counted there
            atomicAdd_system(&P.ctl_peer[kk[q] / P.v.shard]->created[ws], 1ULL);
            offgpu = true;
        } else {
            nrem++;
        }
    }
}

template <int NK, bool MULTI>
__device__ __forceinline__ void zz_mark_push(const ZzParams& P, ZzAsyncSh& S, const ZzTile& t, const int32_t (&kk)[NK],
                                             int nxt, int ws, uint32_t wat)
{
    ZzDevCtl* C = P.ctl;
    const int nr = MULTI ? P.v.nranks : 1, me = MULTI ? P.v.rank : 0;
#pragma unroll
    for (int q = 0; q < NK; ++q) {
        if (kk[q] < 0) continue;
        if ((unsigned int)(kk[q] - t.c_lo) < (unsigned int)(t.c_hi - t.c_lo)) { zz_mark_local(S, t, kk[q], nxt); continue; }
        int r = 0; unsigned int owner; unsigned long long* box = P.inbox;
        if (MULTI) {
            r = kk[q] / P.v.shard;
            owner = (unsigned int)(kk[q] - r * P.v.shard) / (unsigned int)t.per;
            box = P.inbox_peer[r];
        } else {
            owner = (unsigned int)(kk[q] - t.lo) / (unsigned int)t.per;
        }
        const unsigned int pos = atomicAdd(P.inbox_cnt + ((size_t)ws * nr + r) * gridDim.x + owner, 1u);
        if (pos < P.inbox_cap)
            *(volatile unsigned long long*)(box + ((size_t)owner * nr + me) * P.inbox_cap + pos) = ((unsigned long long)wat << 32) | (unsigned int)kk[q];
        else {
            zz_abort_raise<MULTI>(P, ws);
            atomicAdd(&C->dbg[2], 1ULL);   // inbox full
        }
    }
}

// Lattice kernels relax in Gauss-Seidel order over the checkerboard: queue 0 holds the coordinates with (row + column) even,
// queue 1 the odd ones, and the CTA alternates between them.  All readers of a coordinate have the other colour, so two
// coordinates that read each other are never evaluated in the same round -- which is what makes pass-synchronous (Jacobi)
// relaxation chase phantom flips for dozens of rounds when two neighbours keep cancelling each other's speculative flips.
template <int KIND>
__device__ __forceinline__ int zz_colour(const ZzParams& P, int32_t k, int other)
{
    if (KIND != ZZ_KIND_GRID) return other;   // general graphs: whatever is not being processed (pass-synchronous inside the tile)
    const int32_t col = zz_grid_col(P.g, k);
    return (int)((unsigned int)(k - col * P.g.grid_m + col) & 1u);
}

// Warp 0 moves the valid prefixes of this CTA's inbox (one sub-box per sending rank) into the queues (lattice: by colour;
// otherwise into queue `buf`).  An entry is valid when it carries the number of the current window attempt.
template <int KIND, bool MULTI>
__device__ __forceinline__ void zz_drain_inbox(const ZzParams& P, ZzAsyncSh& S, const ZzTile& t, int buf, int ws, uint32_t wat)
{
    const unsigned int lane = threadIdx.x & 31u;
    const int nr = MULTI ? P.v.nranks : 1;
    for (int sr = 0; sr < nr; ++sr) {
        unsigned int head = S.head[sr];
        const unsigned long long* box = P.inbox + ((size_t)blockIdx.x * nr + sr) * P.inbox_cap;
        for (;;) {
            const unsigned int e = head + lane;
            bool ok = false; int32_t k = -1;
            if (e < P.inbox_cap) {
                const unsigned long long v = MULTI ? zz_ld_acq_sys(box + e) : zz_ld_acq(box + e);
                ok = ((uint32_t)(v >> 32) == wat);
                k = (int32_t)(uint32_t)v;
            }
            const unsigned int m = __ballot_sync(0xffffffffu, ok);
            const unsigned int nvalid = (m == 0xffffffffu) ? 32u : (unsigned int)(__ffs((int)~m) - 1);
            if (lane < nvalid && !zz_mark_local(S, t, k, zz_colour<KIND>(P, k, buf))) atomicAdd(&S.dups, 1u);
            head += nvalid;
            if (nvalid < 32u) break;   // the next entry has not been written (yet)
        }
        __syncwarp();
        if (lane == 0) S.head[sr] = head;
    }
    (void)ws;
}

// Has any sender left an entry that has not been consumed?  (lanes 0 .. nranks-1 of warp 0 look at one sub-box each)
template <bool MULTI>
__device__ __forceinline__ bool zz_inbox_nonempty(const ZzParams& P, const ZzAsyncSh& S, uint32_t wat)
{
    const unsigned int lane = threadIdx.x & 31u;
    const int nr = MULTI ? P.v.nranks : 1;
    bool ok = false;
    if ((int)lane < nr && S.head[lane] < P.inbox_cap) {
        const unsigned long long* e = P.inbox + ((size_t)blockIdx.x * nr + lane) * P.inbox_cap + S.head[lane];
        const unsigned long long v = MULTI ? zz_ld_acq_sys(e) : zz_ld_acq(e);
        ok = ((uint32_t)(v >> 32) == wat);
    }
    return __any_sync(0xffffffffu, ok);
}

// Sharded runs: every rank keeps a replica of the records it reads from other ranks at the SAME global index of its own
// (full-length) arrays; the owner pushes every change -- published lists, committed anchors -- into the replicas of the ranks
// that read the coordinate, so that no evaluation ever loads over NVLink.  Bit r of the result: rank r (other than the owner)
// reads coordinate j.
template <int KIND>
__device__ __forceinline__ unsigned int zz_reader_ranks(const ZzParams& P, int32_t j)
{
    unsigned int m = 0;
    if (KIND == ZZ_KIND_GRID) {   // whole lattice columns per rank: only the j -+ M readers can live elsewhere
        const int32_t M = P.g.grid_m;
        if (j - M >= 0) m |= 1u << ((j - M) / P.v.shard);
        if (j + M < P.v.d) m |= 1u << ((j + M) / P.v.shard);
    } else {
        const int32_t q1 = P.dptr[j + 1];
        for (int32_t q = P.dptr[j]; q < q1; ++q) m |= 1u << (P.didx[q] / P.v.shard);
    }
    return m & ~(1u << P.v.rank);
}

template <int KIND, bool MULTI, int MODE>
__device__ __forceinline__ void zz_publish_async(const ZzParams& P, ZzAsyncSh& S, const ZzTile& t, int32_t j, const ZzNodeOut& o,
                                                 uint32_t w0, int nxt, int ws, uint32_t wat)
{
    ZzDevCtl* C = P.ctl;
    int slot;
    const uint32_t cnt = zz_pick_slot(o.hdr0, o.hdr1, w0, 0xffffffffu, slot);
    bool same = (cnt == o.nflip);
    uint32_t flags = o.flags;
    if (same && cnt) {
        const double* fl = P.v.flips + ((size_t)j * 2 + slot) * ZZ_MAXFLIP;
#pragma unroll
        for (int m = 0; m < ZZ_MAXFLIP; ++m)
            if (m < (int)cnt) same = same && (zz_d2u(__ldcg(fl + m)) == zz_d2u(o.fl[m]));
        if (ZZ_MODE_HAS_VEL(MODE)) {
            const double* ft = P.v.fth + ((size_t)j * 2 + slot) * ZZ_MAXFLIP;
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m)
                if (m < (int)cnt) same = same && (zz_d2u(__ldcg(ft + m)) == zz_d2u(o.fth[m]));
        }
    }
    if (!same) {
        const int wsl = (slot == 0) ? 1 : 0;
        const uint32_t newtag = (slot < 0) ? w0 : (((slot == 0) ? o.hdr0 : o.hdr1) >> 4) + 1u;
        if (newtag - w0 >= ZZ_TAG_STRIDE) {
            flags |= ZZ_F_OVERFLOW;   // out of tags for this attempt: retry the window shorter
            atomicAdd(&C->dbg[1], 1ULL);
            C->dbg[2] = (unsigned long long)j | ((unsigned long long)S.state << 32) | ((unsigned long long)o.nflip << 48);   // (S.state: rounds of this window, ZZ_PROF_SCAN)
        } else {
            double* fl = P.v.flips + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
            for (int m = 0; m < ZZ_MAXFLIP; ++m)
                if (m < (int)o.nflip) fl[m] = o.fl[m];
            if (ZZ_MODE_HAS_VEL(MODE)) {
                double* ft = P.v.fth + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
                for (int m = 0; m < ZZ_MAXFLIP; ++m)
                    if (m < (int)o.nflip) ft[m] = o.fth[m];
            }
            reinterpret_cast<volatile uint32_t*>(P.v.kin + j)[6 + wsl] = (newtag << 4) | o.nflip;
            if (MULTI) {   // the same list and header into the replicas of the ranks that read j (plain stores over NVLink)
                for (unsigned int rm = zz_reader_ranks<KIND>(P, j); rm; rm &= rm - 1u) {
                    const int rr = __ffs((int)rm) - 1;
                    double* flr = P.v.flips_peer[rr] + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
                    for (int m = 0; m < ZZ_MAXFLIP; ++m)
                        if (m < (int)o.nflip) flr[m] = o.fl[m];
                    if (ZZ_MODE_HAS_VEL(MODE)) {
                        double* ftr = P.v.fth_peer[rr] + ((size_t)j * 2 + wsl) * ZZ_MAXFLIP;
#pragma unroll
                        for (int m = 0; m < ZZ_MAXFLIP; ++m)
                            if (m < (int)o.nflip) ftr[m] = o.fth[m];
                    }
                    reinterpret_cast<volatile uint32_t*>(P.v.kin_peer[rr] + j)[6 + wsl] = (newtag << 4) | o.nflip;
                }
            }
#ifdef ZZ_PROF_SCAN
            if (P.dbgbuf && newtag - w0 >= 16u && newtag - w0 < 40u) {   // development: who keeps re-publishing?
                const unsigned long long pos = atomicAdd(&C->dbg[7], 1ULL);
                if (pos < 4096ULL) {
                    unsigned long long* r = P.dbgbuf + (size_t)gridDim.x * ZZ_DBG_REC * 4 + pos * 16;
                    r[0] = (unsigned long long)j; r[1] = newtag - w0; r[2] = o.nflip | ((unsigned long long)cnt << 8) | ((unsigned long long)(slot + 1) << 16) | ((unsigned long long)S.state << 32);
                    r[3] = zz_d2u(o.fl[0]); r[4] = zz_d2u(o.fl[1]);
                    r[5] = (unsigned long long)o.hdr0 | ((unsigned long long)o.hdr1 << 32);
                    const double* flo = P.v.flips + ((size_t)j * 2 + (slot < 0 ? 0 : slot)) * ZZ_MAXFLIP;
                    r[6] = zz_d2u(__ldcg(flo)); r[7] = zz_d2u(__ldcg(flo + 1));
                    const int32_t nb[4] = { j - P.g.grid_m, j - 1, j + 1, j + P.g.grid_m };
                    for (int q = 0; q < 4; ++q) r[8 + q] = (unsigned long long)__double_as_longlong(__ldcg(reinterpret_cast<const double*>(P.v.kin + nb[q]) + 3));
                    r[12] = zz_d2u(o.tau); r[13] = w0; r[14] = o.nprop; r[15] = o.nitems;
                }
            }
#endif
            // The list and its header must be VISIBLE (in L2; in the replicas of other GPUs) before any reader is marked: a reader
            // that raced with the stores above (and may have seen the new header with the old contents of the slot) cleared its
            // mark before it read, so the marks below re-queue it; a reader that clears its mark after them reads complete data.
            // A block-scope fence is not enough even for readers of this CTA: everybody reads through L2 (ld.cg).  The same
            // fence orders the increment of `pending` before the inbox entries.
            unsigned int nrem = 0; bool offgpu = false;
            if (KIND == ZZ_KIND_GRID) {
                const int32_t M = P.g.grid_m, N = P.g.grid_n;
                int32_t kk[4];
                if (o.interior) {
                    kk[0] = j - M; kk[1] = j - 1; kk[2] = j + 1; kk[3] = j + M;
                } else {
                    const int32_t col = zz_grid_col(P.g, j), row = j - col * M;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const bool ok = (q == 0) ? (col > 0) : (q == 1) ? (row > 0) : (q == 2) ? (row < M - 1) : (col < N - 1);
                        kk[q] = ok ? j + ((q == 0) ? -M : (q == 1) ? -1 : (q == 2) ? 1 : M) : -1;
                    }
                }
                zz_mark_count<4, MULTI>(P, t, kk, ws, nrem, offgpu);
                if (nrem) { atomicAdd(&C->created[ws], (unsigned long long)nrem); atomicAdd(&C->dbg[3], (unsigned long long)nrem); }
                if (MULTI && offgpu) __threadfence_system(); else __threadfence();
                zz_mark_push<4, MULTI>(P, S, t, kk, nxt, ws, wat);
            } else {
                const int32_t q1 = P.dptr[j + 1];
                for (int32_t q0 = P.dptr[j]; q0 < q1; q0 += 4) {
                    int32_t kk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) kk[q] = (q0 + q < q1) ? P.didx[q0 + q] : -1;
                    zz_mark_count<4, MULTI>(P, t, kk, ws, nrem, offgpu);
                }
                if (nrem) { atomicAdd(&C->created[ws], (unsigned long long)nrem); atomicAdd(&C->dbg[3], (unsigned long long)nrem); }
                if (MULTI && offgpu) __threadfence_system(); else __threadfence();
                for (int32_t q0 = P.dptr[j]; q0 < q1; q0 += 4) {
                    int32_t kk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) kk[q] = (q0 + q < q1) ? P.didx[q0 + q] : -1;
                    zz_mark_push<4, MULTI>(P, S, t, kk, nxt, ws, wat);
                }
            }
        }
    }
    ZzNodeOut oo = o; oo.flags = flags;
    zz_store_spec(P.spec + j, oo);
    if (MODE == ZZ_MODE_REFRESH) reinterpret_cast<double2*>(P.v.rspec)[j] = make_double2(o.tprop, o.tref);
    if (flags & ZZ_F_VIOL) {
        double* vi = P.viol_info + (size_t)j * 3;
        vi[0] = o.viol_t; vi[1] = o.viol_l; vi[2] = o.viol_lb;
    }
    if (flags & ZZ_F_OVERFLOW) {
        zz_abort_raise<MULTI>(P, ws);
        if (o.flags & ZZ_F_OVERFLOW) atomicAdd(&C->dbg[0], 1ULL);   // flips / pool / items of the timeline itself
    }
}

#define ZZ_DBGLOG(kind, cnt) do { if (dbg_on && threadIdx.x == 0 && dbg_n < ZZ_DBG_REC) { unsigned long long* _r = P.dbgbuf + ((size_t)blockIdx.x * ZZ_DBG_REC + dbg_n) * 4; \
    _r[0] = (unsigned long long)(kind); _r[1] = (unsigned long long)(cnt); _r[2] = zz_now(); _r[3] = (unsigned long long)clock64(); dbg_n++; } } while (0)

#ifndef ZZ_PRESCAN
#define ZZ_PRESCAN 1
#endif
template <int KIND, bool MULTI, int MODE>
__device__ __forceinline__ void zz_run_body_async(const ZzParams& P)
{
    ZzDevCtl* C = P.ctl;
    unsigned int dbg_n = 0;
    __shared__ ZzAsyncSh S;
    extern __shared__ unsigned int zz_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int nwc = blockDim.x >> 5;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    const int32_t lo = MULTI ? P.v.lo : 0, hi = MULTI ? P.v.hi : P.v.d;   // owned coordinates
    unsigned long long xep = MULTI ? __ldcg(&C->xrelease) : 0ULL;         // cross-GPU boundary counter (persists)
    ZzTile t;
    t.lo = lo;
    t.per = P.tile_per;
    t.c_lo = lo + (int32_t)blockIdx.x * t.per; if (t.c_lo > hi) t.c_lo = hi;
    t.c_hi = (t.c_lo + t.per < hi) ? t.c_lo + t.per : hi;
    t.gq[0] = P.wl[0] + t.c_lo; t.gq[1] = P.wl[1] + t.c_lo; t.tlist = P.touched[0] + t.c_lo;
    t.dirty = zz_dyn; t.tbits = zz_dyn + P.flag_words;
    unsigned long long epoch = 0;

    ZzCtl ctl; uint32_t cur, wat;
    if (__ldcg(&C->started)) {
        ctl = C->ctl; cur = C->cur;
    } else {
        double F0 = zz_unkey(__ldcg(&C->f0_key));
        zz_ctl_init(ctl, F0 < P.T ? F0 : P.T, P.T, P.delta0, P.target, P.target_flips);
        cur = 0;
    }
    wat = __ldcg(&C->wattempt);   // (the host starts every run of a handle with fresh attempt numbers: they tag the inbox entries)
    unsigned long long nprop_prev = (unsigned long long)P.target | ((unsigned long long)P.target_flips << 32);
    unsigned int windows_done = 0;
    unsigned long long st_iters = 0, st_retries = 0, st_evals = 0, st_rebases = 0;
    bool stop = false;
    unsigned long long profbuf[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    unsigned long long* prof = leader ? profbuf : nullptr;
    unsigned long long tmark = 0;

    // slots of the first attempt of this launch (later ones are prepared one attempt ahead, see below)
    if (threadIdx.x == 0) {
        const int ws0 = (int)(wat % 3u);
        for (int r = 0; r < (MULTI ? P.v.nranks : 1); ++r)   // this rank's slot counters for pushes INTO tile blockIdx.x of rank r
            P.inbox_cnt[((size_t)ws0 * (MULTI ? P.v.nranks : 1) + r) * gridDim.x + blockIdx.x] = 0u;
        if (blockIdx.x == 0) {
            C->created[ws0] = (unsigned long long)gridDim.x; C->done[ws0] = 0ULL; C->abortf[ws0] = 0u; C->doneflag[ws0] = 0u;
            C->touched_cnt[ws0] = 0; C->smin_key[ws0] = ~0ULL; C->nprop_win[ws0] = 0;
        }
    }
    {   // (sharded: the earliest initial proposal over ALL ranks starts the first window; every rank initialised its own slab)
        const ZzXres x0r = zz_boundary<MULTI>(P, epoch, xep, prof, nullptr, &C->f0_key, nullptr, 0u, 0u);
        if (MULTI && !__ldcg(&C->started)) {
            const double F0 = zz_unkey(x0r.minkey);
            zz_ctl_init(ctl, F0 < P.T ? F0 : P.T, P.T, P.delta0, P.target, P.target_flips);
        }
    }

    // Attempt state.  The prologue of an attempt (controller step, slot resets, scan of the tile, tokens) is a lambda because it
    // runs in one of two places: at the top of the loop, or -- after a commit -- BEFORE the barrier that publishes the committed
    // frontier: the scan reads only this tile's own proposal times, which the CTA has just committed itself, so it overlaps with
    // the wait for the other tiles ("pre-scan").  Evaluations start after the barrier, as before.
    ZzCtl saved = ctl;
    uint32_t w0 = 0, watn = 0; int ws = 0; double H = 0.0; int incl = 0; bool dbg_on = false;
    bool prescanned = false;
    auto begin_attempt = [&]() {
        saved = ctl;
        zz_ctl_begin(ctl);
        w0 = cur + 1;
        cur = w0 + ZZ_TAG_STRIDE;   // every tag of this attempt is in [w0, w0 + ZZ_TAG_STRIDE)
        ws = (int)(wat % 3u);
        watn = wat;   // attempt number stamped into inbox entries
        wat++;
        H = ctl.H; incl = ctl.incl;
        dbg_on = P.dbgbuf && windows_done == P.dbg_window;
        ZZ_DBGLOG(1, wat);

        // ---------------- scan: the coordinates of this tile with a proposal inside the window
        ZZ_TIC();
        if (threadIdx.x == 0) {
            const int wz = (int)(wat % 3u);   // slots of the NEXT attempt
            for (int r = 0; r < (MULTI ? P.v.nranks : 1); ++r)
                P.inbox_cnt[((size_t)wz * (MULTI ? P.v.nranks : 1) + r) * gridDim.x + blockIdx.x] = 0u;
            if (blockIdx.x == 0) {
                C->created[wz] = (unsigned long long)gridDim.x; C->done[wz] = 0ULL; C->abortf[wz] = 0u; C->doneflag[wz] = 0u;
                C->touched_cnt[wz] = 0; C->smin_key[wz] = ~0ULL; C->nprop_win[wz] = 0;
            }
            S.n[0] = 0u; S.n[1] = 0u; S.tcount = 0u; S.dups = 0u;
            for (int r = 0; r < ZZ_MAXRANKS; ++r) S.head[r] = 0u;
             S.aborted = 0u; S.state = 0u;
        }
        __syncthreads();
#ifdef ZZ_PROF_SCAN
        const long long sc0 = clock64(); long long sc1 = sc0, sc2 = sc0;
#endif
        // Phase A: warp w takes the 32-coordinate words w, w + #warps, ... of the tile, ZZ_SCAN_B words per batch (all loads of a
        // batch in flight together); the ballot of a word IS its bit word (queued = touched = proposal inside the window).
        const unsigned int nwords = (unsigned int)(t.c_hi - t.c_lo + 31) >> 5;
        for (unsigned int wb = (unsigned int)warp; wb < nwords; wb += nwc * ZZ_SCAN_B) {
            double tj[ZZ_SCAN_B];
#pragma unroll
            for (int u = 0; u < ZZ_SCAN_B; ++u) {
                const unsigned int w = wb + (unsigned int)u * nwc;
                const int32_t j = t.c_lo + (int32_t)(w << 5) + lane;
                tj[u] = (w < nwords && j < t.c_hi) ? __ldcg(P.v.tau + j) : ZZ_INF;
            }
#pragma unroll
            for (int u = 0; u < ZZ_SCAN_B; ++u) {
                const unsigned int w = wb + (unsigned int)u * nwc;
                const unsigned int m = __ballot_sync(0xffffffffu, (tj[u] < H) || (incl && tj[u] == H));
                if (lane == 0 && w < nwords) { t.dirty[w] = m; t.tbits[w] = m; }
            }
        }
        __syncthreads();
#ifdef ZZ_PROF_SCAN
        sc1 = clock64();
#endif
        // Phase B: compaction.  One THREAD per word walks its set bits (about six): count per queue (lattice: per colour), block-wide
        // exclusive scan of the counts, write the entries.  No atomics, a few dozen instructions per word.
        for (unsigned int w0c = 0; w0c < nwords; w0c += blockDim.x) {
            const unsigned int w = w0c + threadIdx.x;
            unsigned int m = (w < nwords) ? t.tbits[w] : 0u;
            const int32_t j0 = t.c_lo + (int32_t)(w << 5);
            unsigned int redm = m;   // bits that go to queue 0
            if (KIND == ZZ_KIND_GRID) {
                redm = 0u;
                for (unsigned int mm = m; mm; mm &= mm - 1u) {
                    const int b = __ffs((int)mm) - 1;
                    if (zz_colour<KIND>(P, j0 + b, 0) == 0) redm |= 1u << b;
                }
            }
            const unsigned int nr = (unsigned int)__popc(redm), nb = (unsigned int)__popc(m & ~redm);
            // exclusive scan of (nr | nb << 16) over the block (counts per chunk stay below 2^16: 32 per word, <= 1024 words)
            unsigned int v = nr | (nb << 16), incl_v = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl_v, off);
                if (lane >= off) incl_v += up;
            }
            if (lane == 31) S.wsum[warp] = incl_v;
            __syncthreads();
            unsigned int pre = 0, tot = 0;
            for (unsigned int q = 0; q < nwc; ++q) { const unsigned int x = S.wsum[q]; if (q < (unsigned int)warp) pre += x; tot += x; }
            const unsigned int excl = pre + incl_v - v;
            unsigned int pr = S.n[0] + (excl & 0xffffu), pb = S.n[1] + (excl >> 16), pt = S.tcount + (excl & 0xffffu) + (excl >> 16);
            for (unsigned int mm = m; mm; mm &= mm - 1u) {
                const int b = __ffs((int)mm) - 1;
                const int32_t j = j0 + b;
                if (redm & (1u << b)) zz_q_put(S, t, 0, pr++, j); else zz_q_put(S, t, 1, pb++, j);
                t.tlist[pt++] = j;
            }
            __syncthreads();
            if (threadIdx.x == 0) { S.n[0] += tot & 0xffffu; S.n[1] += tot >> 16; S.tcount += (tot & 0xffffu) + (tot >> 16); }
            __syncthreads();
        }
#ifdef ZZ_PROF_SCAN
        sc2 = clock64();
#endif
        __syncthreads();
#ifdef ZZ_PROF_SCAN
        if (leader) { const long long sc3 = clock64(); C->dbg[4] += 1; C->dbg[5] += (unsigned long long)(sc1 - sc0); C->dbg[6] += (unsigned long long)(sc3 - sc1); }
#endif
        if (threadIdx.x == 0) {
            // queued evaluations first, then hand back this CTA's token (the second atomic consumes the result of the first, so
            // that it cannot be performed earlier: created >= done at all times)
            const unsigned long long oc = atomicAdd(&C->created[ws], (unsigned long long)S.tcount);
            atomicAdd(&C->done[ws], 1ULL + (oc >> 63));
        }
        ZZ_TOC(0);
        ZZ_DBGLOG(2, S.tcount);
    };

    while (ctl.phase < ZZ_PH_DONE && !stop) {
        if (!prescanned) {
            if (cur > P.tag_limit) {  // list tags are about to run out of bits: forget all of them

                if (MULTI) {   // own records and the replicas of the other ranks' records alike (nobody publishes before the boundary below)
                    for (int32_t j = (int32_t)(blockIdx.x * blockDim.x + threadIdx.x); j < P.v.d; j += (int32_t)(gridDim.x * blockDim.x))
                        reinterpret_cast<unsigned long long*>(P.v.kin + j)[3] = 0ULL;
                } else {
                    for (int32_t j = t.c_lo + (int32_t)threadIdx.x; j < t.c_hi; j += blockDim.x)
                        reinterpret_cast<unsigned long long*>(P.v.kin + j)[3] = 0ULL;
                }
                zz_boundary<MULTI>(P, epoch, xep, prof, nullptr, nullptr, nullptr, 0u, 0u);
                cur = 0; st_rebases++;
            }
            begin_attempt();
        }
        prescanned = false;

        // ---------------- local rounds until the whole window is quiescent
        // Invariant at the top: queue `cq` is complete (local marks of the previous round + drained inbox entries), every
        // thread of the CTA is past a block barrier.  Two block barriers per round.
        int cq = 0;
        unsigned int outcome = ZZ_ST_ABORT;
        for (;;) {
            const unsigned int n = S.n[cq];
            if (n == 0 && S.n[cq ^ 1] != 0u) { cq ^= 1; continue; }   // (uniform: read after a block barrier)
            if (n == 0) {
                if (warp == 0) {
                    ZZ_TIC();
                    unsigned int st = 0;
                    for (;;) {
                        if (zz_inbox_nonempty<MULTI>(P, S, watn)) { st = ZZ_ST_WORK; break; }
                        // order matters: `done` before `created` (see ZzDevCtl); an evaluation that overflowed raised the flag
                        // BEFORE it was added to `done`, so whoever sees the sums agree also sees the flag
                        unsigned long long dsum = 0, csum = 1; unsigned int ab = 0, fl = 0;
                        if (!MULTI) {
                            if (lane == 0) { dsum = zz_ld_acq(&C->done[ws]); csum = zz_ld_acq(&C->created[ws]); ab = zz_ld_acq32(&C->abortf[ws]); }
                        } else if (blockIdx.x == 0) {   // CTA 0 compares the node-wide sums (one peer load per lane and counter) ...
                            const bool have = lane < (unsigned int)P.v.nranks;
                            unsigned long long dv = have ? zz_ld_acq_sys(&P.ctl_peer[lane]->done[ws]) : 0ULL;
                            dv = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), dv, cg::plus<unsigned long long>());
                            unsigned long long cv = have ? zz_ld_acq_sys(&P.ctl_peer[lane]->created[ws]) : 0ULL;
                            cv = cg::reduce(cg::tiled_partition<32>(cg::this_thread_block()), cv, cg::plus<unsigned long long>());
                            dsum = dv; csum = cv;
                            if (lane == 0) {
                                ab = zz_ld_acq32_sys(&C->abortf[ws]);
                                if (dsum == csum && !ab) { __threadfence(); atomicExch(&C->doneflag[ws], 1u); }
                            }
                        } else {                        // ... the other CTAs of the GPU wait for its verdict
                            if (lane == 0) { ab = zz_ld_acq32_sys(&C->abortf[ws]); fl = zz_ld_acq32(&C->doneflag[ws]); if (fl) csum = dsum; }
                        }
                        dsum = __shfl_sync(0xffffffffu, dsum, 0); csum = __shfl_sync(0xffffffffu, csum, 0); ab = __shfl_sync(0xffffffffu, ab, 0);
                        if (ab) { st = ZZ_ST_ABORT; break; }
                        if (dsum == csum) { st = ZZ_ST_DONE; break; }
                    }
                    if (lane == 0) S.state = st;
                    ZZ_TOC(2);
                    ZZ_DBGLOG(4, st);
                }
                __syncthreads();
                if (S.state != ZZ_ST_WORK) { outcome = S.state; break; }
                if (warp == 0) {
                    zz_drain_inbox<KIND, MULTI>(P, S, t, cq, ws, watn);
                    if (lane == 0 && S.dups) {
                        atomicAdd(&C->done[ws], (unsigned long long)S.dups);
                        S.dups = 0u;
                    }
                }
                __syncthreads();
                continue;
            }
            ZZ_TIC();
            ZZ_DBGLOG(3, n);
#ifdef ZZ_PROF_SCAN
            if (threadIdx.x == 0) S.state = (S.state >= 1000u ? S.state : 1000u) + 1u;
#endif
#ifdef ZZ_PROF_TAIL2
            const long long q0 = clock64();
#endif
            const int nq = cq ^ 1;
            const unsigned int other0 = S.n[nq];   // (lattice: the other colour may hold entries of the scan or of the inbox already)
            // the inbox counter and the abort flag are needed only AFTER the evaluations: start the loads now (plain loads:
            // the entries themselves are read with acquire semantics, the flag is only a hint to stop early)
            unsigned int abort_pre = 0;
            if (threadIdx.x == 0) abort_pre = *(volatile const unsigned int*)&C->abortf[ws];
            // entry e goes to warp e % (#warps), lane e / (#warps): a short queue occupies a few lanes of every warp
            const unsigned int estride = P.eval_threads ? P.eval_threads : blockDim.x;
            for (unsigned int e = (unsigned int)lane * nwc + (unsigned int)warp; e < n; e += estride) {
                if ((unsigned int)lane * nwc + (unsigned int)warp >= estride) break;
                const int32_t j = zz_q_get(S, t, cq, e);
                const unsigned int li = (unsigned int)(j - t.c_lo);
                atomicAnd(&t.dirty[li >> 5], ~(1u << (li & 31u)));
                __threadfence_block();   // the mark is cleared before anything is read (a later publication re-queues j)
                ZzNodeOut o;
#ifdef ZZ_PROF_TAIL2
                const long long c0 = clock64();
#endif
                if constexpr (MODE == ZZ_MODE_LOGIT) zz_process_node_logit(P.g, P.v, P.lg, j, H, incl, w0, 0xffffffffu, false, o);
                else if constexpr (MODE == ZZ_MODE_STRONG) zz_process_node_strong(P.g, P.v, P.st, j, H, incl, w0, 0xffffffffu, false, o);
                else zz_process_node_k<KIND, MODE, false>(P.g, P.v, j, H, incl, w0, 0xffffffffu, false, o);   // (sharded: local replicas)
#ifdef ZZ_PROF_TAIL2
                const long long c1 = clock64();
#endif
                zz_publish_async<KIND, MULTI, MODE>(P, S, t, j, o, w0, nq, ws, watn);
#ifdef ZZ_PROF_TAIL2
                if (n <= 8 && e == 0) {   // lone evaluations of tail rounds: cycles in the timeline / in the publication
                    const long long c2 = clock64();
                    atomicAdd(&C->dbg[4], 1ULL); atomicAdd(&C->dbg[5], (unsigned long long)(c1 - c0)); atomicAdd(&C->dbg[6], (unsigned long long)(c2 - c1));
                    atomicAdd(&C->dbg[7], (unsigned long long)o.nitems);
                }
#endif
                st_evals++;
            }
            __syncthreads();   // every publication and every local mark of this round is done
            if (warp == 0) {
                const unsigned int produced = S.n[nq] - other0;   // local marks only: inbox entries were counted by their senders
                if (lane == 0) S.n[cq] = 0u;                          // (before the drain: on the lattice it may refill this queue)
                __syncwarp();
                if (zz_inbox_nonempty<MULTI>(P, S, watn)) zz_drain_inbox<KIND, MULTI>(P, S, t, nq, ws, watn);
                if (lane == 0) {
                    const unsigned long long oc = atomicAdd(&C->created[ws], (unsigned long long)produced);   // (performed before `done`)
                    atomicAdd(&C->done[ws], (unsigned long long)n + (unsigned long long)S.dups + (oc >> 63));
                    S.dups = 0u;
                    if (abort_pre) S.aborted = 1u;
                }
            }
            cq = nq;
            st_iters++;
            if (prof) prof[7] += 1;
            ZZ_TOC(1);
            __syncthreads();
#ifdef ZZ_PROF_TAIL2
            if (threadIdx.x == 0 && n <= 8u) {   // whole tail round, thread 0 of every CTA
                atomicAdd(&C->tprof[5], (unsigned long long)(clock64() - q0)); atomicAdd(&C->dbg[2], 1ULL);
            }
#endif
            if (S.aborted) break;
        }

        // ---------------- the outcome is known to every CTA without a barrier: the window converged (`pending` reached zero:
        // no evaluation queued or in flight anywhere, so every list and every speculative end state is final) or it was aborted
        // (the flag is raised before the overflowing evaluation is subtracted from `pending`, and never lowered).  A grid-wide
        // (sharded: node-wide) barrier is needed only to agree on the earliest flip of a phase-B trial window, or on the room
        // left in the trace buffer.
        const bool overflow = (outcome != ZZ_ST_DONE);
        const unsigned int tcount = S.tcount;
        const bool need_b1 = (ctl.phase == ZZ_PH_B) || (!MULTI && P.record_trace);
        double smin = ZZ_INF;
        if (need_b1) {
            if (threadIdx.x == 0 && tcount) atomicAdd(&C->touched_cnt[ws], tcount);
            if (!overflow && ctl.phase == ZZ_PH_B) {
                ZZ_TIC();
                unsigned long long kmin = ~0ULL;
                for (unsigned int e = threadIdx.x; e < tcount; e += blockDim.x) {
                    const int32_t j = __ldcg(t.tlist + e);
                    const ZzSpecR s = zz_load_spec(P.spec + j);
                    if (s.nflip) {
                        double th, tf, xf; uint32_t h0, h1;
                        zz_ld_kin(P.v.kin + j, th, tf, xf, h0, h1);
                        int slot;
                        zz_pick_slot(h0, h1, w0, 0xffffffffu, slot);
                        const unsigned long long k = zz_key(__ldcg(P.v.flips + ((size_t)j * 2 + slot) * ZZ_MAXFLIP));
                        kmin = k < kmin ? k : kmin;
                    }
                }
                if (kmin != ~0ULL) atomicMin(&C->smin_key[ws], kmin);
                ZZ_TOC(5);
            }
            ZZ_DBGLOG(5, tcount);
            const ZzXres xb = zz_boundary<MULTI>(P, epoch, xep, prof, nullptr, &C->smin_key[ws], nullptr, 0u, 0u);
            ZZ_DBGLOG(6, 0);
            if (!overflow && ctl.phase == ZZ_PH_B && xb.minkey != ~0ULL) smin = zz_unkey(xb.minkey);
        }

        ZzCtl trial = ctl;
        const int act = zz_ctl_end(trial, overflow, smin, nprop_prev);
        if (!MULTI && act == ZZ_ACT_COMMIT && P.record_trace) {
            // every event of this window must fit; otherwise hand the buffer to the host first
            const unsigned long long tl = __ldcg(&C->trace_len);
            const unsigned long long need = (unsigned long long)__ldcg(&C->touched_cnt[ws]) * ZZ_MAXFLIP + 1ULL;
            if (tl + need > P.trace_cap) {
                ctl = saved;
                if (leader) C->need_drain = 1u;
                break;
            }
        }
        ctl = trial;
        if (act == ZZ_ACT_COMMIT) {
            ZZ_TIC();
            // Two passes, so that the lanes of a warp do the same thing: (1) the private record of every touched coordinate (two
            // dependent loads, four stores), collecting the coordinates WITH events -- about one in five -- in the (empty) queue
            // 0; (2) their kinematic records, moment sums and trace records, spread evenly over the CTA.  In one pass nearly every
            // warp iteration had some lane on the long path and 26 lanes waiting for it.
            unsigned int np = 0, nf = 0;
            for (unsigned int e = threadIdx.x; e < tcount; e += blockDim.x) {
                const int32_t j = __ldcg(t.tlist + e);
                const ZzSpecR sp = zz_load_spec(P.spec + j);
                zz_commit_light<MODE>(P, j, sp, np);
                if (sp.nflip) zz_q_put(S, t, 0, atomicAdd(&S.n[0], 1u), j);
            }
            __syncthreads();
            const unsigned int nev = S.n[0];
            for (unsigned int e = threadIdx.x; e < nev; e += blockDim.x) {
                const int32_t j = zz_q_get(S, t, 0, e);
                const ZzSpecR sp = zz_load_spec(P.spec + j);
                zz_commit_heavy<MODE, KIND, MULTI>(P, j, sp, w0, 0xffffffffu, nf);
            }
            cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
            np = cg::reduce(w, np, cg::plus<unsigned int>());
            nf = cg::reduce(w, nf, cg::plus<unsigned int>());
            if (lane == 0 && (np | nf)) {
                atomicAdd(&C->nprop_win[ws], (unsigned long long)np | ((unsigned long long)nf << 32));
                atomicAdd(&C->num, (unsigned long long)np);
                atomicAdd(&C->nacc, (unsigned long long)nf);
            }
            ZZ_TOC(3);
            ZZ_DBGLOG(7, np);
            // Pre-scan: the prologue of the next attempt (its window end follows from the controller state alone -- the proposal
            // count reduced at the barrier below enters one window later) runs before the barrier, on this tile's own freshly
            // committed proposal times.  Not when the run may end here, when the tags are about to be rebased (that needs a
            // barrier of its own first) or when a window is being logged.
            const int wsc = ws; const double Hc = H;          // the committed attempt's slot and window end
            const ZzCtl ctl_c = ctl; const uint32_t cur_c = cur, wat_c = wat;
            // The barrier is split: this CTA ARRIVES first (its committed frontier is published), scans, and only then waits.
            zz_grid_arrive(C, epoch);   // (its block barrier also makes the committed times of this tile visible to the whole CTA)
            if (ZZ_PRESCAN && ctl.phase < ZZ_PH_DONE && !(P.max_windows && windows_done + 1u >= P.max_windows) && !(cur > P.tag_limit) &&
                !P.dbgbuf) {
                begin_attempt();
                prescanned = true;
            }
            // proposals / flips of the window (length controller) and the stop word (bound violation on ANY GPU stops all of
            // them) are reduced on the way
            const ZzXres xc = zz_boundary<MULTI>(P, epoch, xep, prof, &C->nprop_win[wsc], nullptr, &C->viol, 0xffffffffu, ZZ_X_STOP, true);
            ZZ_DBGLOG(8, 0);
            if (leader && P.record_trace) {  // window-end marker (i = 0): lets the host sort window by window
                const unsigned long long pos = atomicAdd(&C->trace_len, 1ULL);
                if (pos < P.trace_cap) {
                    double2* e = reinterpret_cast<double2*>(P.trace + pos);
                    e[0] = make_double2(Hc, __longlong_as_double(0LL));
                    e[1] = make_double2(0.0, 0.0);
                } else {
                    C->trace_full = 1u;
                }
            }
            nprop_prev = xc.sum;
            windows_done++;
            if (xc.flags & ZZ_X_STOP) stop = true;
            if (P.max_windows && windows_done >= P.max_windows) stop = true;
            if (stop && prescanned) {   // the run ends here: forget the attempt that was only scanned (the next launch prepares its slots again)
                ctl = ctl_c; cur = cur_c; wat = wat_c; prescanned = false;
            }
        } else {
            st_retries++;
        }
    }

    if (leader) {
        C->ctl = ctl; C->cur = cur; C->itg = 0; C->wattempt = wat; C->started = 1u;
        for (int q = 0; q < 8; ++q) C->tprof[q] += profbuf[q];
        C->windows += windows_done; C->retries += st_retries; C->iters += st_iters; C->rebases += st_rebases;
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
    st_evals = cg::reduce(w, st_evals, cg::plus<unsigned long long>());
    if (lane == 0 && st_evals) atomicAdd(&C->node_evals, st_evals);
}

// ZZ_ASYNC 1 (default): single-GPU kernels run the asynchronous tile-local relaxation; the sharded (_multi) kernels and
// zz_run_kernel_grid_sync (A/B reference, zzb_run_set("schedule", 0)) run the pass-synchronous body of round 1.
#ifndef ZZ_ASYNC
#define ZZ_ASYNC 1
#endif
template <int KIND, bool MULTI, int MODE, bool ASYNC>
__device__ __forceinline__ void zz_run_dispatch(const ZzParams& P)
{
    if constexpr (ASYNC) zz_run_body_async<KIND, MULTI, MODE>(P);
    else zz_run_body<KIND, MULTI, MODE>(P);
}
#define ZZ_RUN_KERNEL(name, KIND, MULTI, MODE, ASYNC) \
    extern "C" __global__ void __launch_bounds__((KIND == ZZ_KIND_GRID ? ZZ_RUN_BLOCK_GRID : ZZ_RUN_BLOCK_CSR), ZZ_MINB) name(const __grid_constant__ ZzParams P) { zz_run_dispatch<KIND, MULTI, MODE, ASYNC>(P); }
ZZ_RUN_KERNEL(zz_run_kernel_grid, ZZ_KIND_GRID, false, ZZ_MODE_PLAIN, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr, ZZ_KIND_CSR, false, ZZ_MODE_PLAIN, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_multi, ZZ_KIND_GRID, true, ZZ_MODE_PLAIN, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_multi, ZZ_KIND_CSR, true, ZZ_MODE_PLAIN, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_lb, ZZ_KIND_GRID, false, ZZ_MODE_LB, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_lb, ZZ_KIND_CSR, false, ZZ_MODE_LB, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_multi_lb, ZZ_KIND_GRID, true, ZZ_MODE_LB, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_multi_lb, ZZ_KIND_CSR, true, ZZ_MODE_LB, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_sticky, ZZ_KIND_GRID, false, ZZ_MODE_STICKY, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_sticky, ZZ_KIND_CSR, false, ZZ_MODE_STICKY, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_boom, ZZ_KIND_GRID, false, ZZ_MODE_BOOM, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_boom, ZZ_KIND_CSR, false, ZZ_MODE_BOOM, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_logit, ZZ_KIND_CSR, false, ZZ_MODE_LOGIT, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_strong, ZZ_KIND_CSR, false, ZZ_MODE_STRONG, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_refresh, ZZ_KIND_GRID, false, ZZ_MODE_REFRESH, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_refresh, ZZ_KIND_CSR, false, ZZ_MODE_REFRESH, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_sticky_multi, ZZ_KIND_GRID, true, ZZ_MODE_STICKY, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_sticky_multi, ZZ_KIND_CSR, true, ZZ_MODE_STICKY, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_boom_multi, ZZ_KIND_GRID, true, ZZ_MODE_BOOM, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_csr_boom_multi, ZZ_KIND_CSR, true, ZZ_MODE_BOOM, ZZ_ASYNC)
ZZ_RUN_KERNEL(zz_run_kernel_grid_sync, ZZ_KIND_GRID, false, ZZ_MODE_PLAIN, 0)
ZZ_RUN_KERNEL(zz_run_kernel_csr_sync, ZZ_KIND_CSR, false, ZZ_MODE_PLAIN, 0)

#include "zz_seq.cuh"
