// zzb200.cpp -- host side of libzzb200.so: the C-ABI declared in include/zzb200.h on top of the CUDA DRIVER
// API.  libcuda is dlopen'ed on first use, so the library itself loads (and exports its symbols) on a machine
// without a GPU; every compute entry point then fails with ZZB_E_CUDA -- there is no CPU fallback.
//
// The kernels live in zzb200_kernels.cubin (sm_100a), loaded with cuModuleLoadData; the event loop is one
// persistent cooperative kernel (zz_run_kernel) launched with cuLaunchCooperativeKernel, re-launched only
// when the trace buffer has to be drained to the host.
#include <cuda.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zzb200.h"
#include "zz_dev.h"
#include "zz_host_graph.h"
#include "zz_host_logit.h"
#include "zz_host_seq.h"

#define ZZ_BLOCK 256

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int32_t fail(int32_t code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

// ------------------------------------------------------------------------------------------- driver loader
#define ZZ_STR2(x) #x
#define ZZ_STR(x) ZZ_STR2(x)
#define ZZ_DRV_FUNCS(X) \
    X(cuInit) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetName) X(cuDeviceGetAttribute) X(cuDeviceTotalMem) \
    X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease) X(cuCtxPushCurrent) X(cuCtxPopCurrent) \
    X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuMemAlloc) X(cuMemFree) X(cuMemcpyHtoD) \
    X(cuMemcpyDtoH) X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync) X(cuMemsetD8) X(cuMemsetD8Async) X(cuStreamCreate) \
    X(cuStreamDestroy) X(cuStreamSynchronize) X(cuLaunchKernel) X(cuLaunchCooperativeKernel) X(cuEventCreate) \
    X(cuEventDestroy) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime) X(cuGetErrorString) \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuMemGetInfo) X(cuMemHostAlloc) X(cuMemFreeHost) \
    X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle) X(cuModuleGetGlobal) X(cuFuncSetAttribute)

struct Drv {
#define X(name) decltype(&name) p_##name = nullptr;
    ZZ_DRV_FUNCS(X)
#undef X
    void* handle = nullptr;
};
static Drv g_drv;

static int32_t load_driver()
{
    if (g_drv.handle) return ZZB_OK;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(ZZB_E_CUDA, "CUDA driver library (libcuda.so.1) not found: %s -- zzb200 has no CPU fallback", dlerror());
#define X(name)                                                                                     \
    g_drv.p_##name = (decltype(&name))dlsym(h, ZZ_STR(name));                                       \
    if (!g_drv.p_##name) return fail(ZZB_E_CUDA, "symbol %s missing from the CUDA driver", ZZ_STR(name));
    ZZ_DRV_FUNCS(X)
#undef X
    g_drv.handle = h;
    return ZZB_OK;
}

static int32_t cu_fail(CUresult r, const char* what)
{
    const char* s = nullptr;
    if (g_drv.p_cuGetErrorString) g_drv.p_cuGetErrorString(r, &s);
    return fail(ZZB_E_CUDA, "%s failed: %s (CUresult %d)", what, s ? s : "?", (int)r);
}
#define CU(call)                                                     \
    do {                                                             \
        CUresult _r = g_drv.p_##call;                                \
        if (_r != CUDA_SUCCESS) return cu_fail(_r, #call);           \
    } while (0)

// --------------------------------------------------------------------------------------------- global state
#define ZZ_NKERN 22   // event-loop kernels in the image (see zzb_init)
#define ZZ_KERN_BLOCK_IDX(k) ((k) == 12 || (k) == 13 ? 1 : ((k) & 1))
#define ZZ_RUN_BLOCK_OF(r) ZZ_KERN_BLOCK_IDX((r)->kidx())      // the logistic / strong kernels are general-sparse kernels on any graph
#define ZZ_KERN_ASYNC(k) ((k) < 14 || (k) >= 16)   // asynchronous tile-local relaxation (zz_run_body_async)
struct Global {
    bool ready = false;
    CUdevice dev = 0; int dev_id = 0;
    CUcontext ctx = nullptr;
    CUmodule mod = nullptr;
    CUfunction f_init_strong = nullptr;
    CUfunction f_math_probe = nullptr;
    CUfunction f_seq = nullptr, f_seq_logit = nullptr;
    CUfunction f_cummean = nullptr;
    CUfunction f_ts_hist = nullptr, f_ts_scan = nullptr, f_ts_scatter = nullptr, f_ts_sort = nullptr;
    CUfunction f_setup = nullptr, f_init = nullptr, f_init_boom = nullptr, f_run[ZZ_NKERN] = {}, f_export = nullptr, f_grid_tail = nullptr;
    CUstream stream = nullptr;
    CUevent ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
    int sm_count = 0; int blocks_per_sm[ZZ_NKERN] = {}; int run_block[2] = { 256, 256 };   // lattice / general-sparse kernels
    size_t total_mem = 0; char name[128] = { 0 };
};
static Global G;
static std::mutex g_mu;

struct CtxGuard {
    bool pushed = false;
    CtxGuard() { if (G.ctx && g_drv.p_cuCtxPushCurrent(G.ctx) == CUDA_SUCCESS) pushed = true; }
    ~CtxGuard() { if (pushed) { CUcontext c; g_drv.p_cuCtxPopCurrent(&c); } }
};

struct DevBuf {
    CUdeviceptr p = 0; size_t n = 0;
    int32_t alloc(size_t bytes)
    {
        release();
        if (!bytes) bytes = 16;
        CUresult r = g_drv.p_cuMemAlloc(&p, bytes);
        if (r != CUDA_SUCCESS) { p = 0; return cu_fail(r, "cuMemAlloc"); }
        n = bytes;
        return ZZB_OK;
    }
    void release() { if (p) { g_drv.p_cuMemFree(p); p = 0; n = 0; } }
    ~DevBuf() { release(); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static int32_t upload(DevBuf& b, const void* src, size_t bytes)
{
    int32_t st = b.alloc(bytes);
    if (st) return st;
    if (bytes) CU(cuMemcpyHtoD(b.p, src, bytes));
    return ZZB_OK;
}

static std::string default_cubin_path()
{
    const char* env = getenv("ZZB200_CUBIN");
    if (env && *env) return env;
    Dl_info info;
    if (dladdr((void*)&default_cubin_path, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        p = (k == std::string::npos) ? std::string(".") : p.substr(0, k);
        return p + "/zzb200_kernels.cubin";
    }
    return "zzb200_kernels.cubin";
}

// ----------------------------------------------------------------------------------------------- handles
struct zzb_problem_s {
    ZzHostGraph hg;
    DevBuf nptr, nidx, nwt, nwb, nfl, gmu, h, dptr, didx;
    ZzGraph g;
    // subsampled logistic target (zzb_problem_create_logistic)
    bool logit = false;
    ZzHostLogit hl;
    DevBuf l_acp, l_arow, l_aval, l_rp, l_rcol, l_rval, l_y, l_ny, l_u0;
    ZzLogit lg;
    // sequential-chain schedule (zz_seq.cuh): compact matrices + connected components; hs.ok says whether it is available
    ZzHostSeq hs;
    DevBuf s_bcp, s_brow, s_bval, s_tcp, s_trow, s_tval, s_comp, s_ent, s_rowrec, s_orig, s_rcoln;
    ZzSeq sq;
};

struct zzb_run_s {
    zzb_problem_s* prob = nullptr;
    int32_t d = 0;
    uint32_t flags = 0;
    DevBuf kin, flips, priv, tau, kctr, spec, viol, dstamp, acc, s1, s2, wl[3], touched, trace, ctl;
    DevBuf in_x, in_th, in_c;          // staging of the inputs / outputs
    DevBuf out_t, out_x, out_th, out_c, out_acc;
    unsigned long long trace_cap = 0;
    ZzParams P;
    double t0 = 0, T = 0;
    uint64_t seed[2] = { 0, 0 }; int32_t adapt = 0; double factor = 1.8;
    double delta0 = 0, target_frac = 0.25, target_flip_frac = 0.045; unsigned int tag_limit = ZZ_TAG_LIMIT; unsigned int max_windows = 0;
    bool uploaded = false, executed = false, have_inputs = false;
    // results
    std::vector<zzb_event> events;     // sorted, markers removed
    ZzDevCtl hc;                       // last copy of the device control block
    int64_t launches = 0;
    int grid = 0; int kind = 1;
    bool strong = false; double strong_c = 0.0, kappa0 = 0.0; int strong_rule = 0;
    // -1: automatic (sequential chains for the logistic target when available, else 1); 2: sequential chains (zz_seq.cuh);
    // 1: asynchronous tile-local relaxation; 0: pass-synchronous schedule of round 1 (plain ZigZag only)
    int schedule = -1;
    bool seq_capable() const
    {
        return prob && prob->hs.ok && nranks <= 1 && !strong && !grid_n &&
               !(flags & (ZZB_FLAG_LOCAL_BOUND | ZZB_FLAG_STICKY | ZZB_FLAG_BOOMERANG | ZZB_FLAG_REFRESH));
    }
    int sched() const { return schedule >= 0 ? schedule : ((prob && (prob->logit || prob->hs.prefer) && seq_capable() && !max_windows) ? 2 : 1); }
    DevBuf inbox, inbox_cnt; unsigned int inbox_cap = 0, flag_words = 0; int inbox_grid = 0;
    DevBuf dbgbuf; std::vector<unsigned long long> dbghost;
    int tile_per = 0; unsigned int eval_threads = 0; int inbox_nr = 0;
    int seq_warps = 0;                 // sequential chains: warps per chain (0 = automatic)
    bool window_policy_set = false;    // target_frac / target_flip_frac were set by the caller
    // device-side ordering of the trace (zz_tsort_*): the events of the last execute, sorted, still in HBM
    DevBuf cm_work, cm_out;            // cummean on the device (zzb_trace_cummean)
    DevBuf trace_sorted, ts_work; bool dev_sorted = false; unsigned long long n_sorted = 0; bool host_sort_only = false;
    DevBuf trace_map, s3; bool have_map = false;   // subtrace filter / inclusion-time accumulator (sticky)
    int setup_lo = 0, setup_hi = 0;    // coordinates whose records this rank sets up (sharded lattice: slab + halo; otherwise all)
    unsigned int wat_next = 0;         // window-attempt numbers tag the inbox entries: never reused by a later run of this handle
    int kidx() const
    {
        if (strong) return 13;
        if (prob && prob->logit) return 12;
        if (flags & ZZB_FLAG_REFRESH) return 16 + kind;
        if (flags & ZZB_FLAG_BOOMERANG) return (nranks > 1 ? 20 : 10) + kind;
        if (flags & ZZB_FLAG_STICKY) return (nranks > 1 ? 18 : 8) + kind;
        if (!sched() && nranks <= 1 && !(flags & ZZB_FLAG_LOCAL_BOUND)) return 14 + kind;
        return kind + (nranks > 1 ? 2 : 0) + ((flags & ZZB_FLAG_LOCAL_BOUND) ? 4 : 0);
    }
    DevBuf dfth, kappa; bool have_kappa = false;
    DevBuf bmu, bsig; bool have_boom = false; double lambdaref = 0, rho = 0;
    DevBuf rsig, rst, rspec; bool have_refresh = false; double rlambdaref = 0;   // ZigZag with refreshments (ZZB_FLAG_REFRESH)
    DevBuf gridbuf; double grid_dt = 0; long long grid_n = 0;   // device-side discretize
    int rank = 0, nranks = 1, shard = 0, lo = 0, hi = 0;
    CUdeviceptr peer[ZZ_MAXRANKS][8] = {};   // imported mappings: kin, flips, dstamp, wl0, wl1, wl2, touched, ctl
    bool peer_open[ZZ_MAXRANKS] = {};
    bool fetched = false;
    std::vector<double> ft, fx, fth, fc; std::vector<long long> facc; std::vector<double> hs1, hs2;
};

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int32_t zzb_last_error(char* buf, int64_t len)
{
    if (!buf || len <= 0) return ZZB_E_ARG;
    snprintf(buf, (size_t)len, "%s", g_err.c_str());
    return ZZB_OK;
}

int32_t zzb_init(int32_t ndev, const int32_t* dev_ids, const char* cubin_path)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (G.ready) return ZZB_OK;
    if (ndev > 1) return fail(ZZB_E_ARG, "one device per process: shard coordinates across processes (see DESIGN.md, multi-GPU)");
    int32_t st = load_driver();
    if (st) return st;
    CU(cuInit(0));
    int count = 0;
    CU(cuDeviceGetCount(&count));
    if (count <= 0) return fail(ZZB_E_CUDA, "no CUDA device visible -- zzb200 has no CPU fallback");
    G.dev_id = (ndev >= 1 && dev_ids) ? dev_ids[0] : 0;
    if (G.dev_id < 0 || G.dev_id >= count) return fail(ZZB_E_ARG, "device id %d out of range (0..%d)", G.dev_id, count - 1);
    CU(cuDeviceGet(&G.dev, G.dev_id));
    CU(cuDeviceGetName(G.name, sizeof G.name, G.dev));
    CU(cuDeviceTotalMem(&G.total_mem, G.dev));
    int major = 0, minor = 0, coop = 0;
    CU(cuDeviceGetAttribute(&major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, G.dev));
    CU(cuDeviceGetAttribute(&minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, G.dev));
    CU(cuDeviceGetAttribute(&G.sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, G.dev));
    CU(cuDeviceGetAttribute(&coop, CU_DEVICE_ATTRIBUTE_COOPERATIVE_LAUNCH, G.dev));
    if (major != 10) return fail(ZZB_E_CUDA, "device %s is sm_%d%d; the kernels are built for sm_100a only", G.name, major, minor);
    if (!coop) return fail(ZZB_E_CUDA, "device does not support cooperative launch");
    CU(cuDevicePrimaryCtxRetain(&G.ctx, G.dev));  // shared with the CUDA runtime (torch) of this process
    CtxGuard cg;
    std::string path = cubin_path && *cubin_path ? std::string(cubin_path) : default_cubin_path();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return fail(ZZB_E_CUDA, "cannot open kernel image %s (build with __graft_entry__.build())", path.c_str());
    std::vector<char> img;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    img.resize((size_t)sz + 1);
    if (fread(img.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return fail(ZZB_E_CUDA, "short read on %s", path.c_str()); }
    fclose(f);
    CU(cuModuleLoadData(&G.mod, img.data()));
    CU(cuModuleGetFunction(&G.f_setup, G.mod, "zz_setup_kernel"));
    CU(cuModuleGetFunction(&G.f_init, G.mod, "zz_init_kernel"));
    CU(cuModuleGetFunction(&G.f_init_boom, G.mod, "zz_init_kernel_boom"));
    // index = kind (0 lattice, 1 general) + 2 * multi-GPU + 4 * LocalBound
    // 8, 9: sticky ZigZag (single GPU)
    // 10, 11: factorised Boomerang (single GPU)
    // 12: subsampled logistic target (general sparse, single GPU)
    static const char* run_names[ZZ_NKERN] = { "zz_run_kernel_grid", "zz_run_kernel_csr", "zz_run_kernel_grid_multi", "zz_run_kernel_csr_multi",
                                         "zz_run_kernel_grid_lb", "zz_run_kernel_csr_lb", "zz_run_kernel_grid_multi_lb", "zz_run_kernel_csr_multi_lb",
                                         "zz_run_kernel_grid_sticky", "zz_run_kernel_csr_sticky", "zz_run_kernel_grid_boom", "zz_run_kernel_csr_boom",
                                         "zz_run_kernel_csr_logit", "zz_run_kernel_csr_strong",
                                         "zz_run_kernel_grid_sync", "zz_run_kernel_csr_sync",      // 14, 15: round-1 schedule (A/B reference)
                                         "zz_run_kernel_grid_refresh", "zz_run_kernel_csr_refresh",     // 16, 17: ZigZag with refreshments
                                         "zz_run_kernel_grid_sticky_multi", "zz_run_kernel_csr_sticky_multi",   // 18, 19: sticky, sharded
                                         "zz_run_kernel_grid_boom_multi", "zz_run_kernel_csr_boom_multi" };     // 20, 21: Boomerang, sharded
    CU(cuModuleGetFunction(&G.f_init_strong, G.mod, "zz_init_kernel_strong"));
    for (int k = 0; k < ZZ_NKERN; ++k) CU(cuModuleGetFunction(&G.f_run[k], G.mod, run_names[k]));
    CU(cuModuleGetFunction(&G.f_export, G.mod, "zz_export_kernel"));
    CU(cuModuleGetFunction(&G.f_grid_tail, G.mod, "zz_grid_tail_kernel"));
    CU(cuModuleGetFunction(&G.f_math_probe, G.mod, "zz_math_probe_kernel"));
    CU(cuModuleGetFunction(&G.f_cummean, G.mod, "zz_cummean_kernel"));
    CU(cuModuleGetFunction(&G.f_seq, G.mod, "zz_seq_kernel"));
    CU(cuModuleGetFunction(&G.f_seq_logit, G.mod, "zz_seq_kernel_logit"));
    CU(cuModuleGetFunction(&G.f_ts_hist, G.mod, "zz_tsort_hist_kernel"));
    CU(cuModuleGetFunction(&G.f_ts_scan, G.mod, "zz_tsort_scan_kernel"));
    CU(cuModuleGetFunction(&G.f_ts_scatter, G.mod, "zz_tsort_scatter_kernel"));
    CU(cuModuleGetFunction(&G.f_ts_sort, G.mod, "zz_tsort_sort_kernel"));
    CU(cuStreamCreate(&G.stream, CU_STREAM_NON_BLOCKING));
    CU(cuEventCreate(&G.ev0, CU_EVENT_DEFAULT));
    CU(cuEventCreate(&G.ev1, CU_EVENT_DEFAULT));
    CU(cuEventCreate(&G.tev0, CU_EVENT_DEFAULT));
    CU(cuEventCreate(&G.tev1, CU_EVENT_DEFAULT));
    {   // block size the event-loop kernels were compiled for
        static const char* names[2] = { "zz_run_block_grid", "zz_run_block_csr" };
        for (int q = 0; q < 2; ++q) {
            CUdeviceptr sym = 0; size_t symsz = 0;
            if (g_drv.p_cuModuleGetGlobal(&sym, &symsz, G.mod, names[q]) == CUDA_SUCCESS && symsz == sizeof(int))
                CU(cuMemcpyDtoH(&G.run_block[q], sym, sizeof(int)));
        }
    }
    for (int k = 0; k < ZZ_NKERN; ++k) {
        CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&G.blocks_per_sm[k], G.f_run[k], G.run_block[ZZ_KERN_BLOCK_IDX(k)], 0));
        if (G.blocks_per_sm[k] < 1) return fail(ZZB_E_CUDA, "zz_run_kernel does not fit on an SM");
    }
    G.ready = true;
    return ZZB_OK;
}

int32_t zzb_shutdown(void)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!G.ready) return ZZB_OK;
    {
        CtxGuard cg;
        if (G.ev0) g_drv.p_cuEventDestroy(G.ev0);
        if (G.ev1) g_drv.p_cuEventDestroy(G.ev1);
        if (G.stream) g_drv.p_cuStreamDestroy(G.stream);
        if (G.mod) g_drv.p_cuModuleUnload(G.mod);
    }
    g_drv.p_cuDevicePrimaryCtxRelease(G.dev);
    G = Global();
    return ZZB_OK;
}

int32_t zzb_device_info(int32_t* sm_count, int64_t* total_mem, char* name, int64_t name_len)
{
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (sm_count) *sm_count = G.sm_count;
    if (total_mem) *total_mem = (int64_t)G.total_mem;
    if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", G.name);
    return ZZB_OK;
}

// CUDA events on the library's launching stream: which = 0 marks the start, 1 the end of a timed region.
int32_t zzb_event_record(int32_t which)
{
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    CtxGuard cg;
    CU(cuEventRecord(which ? G.tev1 : G.tev0, G.stream));
    return ZZB_OK;
}

int32_t zzb_event_elapsed_ms(float* ms)
{
    if (!G.ready || !ms) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    CtxGuard cg;
    CU(cuEventSynchronize(G.tev1));
    CU(cuEventElapsedTime(ms, G.tev0, G.tev1));
    return ZZB_OK;
}

// Sequential-chain schedule: upload the compact matrices and the component table when the problem qualifies (zz_host_seq.h).
#define ZZ_SEQ_RES_HOST 32u   // = ZZ_SEQ_RES of zz_seq.cuh
#define ZZ_SEQ_MAX_NC 2700   // coordinates per component: 80 bytes of shared memory each (+ scratch, checked in zz_build_seq)
static int32_t upload_seq(zzb_problem_s* p)
{
    const ZzHostSeq& hs = p->hs;
    memset(&p->sq, 0, sizeof p->sq);
    if (!hs.ok) return ZZB_OK;
    int32_t st = 0;
#define UPS(buf, vec) if (!st) st = upload(p->buf, hs.vec.data(), hs.vec.size() * sizeof(hs.vec[0]))
    UPS(s_bcp, bcp); UPS(s_brow, brow); UPS(s_bval, bval); UPS(s_comp, comp); UPS(s_orig, orig);
    if (hs.have_tgt) { UPS(s_tcp, tcp); UPS(s_trow, trow); UPS(s_tval, tval); }
#undef UPS
    if (st) return st;
    p->sq.bcp = p->s_bcp.as<int32_t>(); p->sq.brow = p->s_brow.as<int32_t>(); p->sq.bval = p->s_bval.as<double>();
    if (hs.have_tgt) { p->sq.tcp = p->s_tcp.as<int32_t>(); p->sq.trow = p->s_trow.as<int32_t>(); p->sq.tval = p->s_tval.as<double>(); }
    p->sq.comp = p->s_comp.as<int32_t>();
    bool ident = true;   // chains that are contiguous ranges already: no renumbering, 4 bytes of shared memory per coordinate less
    for (size_t j = 0; j < hs.orig.size() && ident; ++j) ident = (hs.orig[j] == (int32_t)j);
    p->sq.orig = ident ? nullptr : p->s_orig.as<int32_t>();
    p->sq.ncomp = (int32_t)hs.comp.size() - 1; p->sq.ncmax = hs.ncmax; p->sq.colmax = hs.colmax;
    if (p->logit) {   // packed design entries and per-row constants (the control-variate sigmoids are evaluated once, here)
        const ZzHostLogit& hl = p->hl;
        std::vector<ZzSeqEnt> ent(hl.arow.size());
        for (size_t e = 0; e < ent.size(); ++e) {
            const int32_t row = hl.arow[e];
            ZzSeqEnt& q = ent[e];
            q.row = row; q.q0 = hl.rp[row]; q.len = hl.rp[row + 1] - hl.rp[row]; q.pad = 0; q.val = hl.aval[e]; q.pad2 = 0.0;
        }
        std::vector<ZzSeqRow> rr((size_t)hl.n);
        for (int32_t r = 0; r < hl.n; ++r) { rr[r].y = hl.y[r]; rr[r].ny = hl.ny[r]; rr[r].sn0 = zz_sigmoidn(hl.u0[r]); rr[r].ns0 = zz_nsigmoid(hl.u0[r]); }
        std::vector<int32_t> rcoln(hl.rcol.size());
        for (size_t q = 0; q < rcoln.size(); ++q) rcoln[q] = hs.newid[hl.rcol[q]];
        st = upload(p->s_rcoln, rcoln.data(), rcoln.size() * sizeof(int32_t));
        if (!st) st = upload(p->s_ent, ent.data(), ent.size() * sizeof(ZzSeqEnt));
        if (!st) st = upload(p->s_rowrec, rr.data(), rr.size() * sizeof(ZzSeqRow));
        if (st) return st;
        p->sq.ent = p->s_ent.as<ZzSeqEnt>(); p->sq.rowrec = p->s_rowrec.as<ZzSeqRow>(); p->sq.rcoln = p->s_rcoln.as<int32_t>();
    }
    return ZZB_OK;
}

int32_t zzb_problem_create_gaussian(zzb_problem_t* out, int64_t d, const int64_t* colptr, const int64_t* rowval,
                                    const double* nzval, const double* hvec, const int64_t* bnd_colptr,
                                    const int64_t* bnd_rowval, const double* bnd_nzval, const double* bnd_mu)
{
    if (!out || !colptr || !rowval || !nzval) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded -- zzb200 has no CPU fallback");
    if (!bnd_colptr) { bnd_colptr = colptr; bnd_rowval = rowval; bnd_nzval = nzval; }
    std::vector<double> zeros;
    if (!bnd_mu) { zeros.assign((size_t)std::max<int64_t>(d, 1), 0.0); bnd_mu = zeros.data(); }
    zzb_problem_s* p = new zzb_problem_s();
    std::string e = zz_build_graph(p->hg, d, colptr, rowval, nzval, hvec, bnd_colptr, bnd_rowval, bnd_nzval, bnd_mu);
    if (!e.empty()) { delete p; return fail(ZZB_E_GRAPH, "%s", e.c_str()); }
    CtxGuard cg;
    const ZzHostGraph& hg = p->hg;
    int32_t st = 0;
#define UP(buf, vec) if (!st) st = upload(p->buf, hg.vec.data(), hg.vec.size() * sizeof(hg.vec[0]))
    UP(nptr, nptr); UP(nidx, nidx); UP(nwt, nwt); UP(nwb, nwb); UP(nfl, nfl); UP(gmu, gmu); UP(dptr, dptr); UP(didx, didx);
    if (hg.has_h) UP(h, h);
#undef UP
    if (st) { delete p; return st; }
    p->g.nptr = p->nptr.as<int32_t>(); p->g.nidx = p->nidx.as<int32_t>(); p->g.nwt = p->nwt.as<double>();
    p->g.nwb = p->nwb.as<double>(); p->g.nfl = p->nfl.as<uint8_t>(); p->g.gmu = p->gmu.as<double>();
    p->g.h = hg.has_h ? p->h.as<double>() : nullptr; p->g.same = hg.same;
    p->g.grid_m = getenv("ZZB200_NO_GRID") ? 0 : hg.grid_m; p->g.grid_n = hg.grid_n;
    for (int q = 0; q < 5; ++q) p->g.grid_diag[q] = hg.grid_diag[q];
    zz_grid_set_magic(p->g);
    if (d <= (1 << 18)) {   // small problems may run as sequential chains (zzb_run_set("schedule", 2))
        const bool sep = !hg.same;
        zz_build_seq(p->hs, d, bnd_colptr, bnd_rowval, bnd_nzval, sep ? colptr : nullptr, sep ? rowval : nullptr, sep ? nzval : nullptr,
                     nullptr, nullptr, ZZ_SEQ_MAX_NC);
        st = upload_seq(p);
        if (st) { delete p; return st; }
    } else p->hs.why = "prepared for d <= 262144 only";
    *out = p;
    return ZZB_OK;
}

// Target = the subsampled logistic-regression potential of scripts/logistic.jl (grad phi_j = gamma0 x_j - fdot_moving(...),
// :78-107, evaluated through the SelfMoving closure signature, src/sfact.jl:68); sampler = ZigZag(bnd, bnd_mu) as usual.
int32_t zzb_problem_create_logistic(zzb_problem_t* out, int64_t d, int64_t n, const int64_t* a_colptr, const int64_t* a_rowval,
                                    const double* a_nzval, const int64_t* at_colptr, const int64_t* at_rowval,
                                    const double* at_nzval, const double* y, const double* ny, const double* mu_cv,
                                    double gamma0, int64_t k, const int64_t* bnd_colptr, const int64_t* bnd_rowval,
                                    const double* bnd_nzval, const double* bnd_mu)
{
    if (!out || !a_colptr || !a_rowval || !a_nzval || !at_colptr || !at_rowval || !at_nzval || !y || !ny || !mu_cv ||
        !bnd_colptr || !bnd_rowval || !bnd_nzval)
        return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded -- zzb200 has no CPU fallback");
    std::vector<double> zeros;
    if (!bnd_mu) { zeros.assign((size_t)std::max<int64_t>(d, 1), 0.0); bnd_mu = zeros.data(); }
    zzb_problem_s* p = new zzb_problem_s();
    std::string e = zz_build_logit(p->hl, d, n, a_colptr, a_rowval, a_nzval, at_colptr, at_rowval, at_nzval, y, ny, mu_cv, gamma0, k);
    if (!e.empty()) { delete p; return fail(ZZB_E_ARG, "%s", e.c_str()); }
    // neighbour lists: the coordinates sharing a design row with j are its "target" entries (they only carry the
    // dependency: weight 0), the sampler matrix gives the bound and trigger entries as for a Gaussian target
    e = zz_build_graph(p->hg, d, p->hl.dep_cp.data(), p->hl.dep_rv.data(), p->hl.dep_nz.data(), nullptr, bnd_colptr, bnd_rowval,
                       bnd_nzval, bnd_mu);
    if (!e.empty()) { delete p; return fail(ZZB_E_GRAPH, "%s", e.c_str()); }
    CtxGuard cg;
    const ZzHostGraph& hg = p->hg;
    const ZzHostLogit& hl = p->hl;
    int32_t st = 0;
#define UP(buf, vec) if (!st) st = upload(p->buf, hg.vec.data(), hg.vec.size() * sizeof(hg.vec[0]))
    UP(nptr, nptr); UP(nidx, nidx); UP(nwt, nwt); UP(nwb, nwb); UP(nfl, nfl); UP(gmu, gmu); UP(dptr, dptr); UP(didx, didx);
#undef UP
#define UP(buf, vec) if (!st) st = upload(p->buf, hl.vec.data(), hl.vec.size() * sizeof(hl.vec[0]))
    UP(l_acp, acp); UP(l_arow, arow); UP(l_aval, aval); UP(l_rp, rp); UP(l_rcol, rcol); UP(l_rval, rval); UP(l_y, y); UP(l_ny, ny);
    UP(l_u0, u0);
#undef UP
    if (st) { delete p; return st; }
    memset(&p->g, 0, sizeof p->g);
    p->g.nptr = p->nptr.as<int32_t>(); p->g.nidx = p->nidx.as<int32_t>(); p->g.nwt = p->nwt.as<double>();
    p->g.nwb = p->nwb.as<double>(); p->g.nfl = p->nfl.as<uint8_t>(); p->g.gmu = p->gmu.as<double>();
    p->g.h = nullptr; p->g.same = 0; p->g.grid_m = 0; p->g.grid_n = 0;
    zz_grid_set_magic(p->g);
    p->hg.bnd_eq_tgt = false;
    p->logit = true;
    p->lg.acp = p->l_acp.as<int32_t>(); p->lg.arow = p->l_arow.as<int32_t>(); p->lg.aval = p->l_aval.as<double>();
    p->lg.rp = p->l_rp.as<int32_t>(); p->lg.rcol = p->l_rcol.as<int32_t>(); p->lg.rval = p->l_rval.as<double>();
    p->lg.y = p->l_y.as<double>(); p->lg.ny = p->l_ny.as<double>(); p->lg.u0 = p->l_u0.as<double>();
    p->lg.gamma0 = hl.gamma0; p->lg.k = hl.k; p->lg.n = hl.n;
    zz_build_seq(p->hs, d, bnd_colptr, bnd_rowval, bnd_nzval, nullptr, nullptr, nullptr, hl.dep_cp.data(), hl.dep_rv.data(), ZZ_SEQ_MAX_NC);
    st = upload_seq(p);
    if (st) { delete p; return st; }
    *out = p;
    return ZZB_OK;
}

int32_t zzb_problem_free(zzb_problem_t p)
{
    if (!p) return ZZB_OK;
    CtxGuard cg;
    delete p;
    return ZZB_OK;
}

int32_t zzb_run_create(zzb_problem_t p, uint32_t flags, int64_t trace_capacity_events, zzb_run_t* out)
{
    if (!p || !out) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (p->logit && (flags & (ZZB_FLAG_STICKY | ZZB_FLAG_LOCAL_BOUND | ZZB_FLAG_BOOMERANG)))
        return fail(ZZB_E_ARG, "the logistic target runs with the plain ZigZag sampler only (no sticky / LocalBound / Boomerang)");
    if ((flags & ZZB_FLAG_STICKY) && (flags & ZZB_FLAG_LOCAL_BOUND)) return fail(ZZB_E_ARG, "sticky and LocalBound cannot be combined");
    if ((flags & ZZB_FLAG_STICKY) && !p->g.grid_m && p->hg.maxdeg > ZZ_NB_WIDE)
        return fail(ZZB_E_ARG, "the sticky kernels handle columns of at most %d entries (this matrix has %d)", ZZ_NB_WIDE, p->hg.maxdeg);
    if ((flags & ZZB_FLAG_BOOMERANG) && (flags & (ZZB_FLAG_STICKY | ZZB_FLAG_LOCAL_BOUND))) return fail(ZZB_E_ARG, "Boomerang cannot be combined with sticky / LocalBound");
    if ((flags & ZZB_FLAG_REFRESH) && (flags & (ZZB_FLAG_STICKY | ZZB_FLAG_LOCAL_BOUND | ZZB_FLAG_BOOMERANG))) return fail(ZZB_E_ARG, "ZigZag refreshments cannot be combined with sticky / LocalBound / Boomerang");
    if ((flags & ZZB_FLAG_REFRESH) && p->logit) return fail(ZZB_E_ARG, "ZigZag refreshments are not available with the logistic target");
    if ((flags & ZZB_FLAG_REFRESH) && !p->g.grid_m && p->hg.maxdeg > ZZ_NB_WIDE)
        return fail(ZZB_E_ARG, "the refreshment kernels handle columns of at most %d entries (this matrix has %d)", ZZ_NB_WIDE, p->hg.maxdeg);
    if ((flags & ZZB_FLAG_BOOMERANG) && !p->g.grid_m && p->hg.maxdeg > ZZ_NB_WIDE)
        return fail(ZZB_E_ARG, "the Boomerang kernels handle columns of at most %d entries (this matrix has %d)", ZZ_NB_WIDE, p->hg.maxdeg);
    if ((flags & ZZB_FLAG_LOCAL_BOUND) && !p->hg.bnd_eq_tgt)
        return fail(ZZB_E_ARG, "LocalBound builds its bound from the target: create the problem with the sampler matrix equal to the target (bnd_* = NULL) and Z.mu = 0");
    CtxGuard cg;
    zzb_run_s* r = new zzb_run_s();
    r->prob = p; r->d = p->hg.d; r->flags = flags;
    if (const char* e = getenv("ZZB200_DEFAULT_SCHEDULE")) {   // tests: pin the schedule that zzb_run_set("schedule") would otherwise choose
        const int v = atoi(e);
        if (v >= 0 && v <= 1) r->schedule = v;
    }
    const size_t d = (size_t)r->d;
    int32_t st = 0;
#define AL(buf, bytes) if (!st) st = r->buf.alloc(bytes)
    AL(kin, d * sizeof(ZzKin)); AL(flips, d * 2 * ZZ_MAXFLIP * sizeof(double)); AL(priv, d * sizeof(ZzPriv));
    AL(tau, d * 8); AL(kctr, d * 4); AL(spec, d * sizeof(ZzSpec)); AL(viol, d * 24); AL(dstamp, d * 4);
    AL(acc, d * 4); AL(s1, d * 8); AL(s2, d * 8); AL(wl[0], d * 4); AL(wl[1], d * 4); AL(wl[2], d * 4);
    AL(touched, d * 8);   // up to two entries per coordinate (see zz_commit_node)
    AL(ctl, sizeof(ZzDevCtl));
    AL(in_x, d * 8); AL(in_th, d * 8); AL(in_c, d * 8);
    if (flags & ZZB_FLAG_STICKY) { AL(dfth, d * 2 * ZZ_MAXFLIP * sizeof(double)); AL(kappa, d * 8); AL(s3, d * 8); }
    if (flags & ZZB_FLAG_REFRESH) { AL(dfth, d * 2 * ZZ_MAXFLIP * sizeof(double)); AL(rsig, d * 8); AL(rst, d * 16); AL(rspec, d * 16); }
    if (flags & ZZB_FLAG_BOOMERANG) {
        AL(dfth, d * 2 * ZZ_MAXFLIP * sizeof(double)); AL(bsig, d * 8);
        if (!st) st = upload(r->bmu, p->hg.mu.data(), d * 8);
    }
    AL(out_t, d * 8); AL(out_x, d * 8); AL(out_th, d * 8); AL(out_c, d * 8); AL(out_acc, d * 8);
    if (!(flags & ZZB_FLAG_NO_TRACE)) {
        unsigned long long cap = (unsigned long long)std::max<int64_t>(trace_capacity_events, 0);
        const unsigned long long need = (unsigned long long)d * ZZ_MAXFLIP + 2ULL;
        if (cap == 0) {
            size_t fr = 0, tot = 0;
            if (!st && g_drv.p_cuMemGetInfo(&fr, &tot) != CUDA_SUCCESS) fr = (size_t)1 << 30;
            cap = std::max<unsigned long long>(need * 4, 1ULL << 22);
            cap = std::min<unsigned long long>(cap, std::max<unsigned long long>(need, (fr / 4) / sizeof(ZzEvent)));
        }
        if (cap < need) cap = need;
        r->trace_cap = cap;
        AL(trace, (size_t)cap * sizeof(ZzEvent));
    }
#undef AL
    if (st) { delete r; return st == ZZB_E_CUDA ? ZZB_E_NOMEM : st; }
    r->kind = p->g.grid_m ? 0 : 1;
    r->grid = G.sm_count * G.blocks_per_sm[r->kidx()];
    r->shard = r->d; r->hi = r->d;
    *out = r;
    return ZZB_OK;
}

// Tiles of the asynchronous relaxation: `per` coordinates per CTA (a multiple of 32; sharded runs: derived from the nominal
// shard size so that every rank computes the same value), two bit arrays per tile in dynamic shared memory, one inbox per CTA
// for marks that cross a tile boundary.
static int32_t setup_tiles(zzb_run_s* r)
{
    const long long span = r->nranks > 1 ? r->shard : r->d;
    const long long per = (((span + r->grid - 1) / r->grid) + 31) & ~31LL;
    r->tile_per = (int)per;
    r->flag_words = (unsigned int)(per / 32);
    // one sub-box per (receiving tile, sending rank); the slot counters [attempt mod 3][destination rank][destination tile] are the
    // SENDER's and live in its own memory
    const unsigned int cap = (unsigned int)std::max<long long>(1024, 2 * per);
    const size_t nr = (size_t)r->nranks;
    if (r->inbox_grid != r->grid || r->inbox_cap != cap || r->inbox_nr != r->nranks) {
        int32_t st = r->inbox.alloc((size_t)r->grid * nr * cap * 8);
        if (!st) st = r->inbox_cnt.alloc((size_t)3 * nr * r->grid * 4);
        if (st) return st == ZZB_E_CUDA ? ZZB_E_NOMEM : st;
        CU(cuMemsetD8(r->inbox.p, 0xff, (size_t)r->grid * nr * cap * 8));   // attempt tag 0xffffffff: "not written"
        CU(cuMemsetD8(r->inbox_cnt.p, 0, (size_t)3 * nr * r->grid * 4));
        r->inbox_grid = r->grid; r->inbox_cap = cap; r->inbox_nr = r->nranks;
    }
    return ZZB_OK;
}

static void fill_params(zzb_run_s* r)
{
    ZzParams& P = r->P;
    memset(&P, 0, sizeof P);
    P.g = r->prob->g;
    if (r->prob->logit) P.lg = r->prob->lg;
    P.v.d = r->d; P.v.kin = r->kin.as<ZzKin>(); P.v.flips = r->flips.as<double>(); P.v.priv = r->priv.as<ZzPriv>();
    P.v.tau = r->tau.as<double>(); P.v.kctr = r->kctr.as<uint32_t>();
    P.dptr = r->prob->dptr.as<int32_t>(); P.didx = r->prob->didx.as<int32_t>();
    P.spec = r->spec.as<ZzSpec>(); P.viol_info = r->viol.as<double>(); P.dstamp = r->dstamp.as<unsigned int>();
    P.acc = r->acc.as<unsigned int>(); P.s1 = r->s1.as<double>(); P.s2 = r->s2.as<double>();
    for (int k = 0; k < 3; ++k) P.wl[k] = r->wl[k].as<int32_t>();
    P.touched[0] = r->touched.as<int32_t>();
    P.trace = r->trace.as<ZzEvent>(); P.trace_cap = r->trace_cap;
    P.ctl = r->ctl.as<ZzDevCtl>();
    P.inbox = r->inbox.as<unsigned long long>(); P.inbox_cnt = r->inbox_cnt.as<unsigned int>();
    P.inbox_cap = r->inbox_cap; P.flag_words = r->flag_words; P.tile_per = r->tile_per;
    P.record_trace = (r->flags & ZZB_FLAG_NO_TRACE) ? 0 : 1;
    P.trace_map = r->have_map ? r->trace_map.as<int32_t>() : nullptr;
    P.s3 = (r->flags & ZZB_FLAG_STICKY) ? r->s3.as<double>() : nullptr;
    P.grid = r->grid_n ? r->gridbuf.as<double>() : nullptr; P.grid_dt = r->grid_dt; P.grid_n = r->grid_n;
    P.v.local_bound = (r->flags & ZZB_FLAG_LOCAL_BOUND) ? 1 : 0;
    P.v.sticky = (r->flags & ZZB_FLAG_STICKY) ? (1 | ((r->flags & ZZB_FLAG_STICKY_REVERSIBLE) ? ZZ_STICKY_REVERSIBLE : 0) |
                                                 ((r->flags & ZZB_FLAG_STICKY_STRONG_UB) ? ZZ_STICKY_STRONG_UB : 0) |
                                                 ((r->flags & ZZB_FLAG_STICKY_ZZ) ? ZZ_STICKY_ZZ : 0)) : 0;
    P.v.boom = (r->flags & ZZB_FLAG_BOOMERANG) ? 1 : 0;
    P.v.refresh = (r->flags & ZZB_FLAG_REFRESH) ? 1 : 0;
    P.v.fth = (P.v.sticky || P.v.boom || P.v.refresh) ? r->dfth.as<double>() : nullptr;
    if (P.v.refresh) {
        P.v.rsig = r->rsig.as<double>(); P.v.rlam1 = r->rlambdaref / (double)r->d;
        P.v.rst = r->rst.as<double>(); P.v.rspec = r->rspec.as<double>();
    }
    if (P.v.boom) {
        P.v.bmu = r->bmu.as<double>(); P.v.bsig = r->bsig.as<double>();
        P.v.bref_rate = r->lambdaref / (double)r->d; P.v.brho = r->rho; P.v.brhobar = sqrt(1 - r->rho * r->rho);
    }
    P.v.kappa = P.v.sticky ? r->kappa.as<double>() : nullptr;
    P.st.c = r->strong_c; P.st.kappa = r->kappa0; P.st.rule = r->strong_rule; P.st.pad = 0;
    P.v.nranks = r->nranks; P.v.rank = r->rank; P.v.shard = r->shard; P.v.lo = r->lo; P.v.hi = r->hi;
    if (r->nranks > 1) {
        for (int q = 0; q < r->nranks; ++q) {
            const bool me = (q == r->rank);
            P.v.kin_peer[q] = me ? P.v.kin : reinterpret_cast<ZzKin*>(r->peer[q][0]);
            P.v.flips_peer[q] = me ? P.v.flips : reinterpret_cast<double*>(r->peer[q][1]);
            P.v.fth_peer[q] = me ? P.v.fth : reinterpret_cast<double*>(r->peer[q][5]);
            P.inbox_peer[q] = me ? P.inbox : reinterpret_cast<unsigned long long*>(r->peer[q][3]);
            P.inbox_cnt_peer[q] = me ? P.inbox_cnt : reinterpret_cast<unsigned int*>(r->peer[q][4]);
            P.dstamp_peer[q] = me ? P.dstamp : reinterpret_cast<unsigned int*>(r->peer[q][2]);
            for (int k = 0; k < 3; ++k) P.wl_peer[k][q] = P.wl[k];   // (work lists are private to their rank in the asynchronous schedule)
            P.touched_peer[q] = me ? P.touched[0] : reinterpret_cast<int32_t*>(r->peer[q][6]);
            P.ctl_peer[q] = me ? P.ctl : reinterpret_cast<ZzDevCtl*>(r->peer[q][7]);
        }
    }
}

static CUdeviceptr shared_buf(zzb_run_s* r, int k)
{
    switch (k) {
    case 0: return r->kin.p; case 1: return r->flips.p; case 2: return r->dstamp.p;
    case 3: return r->inbox.p; case 4: return r->inbox_cnt.p; case 5: return r->dfth.p ? r->dfth.p : r->wl[2].p;   // (velocity lists of the sticky / Boomerang samplers)
    case 6: return r->touched.p; default: return r->ctl.p;
    }
}

// ---- coordinate sharding across the GPUs of one node (one process per GPU) ----------------
// Rank `rank` of `nranks` owns the global coordinates [rank*shard, (rank+1)*shard); on the lattice the shard is a
// whole number of lattice columns.  Call before zzb_run_upload.
int32_t zzb_run_shard(zzb_run_t r, int32_t rank, int32_t nranks)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (nranks < 1 || nranks > ZZ_MAXRANKS || rank < 0 || rank >= nranks) return fail(ZZB_E_ARG, "bad rank %d of %d", rank, nranks);
    if (nranks > 1 && r->strong) return fail(ZZB_E_ARG, "the strong-bound sticky samplers are not sharded");
    if (nranks > 1 && r->prob->logit) return fail(ZZB_E_ARG, "the logistic target is not sharded (its dependency graph is complete: replicas only)");
    const int64_t d = r->d;
    int64_t shard = (d + nranks - 1) / nranks;
    const int64_t m = r->prob->g.grid_m;
    if (m) shard = ((shard + m - 1) / m) * m;
    r->rank = rank; r->nranks = nranks; r->shard = (int)shard;
    r->lo = (int)std::min<int64_t>(d, shard * rank);
    r->hi = (int)std::min<int64_t>(d, shard * (rank + 1));
    r->grid = G.sm_count * G.blocks_per_sm[r->kidx()];
    CtxGuard cg;
    return setup_tiles(r);   // (the inboxes are exported to the peers: allocate them now)
}

// 8 IPC handles (kin, flips, dstamp, inbox, inbox counters, velocity lists [or an unused work list], touched list, control block) of this rank.
int32_t zzb_run_ipc_export(zzb_run_t r, void* buf, int64_t cap, int64_t* len)
{
    if (!r || !buf || !len) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    const int64_t need = 8 * (int64_t)sizeof(CUipcMemHandle);
    if (cap < need) return fail(ZZB_E_ARG, "buffer too small (%lld < %lld)", (long long)cap, (long long)need);
    CtxGuard cg;
    CUipcMemHandle* h = reinterpret_cast<CUipcMemHandle*>(buf);
    for (int k = 0; k < 8; ++k) CU(cuIpcGetMemHandle(&h[k], shared_buf(r, k)));
    *len = need;
    return ZZB_OK;
}

int32_t zzb_run_ipc_import(zzb_run_t r, int32_t peer_rank, const void* buf, int64_t len)
{
    if (!r || !buf) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (peer_rank < 0 || peer_rank >= r->nranks || peer_rank == r->rank) return fail(ZZB_E_ARG, "bad peer rank %d", peer_rank);
    if (len < 8 * (int64_t)sizeof(CUipcMemHandle)) return fail(ZZB_E_ARG, "short handle buffer");
    CtxGuard cg;
    const CUipcMemHandle* h = reinterpret_cast<const CUipcMemHandle*>(buf);
    for (int k = 0; k < 8; ++k) CU(cuIpcOpenMemHandle(&r->peer[peer_rank][k], h[k], CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS));
    r->peer_open[peer_rank] = true;
    return ZZB_OK;
}

int32_t zzb_run_range(zzb_run_t r, int64_t* lo, int64_t* hi)
{
    if (!r || !lo || !hi) return fail(ZZB_E_ARG, "null argument");
    *lo = r->lo; *hi = r->hi;
    return ZZB_OK;
}

int32_t zzb_run_set(zzb_run_t r, const char* key, double value)
{
    if (!r || !key) return fail(ZZB_E_ARG, "null argument");
    if (!strcmp(key, "delta0")) r->delta0 = value;
    else if (!strcmp(key, "target_frac")) { r->target_frac = value; r->window_policy_set = true; }
    else if (!strcmp(key, "target_flip_frac")) { r->target_flip_frac = value; r->window_policy_set = true; }
    else if (!strcmp(key, "tag_limit")) r->tag_limit = (unsigned int)value;
    else if (!strcmp(key, "max_windows")) r->max_windows = (unsigned int)value;
    else if (!strcmp(key, "host_sort")) r->host_sort_only = value != 0.0;   // order the trace on the host (A/B of the device sort)
    else if (!strcmp(key, "eval_threads")) r->eval_threads = (unsigned int)value;
    else if (!strcmp(key, "seq_warps")) r->seq_warps = (int)value;   // sequential chains: 1 (plain loop), 2, 4 or 8 warps per chain; 0 = automatic
    else if (!strcmp(key, "schedule")) {   // 0 / 1: windowed relaxation (pass-synchronous / asynchronous); 2: sequential chains; -1: automatic
        const int v = (int)value;
        if (v < -1 || v > 2) return fail(ZZB_E_ARG, "schedule must be -1, 0, 1 or 2");
        if (r->executed) return fail(ZZB_E_ARG, "the schedule of a run is chosen before its first zzb_run_execute (zzb_run_reset starts over)");
        if (v == 2 && !r->seq_capable())
            return fail(ZZB_E_ARG, "the sequential-chain schedule is not available for this run (%s)",
                        r->prob->hs.ok ? "plain ZigZag on one GPU only, no device-side discretisation" : r->prob->hs.why.c_str());
        r->schedule = v; r->grid = G.sm_count * G.blocks_per_sm[r->kidx()];
    }
    // switch a sticky run to the strong-bound sampler of src/sparsestickyzz.jl: scalar bound constant c, rule (0 sticky, 1 reversible);
    // kappa[0] of zzb_run_upload_kappa is the thaw rate; coordinates with x0 == 0 start frozen.  Before zzb_run_upload.
    else if (!strcmp(key, "strong_c")) {
        if (!(r->flags & ZZB_FLAG_STICKY) || !(value > 0.0)) return fail(ZZB_E_ARG, "strong_c needs a sticky run and c > 0");
        if (r->prob->hg.maxdeg > ZZ_NB) return fail(ZZB_E_ARG, "the strong-bound kernel handles columns of at most %d entries (this matrix has %d)", ZZ_NB, r->prob->hg.maxdeg);
        r->strong = true; r->strong_c = value; r->grid = G.sm_count * G.blocks_per_sm[r->kidx()];
    }
    else if (!strcmp(key, "strong_rule")) r->strong_rule = (int)value;
    else if (!strcmp(key, "grid")) r->grid = std::max(1, std::min((int)value, G.sm_count * G.blocks_per_sm[r->kidx()]));
    else return fail(ZZB_E_ARG, "unknown tuning key %s", key);
    return ZZB_OK;
}

// Thaw rates kappa_i of the sticky sampler (ss_fact.jl:96); call before zzb_run_upload.
int32_t zzb_run_upload_kappa(zzb_run_t r, const double* kappa)
{
    if (!r || !kappa) return fail(ZZB_E_ARG, "null argument");
    if (!(r->flags & ZZB_FLAG_STICKY)) return fail(ZZB_E_ARG, "run was not created with ZZB_FLAG_STICKY");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    for (int32_t j = 0; j < r->d; ++j) if (!(kappa[j] > 0.0)) return fail(ZZB_E_ARG, "kappa[%d] must be positive", j + 1);
    CtxGuard cg;
    CU(cuMemcpyHtoD(r->kappa.p, kappa, (size_t)r->d * 8));
    r->have_kappa = true;
    r->kappa0 = kappa[0];
    return ZZB_OK;
}

// FactBoomerang parameters (types.jl:62-79): sigma scales refreshed velocities, lambdaref = total refreshment rate (> 0:
// hasrefresh(::FactBoomerang) = true, fact_samplers.jl:18), rho = autoregression of the refreshment.  Before zzb_run_upload.
int32_t zzb_run_upload_boomerang(zzb_run_t r, const double* sigma, double lambdaref, double rho)
{
    if (!r || !sigma) return fail(ZZB_E_ARG, "null argument");
    if (!(r->flags & ZZB_FLAG_BOOMERANG)) return fail(ZZB_E_ARG, "run was not created with ZZB_FLAG_BOOMERANG");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!(lambdaref > 0.0)) return fail(ZZB_E_ARG, "FactBoomerang needs a refreshment rate lambdaref > 0");
    if (!(rho > -1.0 && rho < 1.0)) return fail(ZZB_E_ARG, "rho must lie in (-1, 1)");
    CtxGuard cg;
    CU(cuMemcpyHtoD(r->bsig.p, sigma, (size_t)r->d * 8));
    r->lambdaref = lambdaref; r->rho = rho; r->have_boom = true;
    return ZZB_OK;
}

// ZigZag velocity refreshments (Z = ZigZag(Gamma, mu, sigma; lambdaref > 0): hasrefresh(Z), src/fact_samplers.jl:19, src/sfact.jl:78-114):
// sigma scales the refreshed velocities (theta_i <- sigma_i * (+-1)), lambdaref = total refreshment rate.  Before zzb_run_upload.
int32_t zzb_run_upload_refresh(zzb_run_t r, const double* sigma, double lambdaref)
{
    if (!r || !sigma) return fail(ZZB_E_ARG, "null argument");
    if (!(r->flags & ZZB_FLAG_REFRESH)) return fail(ZZB_E_ARG, "run was not created with ZZB_FLAG_REFRESH");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!(lambdaref > 0.0)) return fail(ZZB_E_ARG, "refreshments need lambdaref > 0");
    for (int32_t j = 0; j < r->d; ++j) if (!(sigma[j] > 0.0)) return fail(ZZB_E_ARG, "sigma[%d] must be positive", j + 1);
    CtxGuard cg;
    CU(cuMemcpyHtoD(r->rsig.p, sigma, (size_t)r->d * 8));
    r->rlambdaref = lambdaref; r->have_refresh = true;
    return ZZB_OK;
}

// (Re)initialise the device state from the inputs already resident in HBM: per-coordinate records, initial
// bounds and first proposal times (sfact.jl:167-187).  No host<->device traffic except the 200-byte control block.
int32_t zzb_run_reset(zzb_run_t r)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!r->have_inputs) return fail(ZZB_E_ARG, "zzb_run_upload must precede zzb_run_reset");
    if ((r->flags & ZZB_FLAG_STICKY) && !r->have_kappa) return fail(ZZB_E_ARG, "zzb_run_upload_kappa must precede zzb_run_upload for a sticky run");
    if ((r->flags & ZZB_FLAG_BOOMERANG) && !r->have_boom) return fail(ZZB_E_ARG, "zzb_run_upload_boomerang must precede zzb_run_upload for a Boomerang run");
    if ((r->flags & ZZB_FLAG_REFRESH) && !r->have_refresh) return fail(ZZB_E_ARG, "zzb_run_upload_refresh must precede zzb_run_upload for a run with ZZB_FLAG_REFRESH");
    if ((r->flags & ZZB_FLAG_REFRESH) && r->nranks > 1) return fail(ZZB_E_ARG, "ZigZag refreshments are not sharded yet");
    if (r->strong && r->nranks > 1) return fail(ZZB_E_ARG, "the strong-bound sticky samplers are not sharded");
    if (r->strong && r->adapt) return fail(ZZB_E_ARG, "adapt is not supported by the strong-bound sticky samplers on the device path");
    for (int q = 0; q < r->nranks; ++q)
        if (q != r->rank && !r->peer_open[q]) return fail(ZZB_E_ARG, "peer %d of a sharded run has not been imported", q);
    fill_params(r);
    r->P.setup_lo = r->setup_lo; r->P.setup_hi = r->setup_hi ? r->setup_hi : r->d;
    r->P.init_lo = (r->nranks > 1 && r->setup_hi && (r->setup_lo > 0 || r->setup_hi < r->d)) ? r->lo : 0;
    r->P.init_hi = (r->nranks > 1 && r->setup_hi && (r->setup_lo > 0 || r->setup_hi < r->d)) ? r->hi : r->d;
    r->P.v.seed0 = r->seed[0]; r->P.v.seed1 = r->seed[1]; r->P.v.adapt = r->adapt; r->P.v.factor = r->factor; r->P.t0 = r->t0;
    CtxGuard cg;
    ZzParams& P = r->P;
    ZzDevCtl hc; memset(&hc, 0, sizeof hc);
    hc.f0_key = ~0ULL; for (int k = 0; k < 3; ++k) hc.smin_key[k] = ~0ULL;
    hc.wattempt = r->wat_next;
    CU(cuMemcpyHtoDAsync(r->ctl.p, &hc, sizeof hc, G.stream));
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)r->d + ZZ_BLOCK - 1) / ZZ_BLOCK, (size_t)G.sm_count * 8);
    if (r->grid_n) CU(cuMemsetD8Async(r->gridbuf.p, 0xff, (size_t)r->grid_n * (size_t)r->d * 8, G.stream));   // all-ones = NaN
    CUdeviceptr px = r->in_x.p, pth = r->in_th.p, pc = r->in_c.p;
    void* a1[] = { &P, &px, &pth, &pc };
    CU(cuLaunchKernel(G.f_setup, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a1, nullptr));
    void* a2[] = { &P };
    if (r->strong) CU(cuLaunchKernel(G.f_init_strong, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a2, nullptr));
    else
    CU(cuLaunchKernel((r->flags & ZZB_FLAG_BOOMERANG) ? G.f_init_boom : G.f_init, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a2, nullptr));
    CU(cuStreamSynchronize(G.stream));
    r->launches += 2;
    r->uploaded = true; r->executed = false; r->fetched = false;
    r->events.clear();
    r->dev_sorted = false; r->n_sorted = 0;
    return ZZB_OK;
}

int32_t zzb_run_upload(zzb_run_t r, double t0, const double* x0, const double* theta0, const double* c,
                       const uint64_t* seed, int32_t adapt, double factor)
{
    if (!r || !x0 || !theta0 || !c || !seed) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    {
        CtxGuard cg;
        size_t nb = (size_t)r->d * 8;
        r->seed[0] = seed[0]; r->seed[1] = seed[1]; r->adapt = adapt; r->factor = factor;
        r->t0 = t0;
        // sharded lattice run: this rank sets up its slab and one lattice column on either side (the replicas it reads); the
        // host-to-device traffic of the node is then ~ d, not N d
        size_t so = 0;
        r->setup_lo = 0; r->setup_hi = r->d;
        if (r->nranks > 1 && r->prob->g.grid_m && !r->strong) {
            r->setup_lo = std::max(0, r->lo - r->prob->g.grid_m); r->setup_hi = std::min(r->d, r->hi + r->prob->g.grid_m);
            so = (size_t)r->setup_lo * 8; nb = (size_t)(r->setup_hi - r->setup_lo) * 8;
        }
        CU(cuMemcpyHtoDAsync(r->in_x.p + so, x0 + so / 8, nb, G.stream));
        if (r->strong && r->strong_rule != 2) {   // sparsestickystate (sparsestickyzz.jl:10-12): x0 == 0 starts frozen = velocity 0 in its record (rule 2: zz_setup_kernel keeps the velocity)
            std::vector<double> th(theta0, theta0 + r->d);
            for (int32_t j = 0; j < r->d; ++j) if (x0[j] == 0.0) th[j] = 0.0;
            CU(cuMemcpyHtoD(r->in_th.p, th.data(), nb));
        } else
        CU(cuMemcpyHtoDAsync(r->in_th.p + so, theta0 + so / 8, nb, G.stream));
        CU(cuMemcpyHtoDAsync(r->in_c.p + so, c + so / 8, nb, G.stream));
        r->have_inputs = true;
    }
    return zzb_run_reset(r);
}

struct ZzTsortHost {   // same layout as ZzTsort in zz_kernels.cu
    CUdeviceptr in, out; unsigned long long n; double tmin, scale; unsigned int nb; unsigned int mode;
    CUdeviceptr cnt, base, fill, ovf;
};

// Order the n records of the device trace buffer by (time, coordinate) on the device (markers dropped); true in *ok when the
// sorted events are in r->trace_sorted, false when a time bucket overflowed (caller falls back to the host sort).
static int32_t device_sort_trace(zzb_run_s* r, unsigned long long n, double tmin, double tmax, bool* ok)
{
    *ok = false;
    const unsigned int nb = (unsigned int)std::max<unsigned long long>(1, n / 128);
    int32_t st = ZZB_OK;
    if (r->trace_sorted.n < (size_t)n * sizeof(zzb_event)) st = r->trace_sorted.alloc((size_t)r->trace_cap * sizeof(zzb_event));
    const size_t wbytes = ((size_t)3 * nb + 2) * 4;
    if (!st && r->ts_work.n < wbytes) st = r->ts_work.alloc(wbytes + 4096);
    if (st) return ZZB_OK;   // no memory for the second buffer: the host sorts
    CU(cuMemsetD8Async(r->ts_work.p, 0, wbytes, G.stream));
    ZzTsortHost Q;
    Q.in = r->trace.p; Q.out = r->trace_sorted.p; Q.n = n; Q.tmin = tmin;
    Q.scale = (tmax > tmin) ? (double)nb / (tmax - tmin) : 0.0; Q.nb = nb; Q.mode = 0;
    Q.cnt = r->ts_work.p; Q.base = r->ts_work.p + (size_t)nb * 4; Q.fill = r->ts_work.p + ((size_t)2 * nb + 1) * 4; Q.ovf = r->ts_work.p + ((size_t)3 * nb + 1) * 4;
    void* a[] = { &Q };
    const unsigned grid = (unsigned)std::min<unsigned long long>((n + ZZ_BLOCK - 1) / ZZ_BLOCK, (unsigned long long)G.sm_count * 16);
    CU(cuLaunchKernel(G.f_ts_hist, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    CU(cuLaunchKernel(G.f_ts_scan, 1, 1, 1, 1024, 1, 1, 0, G.stream, a, nullptr));
    CU(cuLaunchKernel(G.f_ts_scatter, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    CU(cuLaunchKernel(G.f_ts_sort, std::min<unsigned>(nb, (unsigned)G.sm_count * 8), 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    r->launches += 4;
    unsigned int tail[2] = { 0, 0 };   // base[nb] is followed by fill[]; ovf is read separately
    CU(cuMemcpyDtoHAsync(&tail[0], Q.base + (size_t)nb * 4, 4, G.stream));
    CU(cuMemcpyDtoHAsync(&tail[1], Q.ovf, 4, G.stream));
    CU(cuStreamSynchronize(G.stream));
    if (tail[1]) return ZZB_OK;
    r->n_sorted = tail[0];
    *ok = true;
    return ZZB_OK;
}

// sort every window segment (records between two i == 0 markers) by (time, coordinate) and drop the markers
static void absorb_trace_chunk(zzb_run_s* r, std::vector<zzb_event>& chunk)
{
    size_t seg = 0;
    std::vector<std::pair<size_t, size_t>> segs;
    for (size_t k = 0; k < chunk.size(); ++k)
        if (chunk[k].i == 0) { if (k > seg) segs.emplace_back(seg, k); seg = k + 1; }
    if (seg < chunk.size()) segs.emplace_back(seg, chunk.size());
    auto cmp = [](const zzb_event& a, const zzb_event& b) { return a.t < b.t || (a.t == b.t && a.i < b.i); };
    unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (chunk.size() < (1u << 16)) nthr = 1;
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nthr; ++w)
        th.emplace_back([&, w]() { for (size_t s = w; s < segs.size(); s += nthr) std::sort(chunk.begin() + segs[s].first, chunk.begin() + segs[s].second, cmp); });
    for (auto& t : th) t.join();
    size_t total = 0;
    for (auto& s : segs) total += s.second - s.first;
    r->events.reserve(r->events.size() + total);
    for (auto& s : segs) r->events.insert(r->events.end(), chunk.begin() + s.first, chunk.begin() + s.second);
}

// Sequential-chain schedule (zz_seq.cuh): one warp per connected component runs the reference's event loop as written.  Phase 0
// processes every item before T, phase 1 finds the first accepted flip at or after T over all chains (the loop of
// src/sfact.jl:199 ends there), phase 2 processes every item up to that time.  A full trace buffer interrupts a phase: the
// host drains it and relaunches (the chains resume from their saved state).
static int32_t execute_seq(zzb_run_s* r, double T, float* device_ms)
{
    ZzParams& P = r->P;
    zzb_problem_s* pb = r->prob;
    if (!r->seq_capable())
        return fail(ZZB_E_ARG, "the sequential-chain schedule is not available for this run (%s)",
                    pb->hs.ok ? "plain ZigZag on one GPU only, no device-side discretisation" : pb->hs.why.c_str());
    if (!(T < (double)INFINITY)) return fail(ZZB_E_ARG, "the sequential-chain schedule needs a finite end time");
    if (r->executed && !(r->hc.ctl.F < T)) return ZZB_OK;   // `while t' < T` (sfact.jl:199): the last event is already at or after T
    ZzSeq Q = pb->sq;
    // warps per chain (speculative thinning, zz_seq.cuh): 4 when the chains are few (latency of the single chain), 2 when many
    // chains share the SMs (registers: six chains of two warps fit), unless zzb_run_set("seq_warps") says otherwise
    auto smem_for = [&](int w) { return (Q.orig ? 80u : 76u) * (unsigned)Q.ncmax + 8u + (unsigned)w * 8u * (72u + 2u * ((unsigned)Q.colmax + 8u)); };   // state + scratch per warp
    CUfunction f = pb->logit ? G.f_seq_logit : G.f_seq;
    CU(cuFuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)std::min(std::max(smem_for(8), 48u * 1024u), 227u * 1024u)));
    int nw = r->seq_warps;
    if (nw != 1 && nw != 2 && nw != 4 && nw != 8) {
        // automatic: eight warps per chain while every chain has an SM of its own (latency of the single chain), else four --
        // measured (profiles/r02_logit_seq.log): with many chains four warps each give the best throughput even in several
        // waves (888 chains: 1.4e7 events/s with 4 warps in two waves, 1.2e7 with 1 warp each and all resident)
        nw = (Q.ncomp <= G.sm_count) ? 8 : 4;
    }
    while (nw > 1 && P.record_trace && r->trace_cap < (unsigned long long)ZZ_SEQ_RES_HOST * (unsigned long long)Q.ncomp * (unsigned long long)nw)
        nw >>= 1;   // (every warp of every chain may hold one reservation of trace records)
    const unsigned dyn = smem_for(nw);
    if (dyn > 220u * 1024u) return fail(ZZB_E_ARG, "the sequential-chain schedule needs %u bytes of shared memory per chain", dyn);
    if (P.record_trace && r->trace_cap < (unsigned long long)ZZ_SEQ_RES_HOST * (unsigned long long)Q.ncomp)
        return fail(ZZB_E_TRACE, "trace buffer (%llu records) too small for %d chains (%u records reserved at a time)", r->trace_cap, Q.ncomp, ZZ_SEQ_RES_HOST);
    if (r->dev_sorted && r->n_sorted) {   // events of an earlier execute that are still on the device (sorted) join the host vector first
        const size_t at = r->events.size();
        r->events.resize(at + (size_t)r->n_sorted);
        CU(cuMemcpyDtoH(r->events.data() + at, r->trace_sorted.p, (size_t)r->n_sorted * sizeof(zzb_event)));
    }
    r->dev_sorted = false; r->n_sorted = 0;
    const double t_front0 = r->executed ? r->hc.ctl.F : r->t0;
    const size_t ev_start = r->events.size();
    bool drained = !r->events.empty() || r->host_sort_only;
    float total_ms = 0.f;
    ZzDevCtl& hc = r->hc;
    auto launch = [&](int phase) -> int32_t {
        Q.phase = phase;
        void* args[] = { &P, &Q };
        CU(cuEventRecord(G.ev0, G.stream));
        CU(cuLaunchKernel(f, (unsigned)Q.ncomp, 1, 1, 32u * (unsigned)nw, 1, 1, dyn, G.stream, args, nullptr));
        CU(cuEventRecord(G.ev1, G.stream));
        CU(cuStreamSynchronize(G.stream));
        r->launches++;
        float ms = 0.f;
        CU(cuEventElapsedTime(&ms, G.ev0, G.ev1));
        total_ms += ms;
        CU(cuMemcpyDtoH(&hc, r->ctl.p, sizeof(ZzDevCtl)));
        return ZZB_OK;
    };
    auto drain = [&]() -> int32_t {   // unordered records of the buffer (markers dropped) join the host vector
        const unsigned long long n = std::min<unsigned long long>(hc.trace_len, r->trace_cap);
        std::vector<zzb_event> chunk((size_t)n);
        if (n) CU(cuMemcpyDtoH(chunk.data(), r->trace.p, (size_t)n * sizeof(zzb_event)));
        for (const zzb_event& e : chunk) if (e.i != 0) r->events.push_back(e);
        hc.trace_len = 0; hc.need_drain = 0;
        CU(cuMemcpyHtoD(r->ctl.p, &hc, sizeof hc));
        drained = true;
        return ZZB_OK;
    };
    double t_end = T;
    bool viol = false;
    for (int phase = 0; phase < 3 && !viol; ++phase) {
        if (phase == 1) {
            CU(cuMemcpyDtoH(&hc, r->ctl.p, sizeof(ZzDevCtl)));
            hc.smin_key[0] = ~0ULL;
            CU(cuMemcpyHtoD(r->ctl.p, &hc, sizeof hc));
        }
        for (;;) {
            int32_t st = launch(phase);
            if (st) return st;
            if (hc.viol) { viol = true; break; }
            if (!hc.need_drain) break;
            if (!P.record_trace) return fail(ZZB_E_INTERNAL, "sequential schedule: drain requested without a trace");
            st = drain();
            if (st) return st;
        }
        if (phase == 1) {
            if (hc.smin_key[0] == ~0ULL) break;   // no chain will ever flip again: everything before T has been processed
            t_end = zz_unkey(hc.smin_key[0]);
        }
    }
    hc.ctl.F = viol ? std::max(t_front0, std::min(T, hc.viol_t)) : t_end;
    hc.ctl.phase = ZZ_PH_DONE;
    if (P.record_trace) {
        const unsigned long long n = std::min<unsigned long long>(hc.trace_len, r->trace_cap);
        bool ok = false;
        if (!drained && !viol && n) {
            int32_t st = device_sort_trace(r, n, t_front0, hc.ctl.F, &ok);
            if (st) return st;
            if (ok) { r->dev_sorted = true; hc.trace_len = 0; hc.need_drain = 0; }
        }
        if (!ok) {
            int32_t st = drain();
            if (st) return st;
            auto cmp = [](const zzb_event& a, const zzb_event& b) { return a.t < b.t || (a.t == b.t && a.i < b.i); };
            std::sort(r->events.begin() + (std::ptrdiff_t)ev_start, r->events.end(), cmp);
        }
    }
    CU(cuMemcpyHtoD(r->ctl.p, &hc, sizeof hc));
    if (device_ms) *device_ms = total_ms;
    r->executed = true; r->fetched = false;
    if (viol)
        return fail(ZZB_E_BOUND, "Tuning parameter `c` too small. (coordinate %d, t = %.17g, l = %.17g, lb = %.17g)",
                    hc.viol_i, hc.viol_t, hc.viol_l, hc.viol_lb);
    return ZZB_OK;
}

int32_t zzb_run_execute(zzb_run_t r, double T, float* device_ms)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!r->uploaded) return fail(ZZB_E_ARG, "zzb_run_upload must precede zzb_run_execute");
    CtxGuard cg;
    ZzParams& P = r->P;
    r->T = T;
    P.T = T;
    double span = T - r->t0;
    P.delta0 = r->delta0 > 0 ? r->delta0 : std::max(1e-12, 1e-2 * std::min(1.0, span > 0 ? span : 1.0));
    // window length policy: proposals per window (and a cap on the accepted flips) as fractions of d.  Defaults measured on a B200:
    // 0.25 / 0.045 at d = 10^6 (longer windows lengthen the wait for the slowest tile); small problems -- a tile of a few dozen
    // coordinates per SM -- want windows four times longer (d = 10^4: 9.3e6 -> 1.27e7 switches/s)
    double tfrac = r->target_frac, ffrac = r->target_flip_frac;
    if (!r->window_policy_set && r->d <= 30000) { tfrac = 1.0; ffrac = 0.18; }
    P.target = std::max(tfrac * (double)r->d, 4.0);              // proposals per window
    P.target_flips = std::max(ffrac * (double)r->d, 2.0);        // cap on accepted flips per window
    P.tag_limit = r->tag_limit; P.max_windows = r->max_windows;
    float total_ms = 0.f;
    if (device_ms) *device_ms = 0.f;
    if (!(r->t0 < T)) { r->executed = true; return ZZB_OK; }  // `while t' < T` never entered (sfact.jl:199)
    if (r->sched() == 2) return execute_seq(r, T, device_ms);
    unsigned dyn_smem = 0;
    if (ZZ_KERN_ASYNC(r->kidx())) {
        if (r->nranks <= 1) { int32_t st = setup_tiles(r); if (st) return st; }   // (sharded: done by zzb_run_shard, before the IPC export)
        if (!r->tile_per) return fail(ZZB_E_ARG, "zzb_run_shard must precede zzb_run_execute for a sharded run");
        dyn_smem = 2u * 4u * r->flag_words;
        if (dyn_smem > 200u * 1024u) return fail(ZZB_E_ARG, "d = %d is too large for %d tiles (bit arrays of %u bytes per CTA)", r->d, r->grid, dyn_smem);
        if (dyn_smem > 32u * 1024u)
            CU(cuFuncSetAttribute(G.f_run[r->kidx()], CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem));
        P.inbox = r->inbox.as<unsigned long long>(); P.inbox_cnt = r->inbox_cnt.as<unsigned int>(); P.inbox_cap = r->inbox_cap;
        P.flag_words = r->flag_words; P.tile_per = r->tile_per; P.eval_threads = r->eval_threads;
        for (int q = 0; q < r->nranks && r->nranks > 1; ++q) {
            const bool me = (q == r->rank);
            P.inbox_peer[q] = me ? P.inbox : reinterpret_cast<unsigned long long*>(r->peer[q][3]);
            P.inbox_cnt_peer[q] = me ? P.inbox_cnt : reinterpret_cast<unsigned int*>(r->peer[q][4]);
        }
        if (const char* dw = getenv("ZZB200_DBG_WINDOW")) {   // development: log the rounds of one window per CTA (tools/window_trace.py)
            const size_t nb = (size_t)r->grid * ZZ_DBG_REC * 4 * 8 + 4096 * 16 * 8;
            int32_t st = r->dbgbuf.alloc(nb);
            if (st) return st;
            CU(cuMemsetD8Async(r->dbgbuf.p, 0, nb, G.stream));
            P.dbgbuf = r->dbgbuf.as<unsigned long long>(); P.dbg_window = (unsigned int)atoi(dw);
        } else P.dbgbuf = nullptr;
    }
    // events of an earlier execute that are still on the device (sorted) join the host vector first
    if (r->dev_sorted && r->n_sorted) {
        const size_t at = r->events.size();
        r->events.resize(at + (size_t)r->n_sorted);
        CU(cuMemcpyDtoH(r->events.data() + at, r->trace_sorted.p, (size_t)r->n_sorted * sizeof(zzb_event)));
    }
    r->dev_sorted = false; r->n_sorted = 0;
    const double t_front0 = r->executed ? r->hc.ctl.F : r->t0;   // every event of this call is at or after the current frontier
    bool drained = !r->events.empty() || r->host_sort_only;      // (then the host vector stays the one home of the trace)
    for (;;) {
        CU(cuMemsetD8Async(r->ctl.p, 0, 8, G.stream));  // barrier counter
        void* args[] = { &P };
        CU(cuEventRecord(G.ev0, G.stream));
        CU(cuLaunchCooperativeKernel(G.f_run[r->kidx()], (unsigned)r->grid, 1, 1, (unsigned)G.run_block[ZZ_RUN_BLOCK_OF(r)], 1, 1, dyn_smem, G.stream, args));
        CU(cuEventRecord(G.ev1, G.stream));
        CU(cuStreamSynchronize(G.stream));
        r->launches++;
        float ms = 0.f;
        CU(cuEventElapsedTime(&ms, G.ev0, G.ev1));
        total_ms += ms;
        CU(cuMemcpyDtoH(&r->hc, r->ctl.p, sizeof(ZzDevCtl)));
        const ZzDevCtl& hc = r->hc;
        r->wat_next = hc.wattempt + 1u;
        if (hc.trace_full) return fail(ZZB_E_INTERNAL, "trace record dropped (internal protocol error)");
        if (hc.viol) break;
        if (hc.ctl.phase == ZZ_PH_FAIL) return fail(ZZB_E_INTERNAL, "window controller failed");
        const bool last = (hc.ctl.phase == ZZ_PH_DONE) || (r->max_windows && !hc.need_drain);
        if (P.record_trace && last && !drained && !hc.need_drain && hc.trace_len) {
            // everything this call produced is still in HBM: order it there; zzb_trace_copy reads the sorted buffer directly
            bool ok = false;
            int32_t st = device_sort_trace(r, hc.trace_len, std::min(t_front0, hc.ctl.F), hc.ctl.F, &ok);
            if (st) return st;
            if (ok) {
                r->dev_sorted = true;
                ZzDevCtl z = hc; z.trace_len = 0; z.need_drain = 0;
                CU(cuMemcpyHtoD(r->ctl.p, &z, sizeof z));
                r->hc.trace_len = 0;
                break;
            }
        }
        if (P.record_trace && (hc.need_drain || hc.ctl.phase == ZZ_PH_DONE || r->max_windows)) {
            drained = true;
            if (hc.need_drain && hc.trace_len == 0) return fail(ZZB_E_TRACE, "trace buffer (%llu records) too small for one window", r->trace_cap);
            std::vector<zzb_event> chunk((size_t)hc.trace_len);
            if (hc.trace_len) CU(cuMemcpyDtoH(chunk.data(), r->trace.p, (size_t)hc.trace_len * sizeof(zzb_event)));
            absorb_trace_chunk(r, chunk);
            // reset trace_len and need_drain on the device
            ZzDevCtl z = hc; z.trace_len = 0; z.need_drain = 0;
            CU(cuMemcpyHtoD(r->ctl.p, &z, sizeof z));
        }
        if (hc.ctl.phase == ZZ_PH_DONE) break;
        if (r->max_windows && !hc.need_drain) break;  // caller asked for a bounded slice
    }
    if (device_ms) *device_ms = total_ms;
    if (P.dbgbuf) {
        r->dbghost.resize((size_t)r->grid * ZZ_DBG_REC * 4 + 4096 * 16);
        CU(cuMemcpyDtoH(r->dbghost.data(), r->dbgbuf.p, r->dbghost.size() * 8));
        if (const char* f = getenv("ZZB200_DBG_FILE")) { FILE* fp = fopen(f, "wb"); if (fp) { fwrite(r->dbghost.data(), 8, r->dbghost.size(), fp); fclose(fp); } }
    }
    r->executed = true; r->fetched = false;
    if (r->hc.viol == 2u) return fail(ZZB_E_INTERNAL, "sticky sampler: a freezing coordinate was not at 0 (ss_fact.jl:89-91)");
    if (r->hc.viol) {
        if (P.record_trace && r->hc.trace_len) {
            std::vector<zzb_event> chunk((size_t)r->hc.trace_len);
            CU(cuMemcpyDtoH(chunk.data(), r->trace.p, chunk.size() * sizeof(zzb_event)));
            absorb_trace_chunk(r, chunk);
        }
        return fail(ZZB_E_BOUND, "Tuning parameter `c` too small. (coordinate %d, t = %.17g, l = %.17g, lb = %.17g)",
                    r->hc.viol_i, r->hc.viol_t, r->hc.viol_l, r->hc.viol_lb);
    }
    return ZZB_OK;
}

static int32_t fetch_state(zzb_run_s* r)
{
    if (r->fetched) return ZZB_OK;
    if (!r->uploaded) return fail(ZZB_E_ARG, "run has no state yet");
    CtxGuard cg;
    const size_t d = (size_t)r->d;
    const unsigned grid = (unsigned)std::min<size_t>((d + ZZ_BLOCK - 1) / ZZ_BLOCK, (size_t)G.sm_count * 8);
    CUdeviceptr pt = r->out_t.p, px = r->out_x.p, pth = r->out_th.p, pc = r->out_c.p, pa = r->out_acc.p;
    void* a[] = { &r->P, &pt, &px, &pth, &pc, &pa };
    CU(cuLaunchKernel(G.f_export, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    r->launches++;
    r->ft.resize(d); r->fx.resize(d); r->fth.resize(d); r->fc.resize(d); r->facc.resize(d); r->hs1.resize(d); r->hs2.resize(d);
    CU(cuMemcpyDtoHAsync(r->ft.data(), pt, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->fx.data(), px, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->fth.data(), pth, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->fc.data(), pc, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->facc.data(), pa, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->hs1.data(), r->s1.p, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(r->hs2.data(), r->s2.p, d * 8, G.stream));
    CU(cuMemcpyDtoHAsync(&r->hc, r->ctl.p, sizeof(ZzDevCtl), G.stream));
    CU(cuStreamSynchronize(G.stream));
    r->fetched = true;
    return ZZB_OK;
}

int32_t zzb_spdmp_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                      const uint64_t* seed, int32_t adapt, double factor, uint32_t flags, zzb_run_t* out)
{
    if (!out || !c) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags, 0, &r);
    if (st) return st;
    st = zzb_run_upload(r, t0, x0, theta0, c, seed, adapt, factor);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st && st != ZZB_E_BOUND) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    memcpy(c, r->fc.data(), (size_t)r->d * 8);  // adapted bounds (sfact.jl:211)
    *out = r;
    if (st == ZZB_E_BOUND)
        fail(ZZB_E_BOUND, "Tuning parameter `c` too small. (coordinate %d, t = %.17g, l = %.17g, lb = %.17g)", r->hc.viol_i,
             r->hc.viol_t, r->hc.viol_l, r->hc.viol_lb);
    return st;
}

int32_t zzb_sspdmp_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, const double* c,
                       const double* kappa, const uint64_t* seed, uint32_t flags, zzb_run_t* out)
{
    if (!out || !c || !kappa) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_STICKY, 0, &r);
    if (st) return st;
    st = zzb_run_upload_kappa(r, kappa);
    if (!st) st = zzb_run_upload(r, t0, x0, theta0, c, seed, 0, 1.0);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st && st != ZZB_E_BOUND) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    *out = r;
    return st;
}

// sspdmp(...; adapt = true, factor) (src/ss_fact.jl:132-136,159): zzb_sspdmp_run with the adaptation of the bounds; c is in/out
int32_t zzb_sspdmp_adapt_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                             const double* kappa, const uint64_t* seed, int32_t adapt, double factor, uint32_t flags, zzb_run_t* out)
{
    if (!out || !c || !kappa) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_STICKY, 0, &r);
    if (st) return st;
    st = zzb_run_upload_kappa(r, kappa);
    if (!st) st = zzb_run_upload(r, t0, x0, theta0, c, seed, adapt, factor);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st && st != ZZB_E_BOUND) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    memcpy(c, r->fc.data(), (size_t)r->d * 8);
    *out = r;
    return st;
}

// sspdmp3 / sparsestickyzz (src/sparsestickyzz.jl:405-422,192-257): the strong-bound sparse sticky ZigZag as the reference runs
// BASELINE config 4 -- one scalar bound constant c (SparseStickyUpperBounds, :127-142, adapt = false), one thaw rate kappa
// (StickyBarriers), rule 0 = :sticky / 1 = :reversible; coordinates with x0 == 0 start frozen (sparsestickystate, :10-12),
// theta0 gives the velocities of the others.  Time starts at 0 like the reference's SparseState.
int32_t zzb_sspdmp3_run(zzb_problem_t p, const double* x0, const double* theta0, double T, double c, double kappa, int32_t rule,
                        const uint64_t* seed, uint32_t flags, zzb_run_t* out)
{
    if (!out || !x0 || !theta0 || !seed) return fail(ZZB_E_ARG, "null argument");
    if (!p) return fail(ZZB_E_ARG, "null argument");
    if (!(c > 0.0) || !(kappa > 0.0) || (rule != 0 && rule != 1)) return fail(ZZB_E_ARG, "sspdmp3 needs c > 0, kappa > 0 and rule 0 (:sticky) or 1 (:reversible)");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_STICKY, 0, &r);
    if (st) return st;
    const size_t d = (size_t)r->d;
    std::vector<double> kap(d, kappa), cv(d, c);
    st = zzb_run_upload_kappa(r, kap.data());
    if (!st) st = zzb_run_set(r, "strong_c", c);
    if (!st) st = zzb_run_set(r, "strong_rule", (double)rule);
    if (!st) st = zzb_run_upload(r, 0.0, x0, theta0, cv.data(), seed, 0, 1.0);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    *out = r;
    return ZZB_OK;
}

// sspdmp4 / asynchzz (src/asynchzz.jl:250-265,80-147): the strong-bound sticky ZigZag with a bound constant c[i] and a thaw rate
// kappa[i] per coordinate (StrongUpperBounds :2-7,20-28; StickyBarriers((0,0), (:sticky,:sticky), (kappa_i,kappa_i)) :258), start
// time t0; coordinates with x0 == 0 start frozen and continue, once thawed, with theta0 (:112-116,206-213).  The reference runs
// this process with its own parallel schedule (local minima of a PartialQueue, regions coloured for Threads.@threads, :150-245);
// the device runs it with the windowed relaxation like sspdmp3 -- the law is the same, the schedule is not part of it.
int32_t zzb_sspdmp4_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, const double* c,
                        const double* kappa, const uint64_t* seed, uint32_t flags, zzb_run_t* out)
{
    if (!out || !p || !x0 || !theta0 || !c || !kappa || !seed) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_STICKY, 0, &r);
    if (st) return st;
    for (int32_t j = 0; j < r->d; ++j)
        if (!(c[j] > 0.0) || !(kappa[j] > 0.0)) { zzb_run_free(r); return fail(ZZB_E_ARG, "sspdmp4 needs c[i] > 0 and kappa[i] > 0"); }
    st = zzb_run_upload_kappa(r, kappa);
    if (!st) st = zzb_run_set(r, "strong_c", c[0]);
    if (!st) st = zzb_run_set(r, "strong_rule", 2.0);
    if (!st) st = zzb_run_upload(r, t0, x0, theta0, c, seed, 0, 1.0);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    *out = r;
    return ZZB_OK;
}

int32_t zzb_spdmp_boomerang_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                                const double* sigma, double lambdaref, double rho, const uint64_t* seed, int32_t adapt,
                                double factor, uint32_t flags, zzb_run_t* out)
{
    if (!out || !c || !sigma) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_BOOMERANG, 0, &r);
    if (st) return st;
    st = zzb_run_upload_boomerang(r, sigma, lambdaref, rho);
    if (!st) st = zzb_run_upload(r, t0, x0, theta0, c, seed, adapt, factor);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st && st != ZZB_E_BOUND) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    memcpy(c, r->fc.data(), (size_t)r->d * 8);
    *out = r;
    if (st == ZZB_E_BOUND)
        fail(ZZB_E_BOUND, "Tuning parameter `c` too small. (coordinate %d, t = %.17g, l = %.17g, lb = %.17g)", r->hc.viol_i,
             r->hc.viol_t, r->hc.viol_l, r->hc.viol_lb);
    return st;
}

// spdmp / pdmp with Z = ZigZag(Gamma, mu, sigma; lambdaref > 0): the refresh branch of src/sfact.jl:78-114,188-190 (one-call form)
int32_t zzb_spdmp_refresh_run(zzb_problem_t p, double t0, const double* x0, const double* theta0, double T, double* c,
                              const double* sigma, double lambdaref, const uint64_t* seed, int32_t adapt, double factor,
                              uint32_t flags, zzb_run_t* out)
{
    if (!out || !c || !sigma) return fail(ZZB_E_ARG, "null argument");
    zzb_run_t r = nullptr;
    int32_t st = zzb_run_create(p, flags | ZZB_FLAG_REFRESH, 0, &r);
    if (st) return st;
    st = zzb_run_upload_refresh(r, sigma, lambdaref);
    if (!st) st = zzb_run_upload(r, t0, x0, theta0, c, seed, adapt, factor);
    if (!st) st = zzb_run_execute(r, T, nullptr);
    if (st && st != ZZB_E_BOUND) { zzb_run_free(r); return st; }
    int32_t st2 = fetch_state(r);
    if (st2) { zzb_run_free(r); return st2; }
    memcpy(c, r->fc.data(), (size_t)r->d * 8);
    *out = r;
    if (st == ZZB_E_BOUND)
        fail(ZZB_E_BOUND, "Tuning parameter `c` too small. (coordinate %d, t = %.17g, l = %.17g, lb = %.17g)", r->hc.viol_i,
             r->hc.viol_t, r->hc.viol_l, r->hc.viol_lb);
    return st;
}

// Device -> caller buffers in one go (any pointer may be NULL): final (t, x, theta), adapted c, per-coordinate accepted
// counts, moment sums, total proposals.  With pinned destination buffers the copies run at full PCIe speed; this is
// the read-back used by the one-call path of a host wrapper that wants no intermediate copies.
int32_t zzb_run_fetch(zzb_run_t r, double* t, double* x, double* theta, double* c, int64_t* acc, double* s1, double* s2,
                      int64_t* num, int64_t* nacc)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!r->uploaded) return fail(ZZB_E_ARG, "run has no state yet");
    CtxGuard cg;
    const size_t d = (size_t)r->d;
    const unsigned grid = (unsigned)std::min<size_t>((d + ZZ_BLOCK - 1) / ZZ_BLOCK, (size_t)G.sm_count * 8);
    CUdeviceptr pt = r->out_t.p, px = r->out_x.p, pth = r->out_th.p, pc = r->out_c.p, pa = r->out_acc.p;
    void* a[] = { &r->P, &pt, &px, &pth, &pc, &pa };
    CU(cuLaunchKernel(G.f_export, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    r->launches++;
    // a sharded run holds results for its owned coordinates only: copy [lo, hi) into the same positions of the caller's arrays
    const size_t lo = r->nranks > 1 ? (size_t)r->lo : 0, n = r->nranks > 1 ? (size_t)(r->hi - r->lo) : d, o8 = lo * 8, n8 = n * 8;
    if (n) {
        if (t) CU(cuMemcpyDtoHAsync(t + lo, pt + o8, n8, G.stream));
        if (x) CU(cuMemcpyDtoHAsync(x + lo, px + o8, n8, G.stream));
        if (theta) CU(cuMemcpyDtoHAsync(theta + lo, pth + o8, n8, G.stream));
        if (c) CU(cuMemcpyDtoHAsync(c + lo, pc + o8, n8, G.stream));
        if (acc) CU(cuMemcpyDtoHAsync(acc + lo, pa + o8, n8, G.stream));
        if (s1) CU(cuMemcpyDtoHAsync(s1 + lo, r->s1.p + o8, n8, G.stream));
        if (s2) CU(cuMemcpyDtoHAsync(s2 + lo, r->s2.p + o8, n8, G.stream));
    }
    CU(cuMemcpyDtoHAsync(&r->hc, r->ctl.p, sizeof(ZzDevCtl), G.stream));
    CU(cuStreamSynchronize(G.stream));
    if (num) *num = (int64_t)r->hc.num;
    if (nacc) *nacc = (int64_t)r->hc.nacc;
    return ZZB_OK;
}

int32_t zzb_run_stats(zzb_run_t r, int64_t* out, int32_t n)
{
    if (!r || !out) return fail(ZZB_E_ARG, "null argument");
    int32_t st = fetch_state(r);
    if (st) return st;
    int64_t v[24] = { (int64_t)r->hc.windows, (int64_t)r->hc.retries, (int64_t)r->hc.iters, (int64_t)r->hc.node_evals,
                      (int64_t)r->hc.rebases, r->launches, (int64_t)r->grid, (int64_t)G.run_block[r->kind] };
    for (int k = 0; k < 8; ++k) { v[8 + k] = (int64_t)r->hc.tprof[k]; v[16 + k] = (int64_t)r->hc.dbg[k]; }
    for (int32_t k = 0; k < n && k < 24; ++k) out[k] = v[k];
    return ZZB_OK;
}

int32_t zzb_run_counts(zzb_run_t r, int64_t* acc, int64_t* num)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    int32_t st = fetch_state(r);
    if (st) return st;
    if (acc) for (int32_t j = 0; j < r->d; ++j) acc[j] = r->facc[j];
    if (num) *num = (int64_t)r->hc.num;
    return ZZB_OK;
}

int32_t zzb_run_final_state(zzb_run_t r, double* t, double* x, double* theta, double* c)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    int32_t st = fetch_state(r);
    if (st) return st;
    const size_t nb = (size_t)r->d * 8;
    if (t) memcpy(t, r->ft.data(), nb);
    if (x) memcpy(x, r->fx.data(), nb);
    if (theta) memcpy(theta, r->fth.data(), nb);
    if (c) memcpy(c, r->fc.data(), nb);
    return ZZB_OK;
}

int32_t zzb_trace_len(zzb_run_t r, int64_t* n)
{
    if (!r || !n) return fail(ZZB_E_ARG, "null argument");
    if (r->flags & ZZB_FLAG_NO_TRACE) {  // events were counted, not stored
        int32_t st = fetch_state(r);
        if (st) return st;
        *n = (int64_t)r->hc.nacc;
        return ZZB_OK;
    }
    *n = (int64_t)(r->events.size() + (r->dev_sorted ? (size_t)r->n_sorted : 0));
    return ZZB_OK;
}

int32_t zzb_trace_copy(zzb_run_t r, zzb_event* dst, int64_t first, int64_t count)
{
    if (!r || !dst) return fail(ZZB_E_ARG, "null argument");
    if (r->flags & ZZB_FLAG_NO_TRACE) return fail(ZZB_E_ARG, "run was created with ZZB_FLAG_NO_TRACE");
    const size_t nh = r->events.size(), nd = r->dev_sorted ? (size_t)r->n_sorted : 0;
    if (first < 0 || count < 0 || (size_t)(first + count) > nh + nd) return fail(ZZB_E_ARG, "trace range out of bounds");
    size_t f = (size_t)first, c = (size_t)count;
    if (f < nh) {   // part that was drained to the host during the run
        const size_t k = std::min(c, nh - f);
        memcpy(dst, r->events.data() + f, k * sizeof(zzb_event));
        dst += k; f += k; c -= k;
    }
    if (c) {        // part ordered on the device: straight from HBM into the caller's buffer
        CtxGuard cg;
        CU(cuMemcpyDtoH(dst, r->trace_sorted.p + (f - nh) * sizeof(zzb_event), c * sizeof(zzb_event)));
    }
    return ZZB_OK;
}

// Forget the events handed out so far (streaming use: execute a bounded number of windows, copy, clear, repeat).
int32_t zzb_trace_clear(zzb_run_t r)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    r->events.clear();
    r->dev_sorted = false; r->n_sorted = 0;
    return ZZB_OK;
}

int32_t zzb_trace_sums(zzb_run_t r, double* s1, double* s2)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    int32_t st = fetch_state(r);
    if (st) return st;
    if (s1) memcpy(s1, r->hs1.data(), (size_t)r->d * 8);
    if (s2) memcpy(s2, r->hs2.data(), (size_t)r->d * 8);
    return ZZB_OK;
}

int32_t zzb_trace_moments(zzb_run_t r, double* m1, double* m2)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    int32_t st = fetch_state(r);
    if (st) return st;
    // T = time of the last event (trace.jl:186) = the window end of phase C = the frontier when done
    const double Tl = r->hc.ctl.F;
    for (int32_t j = 0; j < r->d; ++j) {
        if (m1) m1[j] = r->hs1[j] * (1 / (2 * Tl));   // scale = 1/(2T), trace.jl:190
        if (m2) m2[j] = r->hs2[j] / (3 * Tl);
    }
    return ZZB_OK;
}

// subtrace at the source (src/trace.jl:275-290): record only the events of the coordinates J (1-based, strictly ascending),
// renumbered to their position in J; nJ = 0 lifts the filter.  Before zzb_run_upload / zzb_run_reset.
int32_t zzb_run_trace_filter(zzb_run_t r, const int64_t* J, int64_t nJ)
{
    if (!r || (nJ > 0 && !J)) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (nJ <= 0) { r->have_map = false; r->uploaded = false; return ZZB_OK; }
    std::vector<int32_t> map((size_t)r->d, 0);
    for (int64_t q = 0; q < nJ; ++q) {
        if (J[q] < 1 || J[q] > r->d || (q && J[q] <= J[q - 1])) return fail(ZZB_E_ARG, "J must be strictly ascending indices in 1..d");
        map[(size_t)J[q] - 1] = (int32_t)(q + 1);
    }
    CtxGuard cg;
    int32_t st = upload(r->trace_map, map.data(), map.size() * 4);
    if (st) return st;
    r->have_map = true;
    r->uploaded = false;   // the launch parameters change: zzb_run_upload / zzb_run_reset must follow
    return ZZB_OK;
}

// inclusion_prob(trace) of src/trace.jl:161-178 for a sticky run, from the device accumulator: fraction of [t0, last event] a
// coordinate's segments have a non-zero end.
int32_t zzb_trace_inclusion(zzb_run_t r, double* p)
{
    if (!r || !p) return fail(ZZB_E_ARG, "null argument");
    if (!(r->flags & ZZB_FLAG_STICKY)) return fail(ZZB_E_ARG, "inclusion probabilities are accumulated by the sticky samplers only");
    int32_t st = fetch_state(r);
    if (st) return st;
    CtxGuard cg;
    CU(cuMemcpyDtoH(p, r->s3.p, (size_t)r->d * 8));
    const double Tl = r->hc.ctl.F;
    for (int32_t j = 0; j < r->d; ++j) p[j] = p[j] / Tl;
    return ZZB_OK;
}

// cummean(trace) (src/trace.jl:203-225): per coordinate the running time average y_i / (2 t) after each of its events, y_i the
// running sum of (x_prev + x)(t - t_prev).  CSR-shaped output: the values of coordinate k (0-based) are entries
// offsets[k] .. offsets[k+1]-1 of times[] / values[] (the reference's leading pair (t0, x0_k) is left to the caller).  When the
// trace is still in HBM (ordered there by the last execute) everything happens on the device: the records are grouped by
// coordinate with the bucket kernels of the trace sort (bucket = coordinate), one thread per coordinate orders and walks its
// records (zz_cummean_kernel); otherwise, or when a coordinate has more than 64 events, the host walks the trace.
struct ZzCummeanHost { CUdeviceptr ev, base, x0; double t0; int32_t d, pad; CUdeviceptr times, values, ovf; };   // = ZzCummean (zz_kernels.cu)
int32_t zzb_trace_cummean(zzb_run_t r, int64_t* offsets, double* times, double* values)
{
    if (!r || !offsets) return fail(ZZB_E_ARG, "null argument");
    if (r->flags & ZZB_FLAG_NO_TRACE) return fail(ZZB_E_ARG, "the run does not record a trace");
    if (r->flags & ZZB_FLAG_BOOMERANG) return fail(ZZB_E_ARG, "cummean assumes piecewise linear paths (not the Boomerang flow)");
    if (r->have_map) return fail(ZZB_E_ARG, "cummean is not available together with a trace filter");
    if (!r->executed) return fail(ZZB_E_ARG, "zzb_run_execute must precede zzb_trace_cummean");
    CtxGuard cg;
    const size_t d = (size_t)r->d;
    const unsigned long long nd = r->dev_sorted ? r->n_sorted : 0ULL;
    const size_t n = r->events.size() + (size_t)nd;
    if (n && (!times || !values)) return fail(ZZB_E_ARG, "null argument");
    bool done = false;
    if (nd && r->events.empty() && nd < 0xfffffff0ULL) {
        const unsigned int nb = (unsigned int)d;
        const size_t wbytes = ((size_t)3 * nb + 2) * 4;
        int32_t st = ZZB_OK;
        if (r->cm_work.n < wbytes) st = r->cm_work.alloc(wbytes + 4096);
        if (!st && r->cm_out.n < (size_t)nd * 16) st = r->cm_out.alloc((size_t)nd * 16);
        if (!st) {
            CU(cuMemsetD8Async(r->cm_work.p, 0, wbytes, G.stream));
            ZzTsortHost Q;
            Q.in = r->trace_sorted.p; Q.out = r->trace.p; Q.n = nd; Q.tmin = 0.0; Q.scale = 0.0; Q.nb = nb; Q.mode = 1;
            Q.cnt = r->cm_work.p; Q.base = r->cm_work.p + (size_t)nb * 4; Q.fill = r->cm_work.p + ((size_t)2 * nb + 1) * 4; Q.ovf = r->cm_work.p + ((size_t)3 * nb + 1) * 4;
            void* a[] = { &Q };
            const unsigned grid = (unsigned)std::min<unsigned long long>((nd + ZZ_BLOCK - 1) / ZZ_BLOCK, (unsigned long long)G.sm_count * 16);
            CU(cuLaunchKernel(G.f_ts_hist, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
            CU(cuLaunchKernel(G.f_ts_scan, 1, 1, 1, 1024, 1, 1, 0, G.stream, a, nullptr));
            CU(cuLaunchKernel(G.f_ts_scatter, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
            ZzCummeanHost Cm;
            Cm.ev = r->trace.p; Cm.base = Q.base; Cm.x0 = r->in_x.p; Cm.t0 = r->t0; Cm.d = (int32_t)d; Cm.pad = 0;
            Cm.times = r->cm_out.p; Cm.values = r->cm_out.p + (size_t)nd * 8; Cm.ovf = Q.ovf;
            void* b[] = { &Cm };
            const unsigned gridc = (unsigned)std::min<size_t>((d + ZZ_BLOCK - 1) / ZZ_BLOCK, (size_t)G.sm_count * 16);
            CU(cuLaunchKernel(G.f_cummean, gridc, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, b, nullptr));
            r->launches += 4;
            std::vector<unsigned int> base(d + 1);
            unsigned int ovf = 0;
            CU(cuMemcpyDtoHAsync(base.data(), Q.base, (d + 1) * 4, G.stream));
            CU(cuMemcpyDtoHAsync(&ovf, Q.ovf, 4, G.stream));
            CU(cuStreamSynchronize(G.stream));
            if (!ovf && base[d] == nd) {
                CU(cuMemcpyDtoH(times, Cm.times, (size_t)nd * 8));
                CU(cuMemcpyDtoH(values, Cm.values, (size_t)nd * 8));
                for (size_t k = 0; k <= d; ++k) offsets[k] = (int64_t)base[k];
                done = true;
            }
        }
    }
    if (!done) {   // host: the trace in time order, one pass
        std::vector<zzb_event> ev(n);
        if (!r->events.empty()) memcpy(ev.data(), r->events.data(), r->events.size() * sizeof(zzb_event));
        if (nd) CU(cuMemcpyDtoH(ev.data() + r->events.size(), r->trace_sorted.p, (size_t)nd * sizeof(zzb_event)));
        std::vector<double> x0(d);
        CU(cuMemcpyDtoH(x0.data(), r->in_x.p, d * 8));
        std::vector<int64_t> cnt(d + 1, 0);
        for (const zzb_event& e : ev) cnt[(size_t)e.i]++;                  // (1-based ids land one slot up: exclusive scan below)
        offsets[0] = 0;
        for (size_t k = 0; k < d; ++k) offsets[k + 1] = offsets[k] + cnt[k + 1];
        std::vector<int64_t> fill(offsets, offsets + d);
        std::vector<double> y(d, 0.0), tp(d, r->t0), xp(x0);
        for (const zzb_event& e : ev) {
            const size_t k = (size_t)e.i - 1;
            y[k] += (xp[k] + e.x) * (e.t - tp[k]);                         // trace.jl:217
            tp[k] = e.t; xp[k] = e.x;
            const int64_t pos = fill[k]++;
            times[pos] = e.t; values[pos] = y[k] / (2 * e.t);              // trace.jl:222
        }
    }
    return ZZB_OK;
}

// Ask for the device-side discretisation x(t0 + k dt), k = 0 .. n_rows-1 (src/trace.jl:94-125); before zzb_run_upload.
int32_t zzb_run_discretize(zzb_run_t r, double dt, int64_t n_rows)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!(dt > 0.0) || n_rows < 1) return fail(ZZB_E_ARG, "discretize needs dt > 0 and n_rows >= 1");
    if (r->flags & ZZB_FLAG_STICKY) return fail(ZZB_E_ARG, "device-side discretize is not available for the sticky sampler");
    CtxGuard cg;
    int32_t st = r->gridbuf.alloc((size_t)n_rows * (size_t)r->d * 8);
    if (st) return st == ZZB_E_CUDA ? ZZB_E_NOMEM : st;
    r->grid_dt = dt; r->grid_n = n_rows;
    r->uploaded = false;   // the parameters change: zzb_run_upload / zzb_run_reset must follow
    return ZZB_OK;
}

int32_t zzb_run_grid(zzb_run_t r, double* xs, int64_t first_row, int64_t n, int64_t* valid_rows)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (!r->grid_n) return fail(ZZB_E_ARG, "zzb_run_discretize was not called");
    if (!r->uploaded) return fail(ZZB_E_ARG, "run has no state yet");
    if (first_row < 0 || n < 0 || first_row + n > r->grid_n) return fail(ZZB_E_ARG, "row range out of bounds");
    CtxGuard cg;
    CU(cuMemcpyDtoH(&r->hc, r->ctl.p, sizeof(ZzDevCtl)));
    const double tend = r->executed ? r->hc.ctl.F : r->t0;   // frontier: every event before it is final
    const unsigned grid = (unsigned)std::min<size_t>(((size_t)r->d + ZZ_BLOCK - 1) / ZZ_BLOCK, (size_t)G.sm_count * 8);
    double te = tend;
    void* a[] = { &r->P, &te };
    CU(cuLaunchKernel(G.f_grid_tail, grid, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    r->launches++;
    if (xs && n) CU(cuMemcpyDtoHAsync(xs, r->gridbuf.p + (size_t)first_row * (size_t)r->d * 8, (size_t)n * (size_t)r->d * 8, G.stream));
    CU(cuStreamSynchronize(G.stream));
    if (valid_rows) {
        long long k = (long long)floor((tend - r->t0) / r->grid_dt) + 2;
        if (k > r->grid_n) k = r->grid_n;
        while (k > 0 && r->t0 + (double)(k - 1) * r->grid_dt > tend) --k;
        *valid_rows = k;
    }
    return ZZB_OK;
}

// Test probe (tests/test_gpu_math.py): the device build of the scalar primitives of zz_math.h on host-supplied arguments.
int32_t zzb_math_probe(int32_t kind, int64_t n, const double* x, const double* y, const double* z, double* o1, double* o2)
{
    if (!G.ready) return fail(ZZB_E_CUDA, "zzb_init has not succeeded");
    if (n < 1 || !x || !o1 || kind < 0 || kind > 4) return fail(ZZB_E_ARG, "bad argument");
    CtxGuard cg;
    const size_t nb = (size_t)n * 8, nin = (kind == 4) ? 8 : nb;
    DevBuf dx, dy, dz, d1, d2;
    int32_t st = dx.alloc(nin);
    if (!st) st = dy.alloc(nin);
    if (!st) st = dz.alloc(nin);
    if (!st) st = d1.alloc(nb);
    if (!st) st = d2.alloc(nb);
    if (st) return st;
    CU(cuMemcpyHtoD(dx.p, x, nin));
    if (y) CU(cuMemcpyHtoD(dy.p, y, nin));
    if (z) CU(cuMemcpyHtoD(dz.p, z, nin));
    int k = kind; long long nn = n;
    void* a[] = { &k, &nn, &dx.p, &dy.p, &dz.p, &d1.p, &d2.p };
    CU(cuLaunchKernel(G.f_math_probe, (unsigned)G.sm_count * 8, 1, 1, ZZ_BLOCK, 1, 1, 0, G.stream, a, nullptr));
    CU(cuStreamSynchronize(G.stream));
    CU(cuMemcpyDtoH(o1, d1.p, nb));
    if (o2) CU(cuMemcpyDtoH(o2, d2.p, nb));
    return ZZB_OK;
}

int32_t zzb_run_error_info(zzb_run_t r, int64_t* i, double* t, double* l, double* lb)
{
    if (!r) return fail(ZZB_E_ARG, "null argument");
    if (i) *i = r->hc.viol_i;
    if (t) *t = r->hc.viol_t;
    if (l) *l = r->hc.viol_l;
    if (lb) *lb = r->hc.viol_lb;
    return ZZB_OK;
}

int32_t zzb_run_free(zzb_run_t r)
{
    if (!r) return ZZB_OK;
    CtxGuard cg;
    for (int q = 0; q < ZZ_MAXRANKS; ++q)
        if (r->peer_open[q]) for (int k = 0; k < 8; ++k) if (r->peer[q][k]) g_drv.p_cuIpcCloseMemHandle(r->peer[q][k]);
    delete r;
    return ZZB_OK;
}

}  // extern "C"
