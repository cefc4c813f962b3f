"""ctypes binding of oracle/libzzoracle.so (TEST INFRASTRUCTURE: the checker, never the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

RNG_SEQ, RNG_CTR = 0, 1
ARITH_INPLACE, ARITH_LAZY = 0, 2
GRAPH_ALL = 4
LOCAL_BOUND = 8
STICKY_REVERSIBLE, STICKY_STRONG_UB = 16, 32   # sspdmp(...; reversible / strong_upperbounds), src/ss_fact.jl:97,111
STICKYZZ = 64                                  # the dense sticky sampler stickyzz / sspdmp2 (src/stickyzz.jl): rate floor, frozen start
PARITY_MODE = RNG_CTR | ARITH_LAZY  # what the GPU must reproduce bit for bit

EVENT_DTYPE = np.dtype([("t", "<f8"), ("i", "<i8"), ("x", "<f8"), ("theta", "<f8")])  # trace.jl:38

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "libzzoracle.so")
        srcs = [os.path.join(ORACLE_DIR, "zz_oracle.c"), os.path.join(ROOT, "zigzagboomerang.jl_b200", "csrc", "zz_math.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            build()
        L = C.CDLL(so)
        L.zzo_spdmp.restype = C.c_void_p
        L.zzo_spdmp.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.zzo_sspdmp.restype = C.c_void_p
        L.zzo_sspdmp.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.zzo_spdmp_boom.restype = C.c_void_p
        L.zzo_spdmp_boom.argtypes = [C.c_int64] + [C.c_void_p] * 9 + [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.zzo_parallel_spdmp.restype = C.c_void_p
        L.zzo_parallel_spdmp.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_double]
        L.zzo_spdmp_logistic.restype = C.c_void_p
        L.zzo_spdmp_logistic.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_double, C.c_int64] + [C.c_void_p] * 4 + [
            C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.zzo_exp.restype = C.c_double
        L.zzo_exp.argtypes = [C.c_double]
        L.zzo_sparsestickyzz.restype = C.c_void_p
        L.zzo_sparsestickyzz.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.zzo_spdmp_refresh.restype = C.c_void_p
        L.zzo_spdmp_refresh.argtypes = [C.c_int64] + [C.c_void_p] * 9 + [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.zzo_sparsestickyzz_ctr.restype = C.c_void_p
        L.zzo_sparsestickyzz_ctr.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.zzo_strongsticky_ctr.restype = C.c_void_p
        L.zzo_strongsticky_ctr.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_void_p]
        L.zzo_queue_script.restype = None
        L.zzo_queue_script.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_void_p]
        L.zzo_sincos.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        L.zzo_randn.restype = C.c_double
        L.zzo_randn.argtypes = [C.c_double, C.c_double]
        L.zzo_status.argtypes = [C.c_void_p]
        L.zzo_loop_seconds.restype = C.c_double
        L.zzo_loop_seconds.argtypes = [C.c_void_p]
        L.zzo_trace_len.restype = C.c_int64
        L.zzo_trace_len.argtypes = [C.c_void_p]
        L.zzo_trace_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.zzo_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.zzo_final_state.argtypes = [C.c_void_p] * 5
        L.zzo_moments.argtypes = [C.c_void_p] * 5
        L.zzo_error_info.argtypes = [C.c_void_p] * 5
        L.zzo_free.argtypes = [C.c_void_p]
        for f in (L.zzo_poisson_time, L.zzo_poisson_time3, L.zzo_log, L.zzo_u01):
            f.restype = C.c_double
        L.zzo_poisson_time.argtypes = [C.c_double] * 3
        L.zzo_poisson_time3.argtypes = [C.c_double] * 4
        L.zzo_log.argtypes = [C.c_double]
        L.zzo_u01.argtypes = [C.c_uint64] * 4
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleResult:
    pass


class BoundError(RuntimeError):
    pass


def block_diagonal(G, K):
    """Gamma2 of test/testparallel.jl:33-42: the entries of G that couple different chunks of d/K coordinates dropped."""
    from zzb200.problems import CSC
    csz = G.n // K
    cols = np.repeat(np.arange(G.n), np.diff(G.colptr))
    keep = (cols // csz) == ((G.rowval - 1) // csz)
    colptr = np.concatenate([[1], 1 + np.cumsum(np.bincount(cols[keep], minlength=G.n))]).astype(np.int64)
    return CSC(G.n, colptr, np.ascontiguousarray(G.rowval[keep]), np.ascontiguousarray(G.nzval[keep]))


def spdmp(target, bound, t0, x0, theta0, T, c, *, h=None, mu=None, seed=(1, 2), adapt=False, factor=1.8,
          mode=PARITY_MODE, kappa=None, boom=None, parallel=None, logistic=None, refresh=None):
    """Run the oracle.  ``target`` / ``bound`` are problems.CSC (target precision and the sampler's Z.Gamma).
    ``parallel = (K, Delta)`` runs the multithreaded parallel_spdmp (src/parallel.jl) on K threads.
    With ``kappa`` (thaw rates) the sticky sampler sspdmp (src/ss_fact.jl) is run instead of spdmp; with
    ``boom = (sigma, lambdaref, rho)`` the factorised Boomerang (F::FactBoomerang in src/sfact.jl)."""
    L = lib()
    d = bound.n
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, theta0, c = f8(x0), f8(theta0), f8(c)
    mu = np.zeros(d) if mu is None else f8(mu)
    h = None if h is None else f8(h)
    sd = np.array(seed, dtype=np.uint64)
    if refresh is not None:   # (sigma, lambdaref): ZigZag velocity refreshments in spdmp (sfact.jl:78-114), Z.lambdaref > 0
        sigma = f8(refresh[0])
        r = L.zzo_spdmp_refresh(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                                _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu), _p(sigma), float(refresh[1]),
                                float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor), int(mode))
    elif logistic is not None:   # dict(A, At, y, ny, mu, gamma0, k): the subsampled logistic target of scripts/logistic.jl (`target` unused)
        lg = logistic
        A, At = lg["A"], lg["At"]
        ly, lny, lmu = f8(lg["y"]), f8(lg["ny"]), f8(lg["mu"])
        r = L.zzo_spdmp_logistic(d, A.nrows, _p(A.colptr), _p(A.rowval), _p(A.nzval), _p(At.colptr), _p(At.rowval), _p(At.nzval),
                                 _p(ly), _p(lny), _p(lmu), float(lg["gamma0"]), int(lg["k"]),
                                 _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                                 float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor), int(mode))
    elif parallel is not None:   # (K threads, Delta): parallel_spdmp of src/parallel.jl; `bound` must be block diagonal
        r = L.zzo_parallel_spdmp(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                                 _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                                 float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor),
                                 int(parallel[0]), float(parallel[1]))
    elif boom is not None:
        sigma = f8(boom[0])
        r = L.zzo_spdmp_boom(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                             _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu), _p(sigma), float(boom[1]), float(boom[2]),
                             float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor), int(mode))
    elif kappa is not None:
        kappa = f8(kappa)
        L.zzo_sspdmp_adapt.restype = C.c_void_p
        L.zzo_sspdmp_adapt.argtypes = L.zzo_sspdmp.argtypes + [C.c_int, C.c_double]
        r = L.zzo_sspdmp_adapt(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                         _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                         float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(kappa), _p(sd), int(mode), int(adapt), float(factor))
    else:
        r = L.zzo_spdmp(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                        _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                        float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor), int(mode))
    try:
        out = _collect(L, r, d, t0, x0, theta0)
        if boom is None:   # the event-based moments assume a piecewise linear path (trace.jl:182-200)
            out.m1, out.m2, out.s1, out.s2 = (np.empty(d) for _ in range(4))
            L.zzo_moments(r, _p(out.m1), _p(out.m2), _p(out.s1), _p(out.s2))
        return out
    finally:
        L.zzo_free(r)


def queue_script(head_vals, head_first, keys0, vals0, ops):
    """Composite queue of src/morepriorityqueues.jl (LinearQueue head + heap tail): peek before and after each update."""
    L = lib()
    hv = np.ascontiguousarray(head_vals, dtype=np.float64)
    k0 = np.ascontiguousarray(keys0, dtype=np.int64); v0 = np.ascontiguousarray(vals0, dtype=np.float64)
    ok = np.ascontiguousarray([o[0] for o in ops], dtype=np.int64); ov = np.ascontiguousarray([o[1] for o in ops], dtype=np.float64)
    pk = np.zeros(len(ops) + 1, np.int64); pv = np.zeros(len(ops) + 1)
    maxkey = int(max([0] + list(k0) + list(ok)))
    L.zzo_queue_script(_p(hv), len(hv), int(head_first), _p(k0), _p(v0), len(k0), maxkey, _p(ok), _p(ov), len(ops), _p(pk), _p(pv))
    return list(zip(pk.tolist(), pv.tolist()))


def _collect(L, r, d, t0, x0, theta0):
    st = L.zzo_status(r)
    if st == 3:
        i = C.c_int64(); t = C.c_double(); l = C.c_double(); lb = C.c_double()
        L.zzo_error_info(r, C.byref(i), C.byref(t), C.byref(l), C.byref(lb))
        raise BoundError("Tuning parameter `c` too small. (i=%d t=%g l=%g lb=%g)" % (i.value, t.value, l.value, lb.value))
    if st != 0:
        raise RuntimeError("oracle failed with status %d" % st)
    out = OracleResult()
    n = L.zzo_trace_len(r)
    out.events = np.empty(n, dtype=EVENT_DTYPE)
    L.zzo_trace_copy(r, _p(out.events), 0, n)
    out.acc = np.empty(d, np.int64)
    num = C.c_int64()
    L.zzo_counts(r, _p(out.acc), C.byref(num))
    out.num = num.value
    out.t, out.x, out.theta, out.c = (np.empty(d) for _ in range(4))
    L.zzo_final_state(r, _p(out.t), _p(out.x), _p(out.theta), _p(out.c))
    out.t0, out.x0, out.theta0 = t0, x0.copy(), theta0.copy()
    out.loop_seconds = L.zzo_loop_seconds(r)
    return out


def sparsestickyzz(G, x0, theta0, T, c, kappa, *, h=None, rule="sticky", adapt=False, multiplier=1.5, seed=(1, 2), ctr=False):
    """The reference's sparse sticky ZigZag (src/sparsestickyzz.jl, `sspdmp3`), restated for the CPU (faithful draw order,
    mode seq).  Scalar ``c`` and ``kappa``; ``theta0`` gives the velocities of the coordinates with ``x0 != 0``.
    ``ctr=True``: the parity arithmetic (per-coordinate streams and thaw clocks, flip-anchored positions)."""
    L = lib()
    if ctr:
        assert not adapt
        d = G.n
        x0c, th0c = np.ascontiguousarray(x0, dtype=np.float64), np.ascontiguousarray(theta0, dtype=np.float64)
        hc = None if h is None else np.ascontiguousarray(h, dtype=np.float64)
        sd = np.array(seed, dtype=np.uint64)
        r = L.zzo_sparsestickyzz_ctr(d, _p(G.colptr), _p(G.rowval), _p(G.nzval), _p(hc), _p(x0c), _p(th0c), float(T), float(c), float(kappa),
                                     {"sticky": 0, "reversible": 1}[rule], _p(sd))
        try:
            return _collect(L, r, d, 0.0, x0c, th0c)
        finally:
            L.zzo_free(r)
    d = G.n
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, theta0 = f8(x0), f8(theta0)
    h = None if h is None else f8(h)
    sd = np.array(seed, dtype=np.uint64)
    r = L.zzo_sparsestickyzz(d, _p(G.colptr), _p(G.rowval), _p(G.nzval), _p(h), _p(x0), _p(theta0), float(T), float(c), float(kappa),
                             {"sticky": 0, "reversible": 1}[rule], int(adapt), float(multiplier), _p(sd))
    try:
        return _collect(L, r, d, 0.0, x0, theta0)
    finally:
        L.zzo_free(r)


STRONG_RULES = {"sticky": 0, "reversible": 1, "keep": 2}


def strongsticky(G, t0, x0, theta0, T, c, kappa, *, h=None, rule="keep", seed=(1, 2)):
    """The strong-bound sticky ZigZag with a bound constant ``c[i]`` and a thaw rate ``kappa[i]`` per coordinate, start time ``t0``
    (oracle/zz_oracle.c:zzo_strongsticky_ctr) -- with ``rule="keep"`` the process of ``asynchzz`` / ``sspdmp4``
    (src/asynchzz.jl): a thawed coordinate continues with the velocity it had when it froze."""
    L = lib()
    d = G.n
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0c, th0c, cv, kv = f8(x0), f8(theta0), f8(c), f8(kappa)
    hc = None if h is None else f8(h)
    sd = np.array(seed, dtype=np.uint64)
    r = L.zzo_strongsticky_ctr(d, _p(G.colptr), _p(G.rowval), _p(G.nzval), _p(hc), float(t0), _p(x0c), _p(th0c), float(T), _p(cv), _p(kv),
                               STRONG_RULES[rule], _p(sd))
    try:
        return _collect(L, r, d, t0, x0c, th0c)
    finally:
        L.zzo_free(r)


# ---------------------------------------------------------------------------------------------------
# Host emulation of the GPU schedule (oracle/zz_window_sim.cpp) -- test-only, see its header.
_wlib = None


def wlib():
    global _wlib
    if _wlib is None:
        so = os.path.join(ORACLE_DIR, "libzzwindowsim.so")
        srcs = [os.path.join(ORACLE_DIR, "zz_window_sim.cpp")] + [
            os.path.join(ROOT, "zigzagboomerang.jl_b200", "csrc", f) for f in ("zz_core.h", "zz_fast.h", "zz_logit.h", "zz_strong.h", "zz_ctl.h", "zz_host_graph.h", "zz_host_logit.h", "zz_math.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            build()
        L = C.CDLL(so)
        L.zzw_spdmp.restype = C.c_void_p
        L.zzw_spdmp.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double,
                                C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint32, C.c_int, C.c_void_p,
                                C.c_void_p, C.c_double, C.c_double]
        L.zzw_spdmp_logistic.restype = C.c_void_p
        L.zzw_spdmp_logistic.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_double, C.c_int64] + [C.c_void_p] * 4 + [
            C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint32]
        L.zzw_sparsesticky.restype = C.c_void_p
        L.zzw_sparsesticky.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_double,
                                       C.c_double, C.c_uint32]
        L.zzw_status.argtypes = [C.c_void_p]
        L.zzw_trace_len.restype = C.c_int64
        L.zzw_trace_len.argtypes = [C.c_void_p]
        L.zzw_trace_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.zzw_counts.argtypes = [C.c_void_p] * 3
        L.zzw_final_state.argtypes = [C.c_void_p] * 5
        L.zzw_sums.argtypes = [C.c_void_p] * 3
        L.zzw_stats.argtypes = [C.c_void_p] * 2
        L.zzw_pass_hist.argtypes = [C.c_void_p] * 2
        L.zzw_item_hist.argtypes = [C.c_void_p] * 2
        L.zzw_error_info.argtypes = [C.c_void_p] * 5
        L.zzw_free.argtypes = [C.c_void_p]
        _wlib = L
    return _wlib


def window_sim(target, bound, t0, x0, theta0, T, c, *, h=None, mu=None, seed=(1, 2), adapt=False, factor=1.8,
               delta0=0.01, target_frac=0.4, tag_limit=0x0F000000, local_bound=False, kappa=None, boom=None, logistic=None,
               strong=None, async_tiles=0, order_seed=1, reversible=False, strong_upperbounds=False, refresh=None, stickyzz=False):
    """``strong = rule`` ("sticky" / "reversible"): the strong-bound sparse sticky timeline (zz_strong.h) with scalar ``c`` and
    ``kappa``, target = ``bound``; contract: :func:`sparsestickyzz` with ``ctr=True``."""
    L = wlib()
    d = bound.n
    # which schedule: 0 = pass-synchronous relaxation; T > 0 = the asynchronous tile-local relaxation with T tiles, tiles and
    # items taking turns in an order derived from order_seed (csrc/zz_kernels.cu: zz_run_body_async)
    L.zzw_set_schedule.restype = None
    L.zzw_set_schedule.argtypes = [C.c_int, C.c_uint64]
    L.zzw_set_schedule(int(async_tiles), int(order_seed))
    f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, theta0, c = f8(x0), f8(theta0), f8(c)
    mu = np.zeros(d) if mu is None else f8(mu)
    h = None if h is None else f8(h)
    sd = np.array(seed, dtype=np.uint64)
    if refresh is not None:   # (sigma, lambdaref): ZigZag with velocity refreshments
        L.zzw_spdmp_refresh.restype = C.c_void_p
        r = L.zzw_spdmp_refresh(C.c_int64(d), _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                                _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu), _p(f8(refresh[0])), C.c_double(float(refresh[1])),
                                C.c_double(float(t0)), _p(x0), _p(theta0), C.c_double(float(T)), _p(c), _p(sd), C.c_int(int(adapt)), C.c_double(float(factor)),
                                C.c_double(float(delta0)), C.c_double(float(target_frac)), C.c_uint32(int(tag_limit)))
        r = C.c_void_p(r)
    elif strong == "keep":   # asynchzz / sspdmp4: per-coordinate c and kappa, start time, velocity kept over a freeze
        L.zzw_strongsticky.restype = C.c_void_p
        L.zzw_strongsticky.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_uint32]
        kv = f8(kappa)
        r = L.zzw_strongsticky(d, _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(h), float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(kv),
                               2, _p(sd), float(delta0), float(target_frac), int(tag_limit))
    elif strong is not None:
        r = L.zzw_sparsesticky(d, _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(h), _p(x0), _p(theta0), float(T), float(c[0]),
                               float(kappa), {"sticky": 0, "reversible": 1}[strong], _p(sd), float(delta0), float(target_frac), int(tag_limit))
    elif logistic is not None:
        lg = logistic
        A, At = lg["A"], lg["At"]
        ly, lny, lmu = f8(lg["y"]), f8(lg["ny"]), f8(lg["mu"])
        r = L.zzw_spdmp_logistic(d, A.nrows, _p(A.colptr), _p(A.rowval), _p(A.nzval), _p(At.colptr), _p(At.rowval), _p(At.nzval),
                                 _p(ly), _p(lny), _p(lmu), float(lg["gamma0"]), int(lg["k"]),
                                 _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                                 float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor),
                                 float(delta0), float(target_frac), int(tag_limit))
    else:
        r = L.zzw_spdmp(d, _p(target.colptr), _p(target.rowval), _p(target.nzval), _p(h),
                        _p(bound.colptr), _p(bound.rowval), _p(bound.nzval), _p(mu),
                        float(t0), _p(x0), _p(theta0), float(T), _p(c), _p(sd), int(adapt), float(factor),
                        float(delta0), float(target_frac), int(tag_limit), int(bool(local_bound)) | (2 if reversible else 0) | (4 if strong_upperbounds else 0) | (8 if stickyzz else 0),
                        _p(None if kappa is None else f8(kappa)),
                        _p(None if boom is None else f8(boom[0])), 0.0 if boom is None else float(boom[1]), 0.0 if boom is None else float(boom[2]))
    try:
        st = L.zzw_status(r)
        if st == 3:
            i = C.c_int64(); t = C.c_double(); l = C.c_double(); lb = C.c_double()
            L.zzw_error_info(r, C.byref(i), C.byref(t), C.byref(l), C.byref(lb))
            raise BoundError("Tuning parameter `c` too small. (i=%d t=%g l=%g lb=%g)" % (i.value, t.value, l.value, lb.value))
        if st != 0:
            raise RuntimeError("window sim failed with status %d" % st)
        out = OracleResult()
        n = L.zzw_trace_len(r)
        out.events = np.empty(n, dtype=EVENT_DTYPE)
        L.zzw_trace_copy(r, _p(out.events), 0, n)
        out.acc = np.empty(d, np.int64)
        num = C.c_int64()
        L.zzw_counts(r, _p(out.acc), C.byref(num))
        out.num = num.value
        out.t, out.x, out.theta, out.c = (np.empty(d) for _ in range(4))
        L.zzw_final_state(r, _p(out.t), _p(out.x), _p(out.theta), _p(out.c))
        out.s1, out.s2 = np.empty(d), np.empty(d)
        L.zzw_sums(r, _p(out.s1), _p(out.s2))
        st = np.zeros(5, np.int64)
        L.zzw_stats(r, _p(st))
        ph = np.zeros(64, np.int64)
        L.zzw_pass_hist(r, _p(ph))
        out.pass_hist = ph
        ih = np.zeros(64, np.int64)
        L.zzw_item_hist(r, _p(ih))
        out.item_hist = ih
        out.stats = dict(windows=int(st[0]), retries=int(st[1]), iters=int(st[2]), node_evals=int(st[3]), max_iters=int(st[4]))
        return out
    finally:
        L.zzw_free(r)


def assert_same_run(a, b, check_state=True):
    """Bit-exact comparison of two runs (events, counters, final state, moment sums)."""
    assert len(a.events) == len(b.events), (len(a.events), len(b.events))
    assert np.array_equal(a.events["i"], b.events["i"])
    for f in ("t", "x", "theta"):
        assert np.array_equal(a.events[f].view(np.uint64), b.events[f].view(np.uint64)), f
    assert a.num == b.num, (a.num, b.num)
    assert np.array_equal(a.acc, b.acc)
    if check_state:
        for f in ("t", "x", "theta", "c"):
            assert np.array_equal(getattr(a, f).view(np.uint64), getattr(b, f).view(np.uint64)), f
    if hasattr(a, "s1") and hasattr(b, "s1"):
        assert np.array_equal(a.s1.view(np.uint64), b.s1.view(np.uint64))
        assert np.array_equal(a.s2.view(np.uint64), b.s2.view(np.uint64))


def sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().zzo_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def boom_discretize(res, mu, dt):
    """``collect(discretize(trace, dt))`` for a FactBoomerang trace (src/trace.jl:94-125 with the rotation flow of
    src/dynamics.jl:29-36) as ``(ts, xs)``; plain numpy, vectorised over coordinates."""
    t, x, th = float(res.t0), res.x0.copy(), res.theta0.copy()
    mu = np.asarray(mu, dtype=np.float64)
    ts, xs = [t], [x.copy()]
    ev = res.events
    k, n = 0, len(ev)

    def move(x, th, tau):
        s, c = np.sin(tau), np.cos(tau)
        return (x - mu) * c + th * s + mu, -(x - mu) * s + th * c

    while True:
        step = dt
        done = False
        while True:
            if k >= n:
                done = True
                break
            ti, i, xi, thi = ev[k]
            if t + step < ti:
                x, th = move(x, th, step)
                t += step
                break
            dd = ti - t
            step -= dd
            x, th = move(x, th, dd)
            t = ti
            x[i - 1] = xi
            th[i - 1] = thi
            k += 1
        if done:
            break
        ts.append(t)
        xs.append(x.copy())
    return np.array(ts), np.array(xs)
