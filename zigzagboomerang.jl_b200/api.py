"""Host-side mirror of the reference's interface for the factorised local-ZigZag path, on top of the C-ABI.

Names, argument order and return shapes follow mschauer/ZigZagBoomerang.jl (paths relative to /root/reference):

    ZigZag(Gamma, mu, sigma; lambdaref, rho)          src/types.jl:19-27
    FactBoomerang(Gamma, mu, lambdaref, sigma; rho)   src/types.jl:62-79
    spdmp(grad, t0, x0, theta0, T, c, [G,] F, args...; factor=1.8, adapt=false, seed)
        -> Xi, (t, x, theta), (acc, num), c           src/sfact.jl:162-214
    pdmp(grad, t0, x0, theta0, T, c, F, args...)      src/sfact.jl:236
    Trace / FactTrace, discretize, mean, cummean, subtrace    src/trace.jl

A device kernel cannot call a host closure, so `grad` must be a :class:`GaussianPotential` descriptor (it is still
callable as ``grad(x, i)`` like the reference's ``grad(x, i, Gamma) = idot(Gamma, i, x)``).  Julia is not available in
this image; ``julia/ZigZagBoomerangB200.jl`` is the same wrapper written against the same C-ABI.
"""
from __future__ import annotations

import ctypes as C
import secrets

import numpy as np

from . import _capi
from ._capi import EVENT_DTYPE, ZZB_E_BOUND, BoundError, ZZBError, check, f8, ptr
from .problems import CSC


class All:
    """`All()` neighbourhood marker (src/sfact.jl:1-2)."""


class Matched:
    """`Matched()` neighbourhood marker (src/sfact.jl:3-4)."""


class ExtendedForm:
    """`ExtendedForm()` / `SelfMoving()` marker (src/types.jl:103-119): the closure takes the extended argument list and moves
    the coordinates it needs itself.  On the device path positions are flip-anchored, so the marker only documents intent."""


SelfMoving = ExtendedForm


class ZigZag:
    """``ZigZag(Gamma, mu, sigma=diag(Gamma).^(-0.5); lambdaref=0.0, rho=0.0)`` (src/types.jl:19-27)."""

    def __init__(self, Gamma: CSC, mu, sigma=None, *, lambdaref: float = 0.0, rho: float = 0.0):
        self.Gamma = Gamma if isinstance(Gamma, CSC) else CSC.from_scipy(Gamma)
        self.mu = f8(mu)
        if sigma is None:
            G = self.Gamma
            cols = np.repeat(np.arange(G.n), np.diff(G.colptr))
            on = (G.rowval - 1) == cols
            diag = np.zeros(G.n)
            diag[cols[on]] = G.nzval[on]
            with np.errstate(divide="ignore"):
                sigma = diag ** -0.5
        self.sigma = sigma
        self.lambdaref = float(lambdaref)
        self.rho = float(rho)
        self.rhobar = float(np.sqrt(1 - rho * rho))


class FactBoomerang:
    """``FactBoomerang(Gamma, mu, lambda, sigma=diag(Gamma).^(-0.5); rho=0.0)`` (src/types.jl:62-79): factorised Boomerang
    dynamics preserving N(mu, inv(Diagonal(Gamma))), refreshment rate ``lambda`` (must be > 0: ``hasrefresh`` is always
    true, src/fact_samplers.jl:18)."""

    def __init__(self, Gamma: CSC, mu, lambdaref: float, sigma=None, *, rho: float = 0.0):
        self.Gamma = Gamma if isinstance(Gamma, CSC) else CSC.from_scipy(Gamma)
        self.mu = f8(mu)
        if sigma is None:
            sigma = ZigZag(self.Gamma, self.mu).sigma
        self.sigma = f8(sigma)
        self.lambdaref = float(lambdaref)
        self.rho = float(rho)
        self.rhobar = float(np.sqrt(1 - rho * rho))


class LocalBound:
    """``LocalBound(c)`` (src/types.jl:121-123): bounds from the target's own first and second directional derivatives
    (src/local.jl:2-6), valid for ``2/c/|theta|`` and then renewed."""

    def __init__(self, c):
        self.c = f8(c)


class GaussianPotential:
    """Target descriptor: ``grad phi_i(x) = idot(Gamma, i, x) - h[i]`` (src/common.jl:16-24)."""

    def __init__(self, Gamma: CSC, h=None):
        self.Gamma = Gamma if isinstance(Gamma, CSC) else CSC.from_scipy(Gamma)
        self.h = None if h is None else f8(h)

    def __call__(self, x, i, *args):  # 1-based i, like the reference closure
        G = self.Gamma
        s = 0.0
        for p in range(G.colptr[i - 1] - 1, G.colptr[i] - 1):
            s += G.nzval[p] * x[G.rowval[p] - 1]
        return s - (0.0 if self.h is None else self.h[i - 1])


class LogisticSubsampled:
    """Target descriptor for the subsampled logistic regression of scripts/logistic.jl: stands for the closure
    ``grad_phi_moving(t, x, theta, i, t', F, A, At, mu, y, ny, k) = gamma0*x[i] - fdot_moving(...)`` (:78-107) together with
    the arguments ``SelfMoving(), A, At, mu, y, ny, k`` the reference forwards to it (:167).  ``A`` / ``At`` are
    :class:`problems.RectCSC` (design matrix and its transpose).  Calling it evaluates the FULL-data partial derivative
    ``grad_phi(x, i, A, At, y, ny)`` (:104) -- the expectation of the subsampled estimate."""

    def __init__(self, A, At, y, ny, mu, gamma0: float = 0.01, k: int = 10):
        self.A, self.At = A, (At if At is not None else A.transpose())
        self.y, self.ny, self.mu = f8(y), f8(ny), f8(mu)
        self.gamma0, self.k = float(gamma0), int(k)

    def __call__(self, x, i, *args):  # 1-based i
        A, At = self.A, self.At
        s = 0.0
        for p in range(A.colptr[i - 1] - 1, A.colptr[i] - 1):
            row = A.rowval[p]
            u = 0.0
            for q in range(At.colptr[row - 1] - 1, At.colptr[row] - 1):
                u += At.nzval[q] * x[At.rowval[q] - 1]
            sg = 1.0 / (1.0 + np.exp(-u))
            s += A.nzval[p] * self.y[row - 1] * (1.0 - sg) - A.nzval[p] * self.ny[row - 1] * sg
        return self.gamma0 * x[i - 1] - s


class FactTrace:
    """``FactTrace`` (src/trace.jl:7-13): initial triple plus the event list
    ``(t, i, x_i, theta_i)``; ``events`` is a structured array with Julia's tuple layout."""

    def __init__(self, F, t0, x0, theta0, events):
        self.F, self.t0, self.x0, self.theta0, self.events = F, float(t0), f8(x0), f8(theta0), events

    def __len__(self):  # Base.length(FT) = 1 + length(events), trace.jl:42
        return 1 + len(self.events)

    def __iter__(self):
        """Iterate ``(t, x)`` like src/trace.jl:44-63 (x is a fresh copy each step; the last event is not applied,
        exactly like the reference's `k == length(FT.events)` branch)."""
        t, x, th = self.t0, self.x0.copy(), self.theta0.copy()
        yield t, x.copy()
        n = len(self.events)
        for k in range(n):
            if k == n - 1:
                yield t, x.copy()
                return
            t2, i, xi, thi = self.events[k]
            x += th * (t2 - t)
            t = t2
            x[i - 1] = xi
            th[i - 1] = thi
            yield t, x.copy()


Trace = FactTrace


def _flow(F, x, th, tau):
    """move_forward of the trace's dynamics over tau: straight lines (ZigZag, src/dynamics.jl:12-17) or the rotation
    around F.mu (FactBoomerang, src/dynamics.jl:29-36); returns the new (x, theta)."""
    if isinstance(F, FactBoomerang):
        s, c = np.sin(tau), np.cos(tau)
        return (x - F.mu) * c + th * s + F.mu, -(x - F.mu) * s + th * c
    return x + th * tau, th


def discretize(trace: FactTrace, dt: float):
    """``collect(discretize(trace, dt))`` (src/trace.jl:94-125) as ``(ts, xs)`` arrays."""
    if isinstance(trace.F, FactBoomerang):
        return _discretize_flow(trace, dt)
    t, x, th = trace.t0, trace.x0.copy(), trace.theta0.copy()
    ts, xs = [t], [x.copy()]
    ev = trace.events
    k, n = 0, len(ev)
    while True:
        step = dt
        done = False
        while True:
            if k >= n:
                done = True
                break
            ti, i, xi, thi = ev[k]
            if t + step < ti:
                x += th * step
                t += step
                break
            d = ti - t
            step -= d
            x += th * d
            t = ti
            x[i - 1] = xi
            th[i - 1] = thi
            k += 1
        if done:
            break
        ts.append(t)
        xs.append(x.copy())
    return np.array(ts), np.array(xs)


def _discretize_flow(trace: FactTrace, dt: float):
    """The same iteration (src/trace.jl:106-125) with a general flow between events."""
    F = trace.F
    t, x, th = trace.t0, trace.x0.copy(), trace.theta0.copy()
    ts, xs = [t], [x.copy()]
    ev = trace.events
    k, n = 0, len(ev)
    while True:
        step = dt
        done = False
        while True:
            if k >= n:
                done = True
                break
            ti, i, xi, thi = ev[k]
            if t + step < ti:
                x, th = _flow(F, x, th, step)
                t += step
                break
            d = ti - t
            step -= d
            x, th = _flow(F, x, th, d)
            t = ti
            x[i - 1] = xi
            th[i - 1] = thi
            k += 1
        if done:
            break
        ts.append(t)
        xs.append(x.copy())
    return np.array(ts), np.array(xs)


def _previous_per_coordinate(tr: FactTrace):
    """For every event of a time-ordered trace: the time and position of the SAME coordinate's previous event (or its initial
    state).  Vectorised (one stable sort by coordinate) -- no Python loop over events."""
    ev = tr.events
    i = ev["i"].astype(np.int64) - 1
    order = np.argsort(i, kind="stable")
    si, st, sx = i[order], ev["t"][order], ev["x"][order]
    first = np.r_[True, si[1:] != si[:-1]] if len(si) else np.zeros(0, bool)
    pt = np.r_[tr.t0, st[:-1]] if len(si) else st
    px = np.r_[0.0, sx[:-1]] if len(si) else sx
    pt = np.where(first, tr.t0, pt)
    px = np.where(first, tr.x0[si] if len(si) else px, px)
    t_prev, x_prev = np.empty_like(pt), np.empty_like(px)
    t_prev[order], x_prev[order] = pt, px
    return i, t_prev, x_prev


def mean(trace: FactTrace):
    """``Statistics.mean(::Trace)`` (src/trace.jl:182-200).  The per-coordinate sums run in event order (np.add.at is
    unbuffered and sequential), i.e. with the roundings of the reference's loop."""
    ev = trace.events
    y = np.zeros(len(trace.x0))
    if not len(ev):
        return y
    T = ev["t"][-1]
    scale = 1 / (2 * T)
    i, t_prev, x_prev = _previous_per_coordinate(trace)
    np.add.at(y, i, (x_prev + ev["x"]) * (ev["t"] - t_prev) * scale)
    return y


def subtrace(tr: FactTrace, J):
    """``subtrace(tr, J)`` (src/trace.jl:275-290); J sorted, 1-based."""
    J = np.asarray(J, dtype=np.int64)
    assert np.all(np.diff(J) > 0)
    pos = np.searchsorted(J, tr.events["i"])
    ok = (pos < len(J)) & (J[np.minimum(pos, len(J) - 1)] == tr.events["i"])
    ev = tr.events[ok].copy()
    ev["i"] = pos[ok] + 1
    return FactTrace(tr.F, tr.t0, tr.x0[J - 1], tr.theta0[J - 1], ev)


class Problem:
    """Device-resident problem: target potential + sampler matrices (``zzb_problem_create_gaussian``)."""

    def __init__(self, target, Z: ZigZag):
        _capi.init()
        self.target, self.Z = target, Z
        if isinstance(target, LogisticSubsampled):   # zzb_problem_create_logistic
            A, At, Gb = target.A, target.At, Z.Gamma
            if A.ncols != Gb.n:
                raise ValueError("design matrix and sampler dimensions differ")
            self.d = Gb.n
            self._h = C.c_void_p()
            check(_capi.lib().zzb_problem_create_logistic(
                C.byref(self._h), self.d, A.nrows, ptr(A.colptr), ptr(A.rowval), ptr(A.nzval), ptr(At.colptr), ptr(At.rowval),
                ptr(At.nzval), ptr(target.y), ptr(target.ny), ptr(target.mu), target.gamma0, target.k,
                ptr(Gb.colptr), ptr(Gb.rowval), ptr(Gb.nzval), ptr(Z.mu)))
            return
        if target.Gamma.n != Z.Gamma.n:
            raise ValueError("target and sampler dimensions differ")
        self.d = target.Gamma.n
        self._h = C.c_void_p()
        Gt, Gb = target.Gamma, Z.Gamma
        check(_capi.lib().zzb_problem_create_gaussian(
            C.byref(self._h), self.d, ptr(Gt.colptr), ptr(Gt.rowval), ptr(Gt.nzval), ptr(target.h),
            ptr(Gb.colptr), ptr(Gb.rowval), ptr(Gb.nzval), ptr(Z.mu)))

    def close(self):
        if self._h:
            _capi.lib().zzb_problem_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Run:
    """One sampler run on the device (staged form of the C-ABI)."""

    def __init__(self, problem: Problem, *, record_trace: bool = True, trace_capacity: int = 0, local_bound: bool = False,
                 kappa=None, boomerang=None, reversible: bool = False, strong_upperbounds: bool = False, refresh=None, stickyzz: bool = False):
        self.problem = problem
        self.d = problem.d
        self._h = C.c_void_p()
        flags = (0 if record_trace else _capi.ZZB_FLAG_NO_TRACE) | (_capi.ZZB_FLAG_LOCAL_BOUND if local_bound else 0) \
            | (_capi.ZZB_FLAG_STICKY if kappa is not None else 0) | (_capi.ZZB_FLAG_BOOMERANG if boomerang is not None else 0) \
            | (_capi.ZZB_FLAG_STICKY_REVERSIBLE if (kappa is not None and reversible) else 0) \
            | (_capi.ZZB_FLAG_STICKY_STRONG_UB if (kappa is not None and strong_upperbounds) else 0) \
            | (_capi.ZZB_FLAG_STICKY_ZZ if (kappa is not None and stickyzz) else 0) \
            | (_capi.ZZB_FLAG_REFRESH if refresh is not None else 0)
        self.record_trace = record_trace
        check(_capi.lib().zzb_run_create(problem._h, flags, int(trace_capacity), C.byref(self._h)))
        if kappa is not None:
            self._kappa = f8(kappa)
            check(_capi.lib().zzb_run_upload_kappa(self._h, ptr(self._kappa)))
        if refresh is not None:     # a ZigZag with lambdaref > 0: velocity refreshments theta_i <- sigma_i * (+-1) (src/sfact.jl:78-114)
            self._rsigma = f8(refresh.sigma)
            check(_capi.lib().zzb_run_upload_refresh(self._h, ptr(self._rsigma), float(refresh.lambdaref)))
        if boomerang is not None:   # a FactBoomerang: sigma, lambdaref, rho (its Gamma / mu are the problem's sampler matrices)
            self._sigma = f8(boomerang.sigma)
            check(_capi.lib().zzb_run_upload_boomerang(self._h, ptr(self._sigma), boomerang.lambdaref, boomerang.rho))

    def set(self, **kw):
        for k, v in kw.items():
            check(_capi.lib().zzb_run_set(self._h, k.encode(), float(v)))
        return self

    def upload(self, t0, x0, theta0, c, seed=(1, 2), adapt=False, factor=1.8):
        self._x0, self._th0, self._c = f8(x0), f8(theta0), f8(c)
        sd = np.array(seed, dtype=np.uint64)
        check(_capi.lib().zzb_run_upload(self._h, float(t0), ptr(self._x0), ptr(self._th0), ptr(self._c), ptr(sd),
                                         int(bool(adapt)), float(factor)))
        self.t0 = float(t0)
        return self

    def discretize(self, dt: float, n_rows: int):
        """Ask for ``x(t0 + k dt)``, ``k = 0 .. n_rows-1`` to be produced on the device while the run is committed
        (``collect(discretize(trace, dt))``, src/trace.jl:94-125, without handing the trace back).  Before :meth:`upload`."""
        check(_capi.lib().zzb_run_discretize(self._h, float(dt), int(n_rows)))
        self._grid = (float(dt), int(n_rows))
        return self

    def grid(self):
        """``(ts, xs)`` of the device-side discretisation, cut to the rows at or before the simulated frontier."""
        dt, n = self._grid
        xs = np.empty((n, self.d))
        valid = C.c_int64()
        check(_capi.lib().zzb_run_grid(self._h, ptr(xs), 0, n, C.byref(valid)))
        k = valid.value
        return self.t0 + dt * np.arange(k), xs[:k]

    # ---- sharding over the GPUs of one node (see multigpu.py) -------------------------------------------------
    def shard(self, rank: int, nranks: int):
        check(_capi.lib().zzb_run_shard(self._h, int(rank), int(nranks)))
        return self

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(8 * 64)
        n = C.c_int64()
        check(_capi.lib().zzb_run_ipc_export(self._h, buf, len(buf), C.byref(n)))
        return buf.raw[: n.value]

    def ipc_import(self, peer_rank: int, blob: bytes):
        check(_capi.lib().zzb_run_ipc_import(self._h, int(peer_rank), blob, len(blob)))

    def owned_range(self):
        lo, hi = C.c_int64(), C.c_int64()
        check(_capi.lib().zzb_run_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def trace_filter(self, J):
        """``subtrace`` at the source (src/trace.jl:275-290): record only the events of the coordinates `J` (1-based, strictly
        ascending), renumbered to their position in `J`.  Before :meth:`upload`.  ``None`` / empty lifts the filter."""
        J = np.ascontiguousarray([] if J is None else J, dtype=np.int64)
        check(_capi.lib().zzb_run_trace_filter(self._h, ptr(J), len(J)))
        return self

    def cummean(self):
        """``cummean(trace)`` (src/trace.jl:203-225) of this run, computed on the device while the trace is still in HBM:
        ``(offsets, times, values)`` -- the running time averages of coordinate ``k`` (0-based) after each of its events are
        ``values[offsets[k]:offsets[k+1]]`` at ``times[...]``; :func:`cummean_lists` turns this into the reference's list of pairs
        (with the leading ``(t0, x0_k)``)."""
        n = C.c_int64()
        check(_capi.lib().zzb_trace_len(self._h, C.byref(n)))
        off = np.empty(self.d + 1, dtype=np.int64)
        times, values = np.empty(max(n.value, 1)), np.empty(max(n.value, 1))
        check(_capi.lib().zzb_trace_cummean(self._h, ptr(off), ptr(times), ptr(values)))
        return off, times[:n.value], values[:n.value]

    def inclusion_prob(self):
        """``inclusion_prob(trace)`` (src/trace.jl:161-178) of a sticky run, from a device accumulator (no trace needed)."""
        p = np.empty(self.d)
        check(_capi.lib().zzb_trace_inclusion(self._h, ptr(p)))
        return p

    def reset(self):
        """Re-initialise the device state from the inputs already resident in HBM (no host traffic)."""
        check(_capi.lib().zzb_run_reset(self._h))
        return self

    def execute(self, T) -> float:
        """Run to T; returns the CUDA-event time of the event-loop kernel(s) in milliseconds."""
        ms = C.c_float()
        st = _capi.lib().zzb_run_execute(self._h, float(T), C.byref(ms))
        self.device_ms = ms.value
        check(st)
        return ms.value

    def fetch_into(self, *, t=None, x=None, theta=None, c=None, acc=None, s1=None, s2=None):
        """Read results straight into caller-provided arrays (float64 / int64 of length d; pinned memory copies at
        PCIe speed).  Returns (num, nacc)."""
        num, nacc = C.c_int64(), C.c_int64()
        check(_capi.lib().zzb_run_fetch(self._h, ptr(t), ptr(x), ptr(theta), ptr(c), ptr(acc), ptr(s1), ptr(s2),
                                        C.byref(num), C.byref(nacc)))
        return num.value, nacc.value

    def counts(self):
        acc = np.empty(self.d, np.int64)
        num = C.c_int64()
        check(_capi.lib().zzb_run_counts(self._h, ptr(acc), C.byref(num)))
        return acc, num.value

    def final_state(self):
        t, x, th, c = (np.empty(self.d) for _ in range(4))
        check(_capi.lib().zzb_run_final_state(self._h, ptr(t), ptr(x), ptr(th), ptr(c)))
        return t, x, th, c

    def events(self, out=None):
        """The trace so far, ordered by (time, coordinate).  `out`: an optional preallocated EVENT_DTYPE array (e.g. in pinned
        memory: the events ordered on the device are then copied from HBM at PCIe speed); a view of its first n entries is returned."""
        n = C.c_int64()
        check(_capi.lib().zzb_trace_len(self._h, C.byref(n)))
        if out is not None and len(out) >= n.value:
            ev = out[: n.value]
        else:
            ev = np.empty(n.value, dtype=EVENT_DTYPE)
        if n.value:
            check(_capi.lib().zzb_trace_copy(self._h, ptr(ev), 0, n.value))
        return ev

    def clear_events(self):
        check(_capi.lib().zzb_trace_clear(self._h))

    def n_events(self) -> int:
        n = C.c_int64()
        check(_capi.lib().zzb_trace_len(self._h, C.byref(n)))
        return n.value

    def moments(self):
        m1, m2 = np.empty(self.d), np.empty(self.d)
        check(_capi.lib().zzb_trace_moments(self._h, ptr(m1), ptr(m2)))
        return m1, m2

    def sums(self):
        s1, s2 = np.empty(self.d), np.empty(self.d)
        check(_capi.lib().zzb_trace_sums(self._h, ptr(s1), ptr(s2)))
        return s1, s2

    def stats(self):
        out = np.zeros(24, np.int64)
        check(_capi.lib().zzb_run_stats(self._h, ptr(out), 24))
        keys = ("windows", "retries", "passes", "node_evals", "rebases", "launches", "grid", "block",
                "ns_scan", "ns_relax", "ns_tail", "ns_commit", "ns_barrier", "ns_phaseb", "n_barriers", "n_tail_passes") + tuple("dbg%d" % q for q in range(8))
        return dict(zip(keys, (int(v) for v in out)))

    def close(self):
        if self._h:
            _capi.lib().zzb_run_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_problem(grad, F):
    if isinstance(grad, Problem):
        return grad, False
    if not isinstance(grad, (GaussianPotential, LogisticSubsampled)):
        raise TypeError("the B200 path needs a target descriptor (GaussianPotential / LogisticSubsampled), not a closure: "
                        "a device kernel cannot call back into the host")
    if isinstance(grad, LogisticSubsampled) and not isinstance(F, ZigZag):
        raise TypeError("the logistic target runs with ZigZag dynamics only")
    if not isinstance(F, (ZigZag, FactBoomerang)):
        raise TypeError("only ZigZag and FactBoomerang dynamics are implemented on the device path")
    if isinstance(F, ZigZag) and F.lambdaref != 0.0 and isinstance(grad, LogisticSubsampled):
        raise NotImplementedError("refreshments (lambdaref > 0) with the logistic target are not implemented on the device path")
    return Problem(grad, F), True


def spdmp(grad, t0, x0, theta0, T, c, *rest, factor=1.8, adapt=False, seed=None, record_trace=True, tune=None,
          discretize_dt=None, trace_filter=None):
    """``spdmp(grad, t0, x0, theta0, T, c, [G,] F, args...; factor=1.8, adapt=false, seed=Seed())``
    = ``Xi, (t, x, theta), (acc, num), c`` (src/sfact.jl:162-214).

    `G` may be ``Matched()`` / ``All()``; both give the same event law (the reference only moves different
    coordinate sets eagerly), so they share one kernel.  Trailing `args` (the reference forwards them to the
    closure) are accepted and ignored.  ``discretize_dt=dt`` additionally produces ``collect(discretize(Xi, dt))`` on the
    device (``Xi.grid = (ts, xs)`` up to ``T``), so that ``record_trace=False`` runs still return the thinned path.  Raises :class:`BoundError` with the reference's message when a proposal is
    accepted with ``l >= lb`` and ``adapt`` is false (src/sfact.jl:124).
    """
    rest = list(rest)
    if rest and isinstance(rest[0], (All, Matched)):
        rest.pop(0)
    if not rest:
        raise TypeError("spdmp: missing sampler F")
    F = rest.pop(0)
    local_bound = isinstance(c, LocalBound)   # spdmp(grad, t0, x0, th0, T, C::LocalBound, F, ...) src/local.jl:95,148
    if local_bound and isinstance(grad, LogisticSubsampled):
        raise NotImplementedError("LocalBound with the logistic target is not implemented on the device path")
    if local_bound:
        c = c.c
        if isinstance(grad, GaussianPotential):  # the bound comes from the target: F.Gamma / F.mu are ignored (local.jl:2-6)
            F = ZigZag(grad.Gamma, np.zeros(grad.Gamma.n), F.sigma, lambdaref=F.lambdaref, rho=F.rho)
    prob, own = _as_problem(grad, F)
    if seed is None:  # Seed() = fresh entropy (src/ZigZagBoomerang.jl:10)
        seed = (secrets.randbits(64), secrets.randbits(64))
    boom = F if isinstance(F, FactBoomerang) else None
    if boom is not None and local_bound:
        raise NotImplementedError("LocalBound with FactBoomerang is not implemented on the device path")
    refr = F if (isinstance(F, ZigZag) and F.lambdaref != 0.0) else None   # hasrefresh(Z) (src/fact_samplers.jl:19)
    if refr is not None and local_bound:
        raise NotImplementedError("LocalBound with ZigZag refreshments is not implemented on the device path")
    run = Run(prob, record_trace=record_trace, local_bound=local_bound, boomerang=boom, refresh=refr)
    try:
        if tune:
            run.set(**tune)
        if discretize_dt is not None:
            run.discretize(discretize_dt, int(np.floor((T - t0) / discretize_dt)) + 1)
        if trace_filter is not None:   # the returned trace is subtrace(Xi, J), produced on the device (src/trace.jl:275-290)
            run.trace_filter(trace_filter)
        run.upload(t0, x0, theta0, c, seed=seed, adapt=adapt, factor=factor)
        run.execute(T)
        t, x, th, cc = run.final_state()
        acc, num = run.counts()
        ev = run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE)
        if trace_filter is not None:
            Jm = np.asarray(trace_filter, dtype=np.int64) - 1
            Xi = FactTrace(F, t0, f8(x0)[Jm], f8(theta0)[Jm], ev)
        else:
            Xi = FactTrace(F, t0, x0, theta0, ev)
        Xi.grid = run.grid() if discretize_dt is not None else None
        Xi.moments = run.moments() if (num and boom is None) else None   # event-based moments assume linear segments
        Xi.stats = run.stats()
        Xi.device_ms = run.device_ms
        return Xi, (t, x, th), (acc, num), (LocalBound(cc) if local_bound else cc)
    finally:
        run.close()
        if own:
            prob.close()


def pdmp(grad, t0, x0, theta0, T, c, F, *args, **kw):
    """``pdmp(grad, t0, x0, theta0, T, c, F::ZigZag, args...)`` = ``spdmp(..., All(), F, ...)`` (src/sfact.jl:236)."""
    return spdmp(grad, t0, x0, theta0, T, c, All(), F, *args, **kw)


def sspdmp(grad, t0, x0, theta0, T, c, *rest, seed=None, record_trace=True, tune=None, reversible=False, strong_upperbounds=False,
           adapt=False, factor=1.5, **unsupported):
    """``sspdmp(grad, t0, x0, theta0, T, c, [G,] F::ZigZag, kappa, args...; reversible=false, strong_upperbounds=false)`` =
    ``Xi, (t, x, theta), (acc, num), c`` (src/ss_fact.jl:159-217): sticky ZigZag -- coordinates freeze when they hit 0 and thaw
    after an Exp(kappa_i) time.  ``reversible``: a thawing coordinate re-enters with a random sign (:111-113);
    ``strong_upperbounds``: a freeze reschedules nobody (:97-107).  `acc` is the number of accepted reflections (a scalar, like
    the reference).  ``adapt=true``: an accepted proposal with ``l > lb`` multiplies ``c[i]`` by ``factor`` instead of raising
    (:132-136); the reference then also RESETS its diagnostic counters acc and num (:134) -- the device returns the totals."""
    rest = list(rest)
    if rest and (rest[0] is None or isinstance(rest[0], (All, Matched))):
        rest.pop(0)
    F, kappa = rest[0], rest[1]
    if not isinstance(F, ZigZag):   # the reference's method is sspdmp(..., F::ZigZag, kappa, ...) (src/ss_fact.jl:159)
        raise TypeError("sspdmp: the sticky sampler is defined for F::ZigZag only (src/ss_fact.jl:159), got %s" % type(F).__name__)
    prob, own = _as_problem(grad, F)
    if seed is None:
        seed = (secrets.randbits(64), secrets.randbits(64))
    run = Run(prob, record_trace=record_trace, kappa=np.broadcast_to(f8(kappa), (prob.d,)).copy(), reversible=bool(reversible),
              strong_upperbounds=bool(strong_upperbounds))
    try:
        if tune:
            run.set(**tune)
        run.upload(t0, x0, theta0, c, seed=seed, adapt=bool(adapt), factor=float(factor))
        run.execute(T)
        t, x, th, cc = run.final_state()
        acc, num = run.counts()
        ev = run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE)
        Xi = FactTrace(F, t0, x0, theta0, ev)
        Xi.stats = run.stats()
        Xi.acc_per_coordinate = acc
        Xi.sums = run.sums()
        Xi.inclusion_prob = run.inclusion_prob() if num else None   # src/trace.jl:161-178 from the device accumulator
        return Xi, (t, x, th), (int(acc.sum()), num), cc
    finally:
        run.close()
        if own:
            prob.close()


def sspdmp3(grad, u0, T, c, G, Z, kappa, *args, rule="reversible", adapt=False, factor=1.5, seed=None, record_trace=True, tune=None,
            clusteralpha=1.0):
    """``sspdmp3(grad, u0, T, c, nothing, Z, kappa, args...; adapt=false, factor=1.5, rule=:reversible, clusterα=1.0)`` =
    ``trace, acc, uT`` (src/sparsestickyzz.jl:405-422): the strong-bound sparse sticky ZigZag -- one bound constant ``c`` valid
    for ``1/c`` and then renewed (:136-142), a reflection reschedules nobody else, coordinates stick at 0 and thaw at rate
    ``kappa`` each.  ``u0 = (x0, theta0)``: coordinates with ``x0 == 0`` start frozen (``sparsestickystate``, :10-12; the
    reference draws the velocities of the others from the global RNG -- here they are an argument).  Returns
    ``trace, (acc, num), (t, x, theta)``; `acc` = accepted reflections.  Not on the device path: ``adapt``, ``clusterα < 1``."""
    if adapt:
        raise NotImplementedError("sspdmp3(...; adapt=true) is not implemented on the device path")
    if clusteralpha != 1.0:
        raise NotImplementedError("sspdmp3(...; clusterα < 1) is not implemented on the device path")
    if not isinstance(Z, ZigZag):
        raise TypeError("sspdmp3: Z must be a ZigZag (its Gamma gives the dependency structure, src/sparsestickyzz.jl:407)")
    x0, th0 = u0
    prob, own = _as_problem(grad, Z)
    if seed is None:
        seed = (secrets.randbits(64), secrets.randbits(64))
    d = prob.d
    run = Run(prob, record_trace=record_trace, kappa=np.full(d, float(kappa)))
    try:
        run.set(strong_c=float(c), strong_rule={"sticky": 0, "reversible": 1}[rule], **(tune or {}))
        run.upload(0.0, x0, th0, np.full(d, float(c)), seed=seed)
        run.execute(T)
        t, x, th, _ = run.final_state()
        acc, num = run.counts()
        ev = run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE)
        Xi = FactTrace(Z, 0.0, f8(x0), np.where(f8(x0) != 0.0, f8(th0), 0.0), ev)
        Xi.stats = run.stats()
        Xi.device_ms = run.device_ms
        Xi.acc_per_coordinate = acc
        return Xi, (int(acc.sum()), num), (t, x, th)
    finally:
        run.close()
        if own:
            prob.close()


def sspdmp2(grad, t0, x0, v0, T, c, G, Z, kappa, *args, strong_upperbounds=False, adapt=False, factor=1.5, seed=None, record_trace=True,
            tune=None):  # noqa: D401
    """``sspdmp2(grad, t, x0, v0, T, c, nothing, Z, kappa, args...; strong_upperbounds=false, adapt=false, factor=1.5)`` =
    ``trace, acc`` (src/stickyzz.jl:322-338): the dense sticky ZigZag ``stickyzz`` -- the loop of ``sspdmp`` (affine bounds, a
    reflection reschedules its neighbourhood, coordinates stick at 0 and thaw at rate ``kappa[i]`` with the velocity they had)
    with proposal times drawn at rate ``0.01 + (a + b t)^+`` (``queue_time!``, :144-165) and coordinates that start at 0 starting
    frozen (:198-206).  ``adapt=true`` multiplies ``c[i]`` by ``factor`` when a proposal is accepted with ``l > lb`` (:305-309).
    Returns ``trace, (acc, num)``; ``trace.final`` holds ``(t, x, theta)``, ``trace.c`` the (adapted) bounds."""
    if not isinstance(Z, ZigZag):
        raise TypeError("sspdmp2: Z must be a ZigZag (src/stickyzz.jl:323-331)")
    prob, own = _as_problem(grad, Z)
    if seed is None:
        seed = (secrets.randbits(64), secrets.randbits(64))
    d = prob.d
    kv = np.full(d, float(kappa)) if np.isscalar(kappa) else f8(kappa)
    run = Run(prob, record_trace=record_trace, kappa=kv, strong_upperbounds=strong_upperbounds, stickyzz=True)
    try:
        if tune:
            run.set(**tune)
        run.upload(t0, x0, v0, c, seed=seed, adapt=bool(adapt), factor=float(factor))
        run.execute(T)
        t, x, th, cc = run.final_state()
        acc, num = run.counts()
        ev = run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE)
        Xi = FactTrace(Z, t0, f8(x0), np.where(f8(x0) != 0.0, f8(v0), 0.0), ev)
        Xi.stats = run.stats()
        Xi.device_ms = run.device_ms
        Xi.acc_per_coordinate = acc
        Xi.final = (t, x, th)
        Xi.c = cc
        return Xi, (int(acc.sum()), num)
    finally:
        run.close()
        if own:
            prob.close()


def sspdmp4(coloring, grad, t0, x0, v0, T, c, G, Z, kappa, *args, adapt=False, factor=1.5, seed=None, record_trace=True, tune=None):
    """``sspdmp4(Coloring, grad, t, x0, v0, T, c, nothing, Z, kappa, args...; adapt=false, factor=1.5)`` = ``trace, acc``
    (src/asynchzz.jl:250-265): the strong-bound sticky ZigZag of ``asynchzz`` -- a bound constant ``c[i]`` per coordinate, valid
    for ``1/c[i]`` and then renewed (``StrongUpperBounds``, :2-7,20-28), a reflection reschedules nobody else, coordinates stick at
    0 and thaw at rate ``kappa[i]``, continuing with the velocity they had (rule ``:sticky``, :206-213); coordinates with
    ``x0 == 0`` start frozen (:112-116).  The reference runs this process with its own thread schedule (local minima of a
    ``PartialQueue`` over regions coloured by `Coloring`, :150-245); the device uses its windowed relaxation, so `coloring` is
    accepted and ignored -- the law does not depend on the schedule.  Returns ``trace, (acc, num)``; ``trace.final`` holds
    ``(t, x, theta)``.  Not on the device path: ``adapt``."""
    if adapt:
        raise NotImplementedError("sspdmp4(...; adapt=true) is not implemented on the device path")
    if not isinstance(Z, ZigZag):
        raise TypeError("sspdmp4: Z must be a ZigZag (its Gamma gives the dependency structure, src/asynchzz.jl:252-256)")
    prob, own = _as_problem(grad, Z)
    if seed is None:
        seed = (secrets.randbits(64), secrets.randbits(64))
    d = prob.d
    cv = np.full(d, float(c)) if np.isscalar(c) else f8(c)
    kv = np.full(d, float(kappa)) if np.isscalar(kappa) else f8(kappa)
    if not ((cv > 0).all() and (kv > 0).all()):
        raise ValueError("sspdmp4 needs c[i] > 0 and kappa[i] > 0")
    run = Run(prob, record_trace=record_trace, kappa=kv)
    try:
        run.set(strong_c=float(cv[0]), strong_rule=2, **(tune or {}))
        run.upload(t0, x0, v0, cv, seed=seed)
        run.execute(T)
        t, x, th, _ = run.final_state()
        acc, num = run.counts()
        ev = run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE)
        Xi = FactTrace(Z, t0, f8(x0), np.where(f8(x0) != 0.0, f8(v0), 0.0), ev)
        Xi.stats = run.stats()
        Xi.device_ms = run.device_ms
        Xi.acc_per_coordinate = acc
        Xi.final = (t, x, th)
        return Xi, (int(acc.sum()), num)
    finally:
        run.close()
        if own:
            prob.close()


class FactSampler:
    """``FactSampler(grad, u0, c, [G,] F; factor=1.8, adapt=false, seed)`` with ``u0 = (t0, (x0, theta0))`` -- the pull-style
    interface of src/sfactiter.jl:5-64.  Iterating yields ``(t, (t, i, x_i, theta_i))`` pairs, one per accepted event, in
    time order, for ever; the device simulates `windows_per_pull` windows ahead at a time and hands the events over."""

    def __init__(self, grad, u0, c, *rest, factor=1.8, adapt=False, seed=None, windows_per_pull=16):
        rest = list(rest)
        if rest and (rest[0] is None or isinstance(rest[0], (All, Matched))):
            rest.pop(0)
        self.F = rest[0]
        self.grad, self.u0, self.c = grad, u0, c
        self.factor, self.adapt = factor, adapt
        self.seed = seed if seed is not None else (secrets.randbits(64), secrets.randbits(64))
        self.windows_per_pull = int(windows_per_pull)

    def _open(self):
        """Problem and run for this sampler: the same dispatch as spdmp (FactBoomerang dynamics, LocalBound bounds)."""
        F = self.F
        local_bound = isinstance(self.c, LocalBound)
        if local_bound and isinstance(self.grad, LogisticSubsampled):
            raise NotImplementedError("LocalBound with the logistic target is not implemented on the device path")
        if local_bound and isinstance(F, FactBoomerang):
            raise NotImplementedError("LocalBound with FactBoomerang is not implemented on the device path")
        if local_bound and isinstance(self.grad, GaussianPotential):   # the bound comes from the target (src/local.jl:2-6)
            F = ZigZag(self.grad.Gamma, np.zeros(self.grad.Gamma.n), F.sigma, lambdaref=F.lambdaref, rho=F.rho)
        prob, own = _as_problem(self.grad, F)
        run = Run(prob, record_trace=True, local_bound=local_bound, boomerang=F if isinstance(F, FactBoomerang) else None,
                  refresh=F if (isinstance(F, ZigZag) and F.lambdaref != 0.0) else None)
        return prob, own, run, local_bound

    def chunks(self):
        """The same stream as iteration, but handed over as numpy record arrays (EVENT_DTYPE, time-ordered), one per pull of
        `windows_per_pull` windows: no Python object per event."""
        t0, (x0, th0) = self.u0
        prob, own, run, local_bound = self._open()
        try:
            run.set(max_windows=self.windows_per_pull)
            run.upload(t0, x0, th0, self.c.c if local_bound else self.c, seed=self.seed, adapt=self.adapt, factor=self.factor)
            while True:
                run.execute(np.inf)          # a bounded slice of windows; the controller state persists on the device
                ev = run.events()
                run.clear_events()
                yield ev
        finally:
            run.close()
            if own:
                prob.close()

    def __iter__(self):
        for ev in self.chunks():
            for t, i, x, th in zip(ev["t"].tolist(), ev["i"].tolist(), ev["x"].tolist(), ev["theta"].tolist()):
                yield t, (t, i, x, th)


def trace(FS: FactSampler, T):
    """``trace(FS, T)`` (src/sfactiter.jl:66-79), including its quirk: the first event of the iteration is consumed and
    not stored; events with t > T end the collection."""
    t0, (x0, th0) = FS.u0
    parts, first = [], True
    gen = FS.chunks()
    for ev in gen:
        if first and len(ev):
            ev, first = ev[1:], False          # the first event is consumed, not stored (sfactiter.jl:68-70)
        k = int(np.searchsorted(ev["t"], T, side="right"))   # events are time-ordered: stop at the first one with t > T
        parts.append(ev[:k])
        if k < len(ev):
            break
    gen.close()
    return FactTrace(FS.F, t0, x0, th0, np.concatenate(parts) if parts else np.empty(0, dtype=EVENT_DTYPE))


def cummean(tr: FactTrace):
    """``cummean(trace)`` (src/trace.jl:203-225): per coordinate the running time-average at each of its events, as a list of
    ``(times, values)`` pairs starting with ``(t0, x0_i)``."""
    ev = tr.events
    d = len(tr.x0)
    i, t_prev, x_prev = _previous_per_coordinate(tr)
    term = (x_prev + ev["x"]) * (ev["t"] - t_prev)
    order = np.argsort(i, kind="stable")
    si = i[order]
    bounds = np.searchsorted(si, np.arange(d + 1))
    out = []
    for k in range(d):
        sel = order[bounds[k]:bounds[k + 1]]
        tt = ev["t"][sel]
        run = np.cumsum(term[sel])              # sequential partial sums, as the reference's y[i] += ...
        out.append((np.r_[tr.t0, tt], np.r_[tr.x0[k], run / (2 * tt)]))
    return out


def cummean_lists(t0, x0, offsets, times, values):
    """The reference's shape of ``cummean`` from the CSR triple of :meth:`Run.cummean`: per coordinate ``(times, values)`` starting
    with ``(t0, x0_k)`` (views and one concatenation per coordinate, no arithmetic)."""
    return [(np.r_[t0, times[offsets[k]:offsets[k + 1]]], np.r_[x0[k], values[offsets[k]:offsets[k + 1]]]) for k in range(len(x0))]


def inclusion_prob(tr: FactTrace):
    """``inclusion_prob(trace)`` (src/trace.jl:161-178): fraction of time a coordinate is away from 0 (sticky samplers).
    Julia parses ``x[i] != 0 | xi != 0`` as ``x[i] != (0 | xi) != 0``; the evident intent (either end non-zero) is used."""
    ev = tr.events
    y = np.zeros(len(tr.x0))
    if not len(ev):
        return y
    T = ev["t"][-1]
    i, t_prev, x_prev = _previous_per_coordinate(tr)
    np.add.at(y, i, ((x_prev != 0) | (ev["x"] != 0)).astype(np.float64) * (ev["t"] - t_prev) / T)
    return y
