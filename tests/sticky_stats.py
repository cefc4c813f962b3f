"""Occupancy statistics of sticky traces (shared by the CPU and GPU law tests of config 4)."""
import numpy as np


def chain_precision(zzb, p):
    """Gamma of test/sparsesticky.jl:21-23: tridiag(-1, 2, -1) with corners 1, plus 0.1 I."""
    import scipy.sparse as sp
    M = sp.diags([-np.ones(p - 1), 2 * np.ones(p), -np.ones(p - 1)], [-1, 0, 1]).tolil()
    M[0, 0] = 1
    M[p - 1, p - 1] = 1
    return zzb.CSC.from_scipy((M + 0.1 * sp.eye(p)).tocsc())


def occupancy(ev, d, x0, T):
    """Per coordinate: fraction of [0, T] spent away from 0 (velocity != 0 after the coordinate's latest event) and the time
    average of x^2 over the piecewise linear path between its own events (frozen stretches contribute 0)."""
    order = np.lexsort((ev["t"], ev["i"]))
    e = ev[order]
    e = e[e["t"] <= T]
    idx = np.searchsorted(e["i"], np.arange(1, d + 2))
    occ, m2 = np.zeros(d), np.zeros(d)
    for j in range(d):
        seg = e[idx[j]:idx[j + 1]]
        t = np.concatenate([[0.0], seg["t"]])
        x = np.concatenate([[x0[j]], seg["x"]])
        moving = np.concatenate([[x0[j] != 0], seg["theta"] != 0])
        dur = np.diff(np.concatenate([t, [T]]))
        occ[j] = (dur * moving).sum() / T
        a, b, dt = x[:-1], x[1:], np.diff(t)
        m2[j] = (dt * (a * a + a * b + b * b)).sum() / (3 * T)
    return occ, m2
