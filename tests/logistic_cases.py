"""Inputs of the subsampled logistic-regression cases (config 3 of BASELINE.json, scripts/logistic.jl) shared by the CPU
tests (oracle / host emulation of the schedule) and the GPU parity tests."""
import numpy as np

# (levels, continuous regressors, rows per column m, generator seed, T): small designs in which every column is non-empty
SMALL = [((3, 3), 2, 8, 1, 20.0), ((4, 4), 2, 10, 3, 30.0), ((5, 4), 1, 12, 7, 10.0), ((2, 3, 2), 2, 10, 5, 15.0)]
FULL = ((20, 20), 2, 20, 2)   # README / BASELINE: n = 8840, p = 442


def make(zzb, levels, r, m, seed, k=10):
    cfg = zzb.logistic_config(levels=levels, r=r, m=m, seed=seed)
    assert np.diff(cfg["A"].colptr).min() >= 1
    cfg["logistic"] = dict(A=cfg["A"], At=cfg["At"], y=cfg["y"], ny=cfg["ny"], mu=cfg["mu"], gamma0=cfg["gamma0"], k=k)
    return cfg


def replicas(zzb, cfg, R):
    out = zzb.replicate_logistic(cfg, R)
    lg = cfg["logistic"]
    out["logistic"] = dict(A=out["A"], At=out["At"], y=out["y"], ny=out["ny"], mu=out["mu"], gamma0=lg["gamma0"], k=lg["k"])
    return out


def run_oracle(O, cfg, T, *, seed=(5, 6), adapt=True, factor=5.0, mode=None, c=None):
    """spdmp(grad_phi_moving, t0, x0, th0, T, c, Zdrop, SelfMoving(), A, At, mu, y, ny, k; adapt, factor) (scripts/logistic.jl:167)"""
    return O.spdmp(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], T, cfg["c"] if c is None else c, mu=cfg["mu"],
                   adapt=adapt, factor=factor, logistic=cfg["logistic"], seed=seed, mode=O.PARITY_MODE if mode is None else mode)


class R:
    pass


def run_device(zzb, cfg, T, *, seed=(5, 6), adapt=True, factor=5.0, tune=None, c=None):
    lg = cfg["logistic"]
    grad = zzb.LogisticSubsampled(lg["A"], lg["At"], lg["y"], lg["ny"], lg["mu"], lg["gamma0"], lg["k"])
    Z = zzb.ZigZag(cfg["Gamma_drop"], cfg["mu"], cfg["sigma"], rho=0.5, lambdaref=0.0)
    Xi, (t, x, th), (acc, num), cc = zzb.spdmp(grad, 0.0, cfg["x0"], cfg["theta0"], T, cfg["c"] if c is None else c, Z, zzb.SelfMoving(),
                                               lg["A"], lg["At"], lg["mu"], lg["y"], lg["ny"], lg["k"], adapt=adapt, factor=factor,
                                               seed=seed, tune=tune)
    r = R()
    r.events, r.t, r.x, r.theta, r.c, r.acc, r.num = Xi.events, t, x, th, cc, acc, num
    r.s1, r.s2 = None, None
    del r.s1, r.s2
    r.stats, r.device_ms = Xi.stats, Xi.device_ms
    return r, Xi
