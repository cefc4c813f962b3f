"""Inputs of the committed regression fixtures under tests/golden/ (written by tests/golden/make_golden.py from the CPU
oracle, mode ctr|lazy).  Shared by the generator, the CPU test (oracle vs fixture) and the GPU test (CUDA path vs fixture --
no oracle involved).  They pin OUR RNG / arithmetic contract; the reference holds no golden vector for this path."""
import numpy as np


def summary(events, num, acc):
    t = np.ascontiguousarray(events["t"]); x = np.ascontiguousarray(events["x"]); th = np.ascontiguousarray(events["theta"])
    xr = lambda a: int(np.bitwise_xor.reduce(a.view(np.uint64))) if len(a) else 0
    return dict(num=int(num), n_events=int(len(events)), acc_sum=int(np.sum(acc)), first_i=events["i"][:32].tolist(),
                first_t_hex=[float(v).hex() for v in t[:8]], last_t_hex=float(t[-1]).hex(),
                xor_t=xr(t), xor_x=xr(x), xor_theta=xr(th))


def case_inputs(zzb, name):
    """-> dict(G, Zg, x0, th0, c, T, seed, kind, extra)"""
    if name == "gmrf16_T3":          # config 2 in small: lattice GMRF, c = column norms
        G, x0, th0, c = zzb.gmrf_config(16)
        return dict(G=G, Zg=G, x0=x0, th0=th0, c=c, T=3.0, seed=(1, 2), kind="zigzag")
    if name == "gmrf12_tight":       # c = sqrt(eps) (scripts/example.jl:39)
        G, x0, th0, c = zzb.gmrf_config(12, tight=True)
        return dict(G=G, Zg=G, x0=x0, th0=th0, c=c, T=4.0, seed=(3, 4), kind="zigzag")
    if name == "spd8_adapt":         # test/maintest.jl:4-34 shapes: Z = ZigZag(0.9 Gamma, 0), adaptation of c
        G = zzb.random_spd(8, seed=2)
        rng = np.random.default_rng(5)
        return dict(G=G, Zg=G.scaled(0.9), x0=rng.random(8), th0=rng.choice(np.array([-1.0, 1.0]), 8), c=0.05 * G.colnorms(),
                    T=60.0, seed=(11, 12), kind="zigzag", adapt=True)
    if name == "localbound16":       # src/local.jl
        G, x0, th0, _ = zzb.gmrf_config(16)
        return dict(G=G, Zg=G, x0=x0, th0=th0, c=np.full(G.n, 0.5), T=3.0, seed=(1, 2), kind="localbound", adapt=True)
    if name == "sticky12":           # src/ss_fact.jl
        G, x0, th0, c = zzb.gmrf_config(12)
        return dict(G=G, Zg=G, x0=x0, th0=th0, c=c, T=5.0, seed=(1, 2), kind="sticky", kappa=np.full(G.n, 0.7))
    if name == "boomerang12":        # F::FactBoomerang in src/sfact.jl
        G = zzb.grid_precision(12, 12)
        rng = np.random.default_rng(9)
        diag = G.to_scipy().diagonal()
        return dict(G=G, Zg=G, x0=rng.standard_normal(G.n), th0=rng.standard_normal(G.n) / np.sqrt(diag), c=G.colnorms(), T=6.0,
                    seed=(2, 3), kind="boomerang", boom=(diag ** -0.5, 5.0, 0.2))
    if name == "logistic26":         # scripts/logistic.jl in small: subsampled logistic regression, adapt = true, factor = 5
        import logistic_cases as LC
        return dict(kind="logistic", cfg=LC.make(zzb, (4, 4), 2, 10, 3), T=30.0, seed=(5, 6))
    raise KeyError(name)


CASES = ["gmrf16_T3", "gmrf12_tight", "spd8_adapt", "localbound16", "sticky12", "boomerang12"]
LOGISTIC_CASES = ["logistic26"]   # checked by tests/test_logistic.py (oracle) and tests/test_gpu_zz_logistic.py (device)


def run_oracle(O, k):
    if k["kind"] == "logistic":
        import logistic_cases as LC
        r = LC.run_oracle(O, k["cfg"], k["T"], seed=k["seed"])
        return summary(r.events, r.num, r.acc)
    mode = O.PARITY_MODE | (O.LOCAL_BOUND if k["kind"] == "localbound" else 0)
    r = O.spdmp(k["G"], k["Zg"], 0.0, k["x0"], k["th0"], k["T"], k["c"], seed=k["seed"], mode=mode, adapt=k.get("adapt", False),
                kappa=k.get("kappa"), boom=k.get("boom"))
    return summary(r.events, r.num, r.acc)


def run_device(zzb, k):
    if k["kind"] == "logistic":
        import logistic_cases as LC
        r, _ = LC.run_device(zzb, k["cfg"], k["T"], seed=k["seed"])
        return summary(r.events, r.num, r.acc)
    n = k["G"].n
    tgt = zzb.GaussianPotential(k["G"])
    if k["kind"] == "sticky":
        Xi, _, (acc, num), _ = zzb.sspdmp(tgt, 0.0, k["x0"], k["th0"], k["T"], k["c"], zzb.ZigZag(k["Zg"], np.zeros(n)), k["kappa"], seed=k["seed"])
    elif k["kind"] == "boomerang":
        sigma, lref, rho = k["boom"]
        Xi, _, (acc, num), _ = zzb.spdmp(tgt, 0.0, k["x0"], k["th0"], k["T"], k["c"], zzb.FactBoomerang(k["Zg"], np.zeros(n), lref, sigma, rho=rho),
                                         seed=k["seed"], adapt=k.get("adapt", False))
    else:
        c = zzb.LocalBound(k["c"]) if k["kind"] == "localbound" else k["c"]
        Xi, _, (acc, num), _ = zzb.spdmp(tgt, 0.0, k["x0"], k["th0"], k["T"], c, zzb.ZigZag(k["Zg"], np.zeros(n)), seed=k["seed"],
                                         adapt=k.get("adapt", False))
    return summary(Xi.events, num, acc)
