"""GPU parity: the CUDA event loop (through the C-ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


class R:
    pass


def run_gpu(zzb, Gt, Gb, t0, x0, th0, T, c, *, h=None, mu=None, seed=(1, 2), adapt=False, factor=1.8, tune=None):
    Z = zzb.ZigZag(Gb, np.zeros(Gb.n) if mu is None else mu)
    Xi, (t, x, th), (acc, num), cc = zzb.spdmp(zzb.GaussianPotential(Gt, h), t0, x0, th0, T, c, Z, seed=seed,
                                               adapt=adapt, factor=factor, tune=tune)
    r = R()
    r.events, r.t, r.x, r.theta, r.c, r.acc, r.num = Xi.events, t, x, th, cc, acc, num
    r.stats = Xi.stats
    return r, Xi


@pytest.mark.parametrize("n,T", [(4, 10.0), (16, 5.0), (32, 3.0), (100, 2.0)])
def test_grid_gmrf_bit_exact(gpu, n, T):
    G, x0, th0, c = gpu.gmrf_config(n)
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c)
    got, Xi = run_gpu(gpu, G, G, 0.0, x0, th0, T, c)
    O.assert_same_run(ref, got)
    m1, m2 = Xi.moments
    assert np.allclose(m1, ref.m1, rtol=1e-12, atol=1e-300)
    assert np.allclose(m2, ref.m2, rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("delta0,frac", [(1e-3, 0.05), (0.5, 3.0)])
def test_window_length_does_not_change_results(gpu, delta0, frac):
    G, x0, th0, c = gpu.gmrf_config(24)
    ref = O.spdmp(G, G, 0.0, x0, th0, 4.0, c)
    got, _ = run_gpu(gpu, G, G, 0.0, x0, th0, 4.0, c, tune=dict(delta0=delta0, target_frac=frac))
    O.assert_same_run(ref, got)


def test_rectangular_grid_tight_bound(gpu):
    G = gpu.grid_precision(5, 9)
    rng = np.random.default_rng(3)
    x0, th0 = rng.standard_normal(45), rng.choice(np.array([-1.0, 1.0]), 45)
    c = np.full(45, np.sqrt(np.finfo(float).eps))
    ref = O.spdmp(G, G, 0.0, x0, th0, 8.0, c)
    got, _ = run_gpu(gpu, G, G, 0.0, x0, th0, 8.0, c)
    O.assert_same_run(ref, got)


@pytest.mark.parametrize("seed", range(4))
def test_random_sparse_with_mu_h_adapt(gpu, seed):
    """General sparse columns (CSR kernel), target != sampler matrix, Z.mu != 0, linear term, adaptation of c."""
    d = 60
    Gt = gpu.random_sparse_spd(d, deg=2 + seed % 3, seed=seed)
    Gb = Gt.scaled(0.9)
    rng = np.random.default_rng(seed)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
    mu, h = 0.1 * rng.standard_normal(d), 0.3 * rng.standard_normal(d)
    c = 0.2 * Gt.colnorms()
    ref = O.spdmp(Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True)
    got, _ = run_gpu(gpu, Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True)
    O.assert_same_run(ref, got)
    assert (got.c != c).any()


def test_dense_columns_slow_path_and_pdmp(gpu):
    G = gpu.random_spd(12, density=0.5)
    rng = np.random.default_rng(0)
    x0, th0 = rng.random(12), rng.choice(np.array([-1.0, 1.0]), 12)
    c = 2.0 * G.colnorms()
    ref = O.spdmp(G, G, 0.0, x0, th0, 50.0, c, adapt=True)
    Xi, (t, x, th), (acc, num), cc = gpu.pdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 50.0, c, gpu.ZigZag(G, np.zeros(12)),
                                              seed=(1, 2), adapt=True)
    assert num == ref.num and np.array_equal(acc, ref.acc)
    assert np.array_equal(Xi.events["t"].view(np.uint64), ref.events["t"].view(np.uint64))


def test_bound_violation_error(gpu):
    """error("Tuning parameter `c` too small.") of sfact.jl:124 through the C-ABI (ZZB_E_BOUND)."""
    G, x0, th0, c = gpu.gmrf_config(8)
    with pytest.raises(gpu.BoundError, match="Tuning parameter `c` too small"):
        run_gpu(gpu, G, G.scaled(0.5), 0.0, x0, th0, 5.0, np.full(G.n, 1e-9))


def test_empty_horizon(gpu):
    G, x0, th0, c = gpu.gmrf_config(8)
    got, _ = run_gpu(gpu, G, G, 0.0, x0, th0, 0.0, c)
    assert len(got.events) == 0 and got.num == 0 and np.array_equal(got.x, x0)


def test_tag_rebase_and_trace_draining(gpu):
    """Tiny iteration-tag budget (forces rebases) and a trace buffer that only holds a few windows (forces the
    drain-and-relaunch protocol): results must not change."""
    G, x0, th0, c = gpu.gmrf_config(16)
    ref = O.spdmp(G, G, 0.0, x0, th0, 6.0, c)
    prob = gpu.Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(G.n)))
    run = gpu.Run(prob, record_trace=True, trace_capacity=1)  # clamped to the minimum: d * MAXFLIP + 2 records
    run.set(tag_limit=40)
    run.upload(0.0, x0, th0, c, seed=(1, 2))
    run.execute(6.0)
    ev = run.events()
    st = run.stats()
    assert st["rebases"] > 0 and st["launches"] > 4
    assert np.array_equal(ev["i"], ref.events["i"]) and np.array_equal(ev["t"].view(np.uint64), ref.events["t"].view(np.uint64))
    acc, num = run.counts()
    assert num == ref.num


def test_maintest_moments_on_gpu(gpu):
    """test/maintest.jl:37-61 statistical acceptance test run on the device path (d = 8, T = 1000)."""
    import math
    d, T = 8, 1000.0
    G = gpu.random_spd(d, seed=2)
    rng = np.random.default_rng(5)
    x0, th0 = rng.random(d), rng.choice(np.array([-1.0, 1.0]), d)
    c = 0.7 * G.colnorms()
    Z = gpu.ZigZag(G.scaled(0.9), np.zeros(d))
    Xi, _, (acc, num), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, Z, seed=(11, 12))
    ts, xs = gpu.discretize(Xi, 0.5)
    Sigma = np.linalg.inv(G.to_scipy().toarray())
    assert np.mean(np.abs(xs.mean(axis=0))) < 2 / math.sqrt(T)
    assert np.mean(np.abs(np.cov(xs.T) - Sigma)) < 2.5 / math.sqrt(T)
    ref = O.spdmp(G, G.scaled(0.9), 0.0, x0, th0, T, c, seed=(11, 12))
    assert num == ref.num and np.array_equal(Xi.events["t"].view(np.uint64), ref.events["t"].view(np.uint64))


def test_full_size_properties(gpu):
    """d = 10^6 (BASELINE configs[4] on one GPU): size-independent properties -- the result does not depend on the
    window length, the trace is time-sorted, counters agree with the trace, moment sums agree with the trace."""
    G, x0, th0, c = gpu.gmrf_config(1000)
    prob = gpu.Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(G.n)))
    T = 0.25
    res = []
    for frac, rec in ((0.05, True), (0.2, False)):
        run = gpu.Run(prob, record_trace=rec)
        run.set(target_frac=frac)
        run.upload(0.0, x0, th0, c, seed=(3, 4))
        run.execute(T)
        acc, num = run.counts()
        res.append((acc, num, run.final_state(), run.sums(), run.events() if rec else None))
        run.close()
    (a0, n0, f0, s0, ev), (a1, n1, f1, s1, _) = res
    assert n0 == n1 and np.array_equal(a0, a1)
    for u, v in zip(f0 + s0, f1 + s1):
        assert np.array_equal(u.view(np.uint64), v.view(np.uint64))
    assert len(ev) == a0.sum() and np.all(np.diff(ev["t"]) >= 0) and ev["t"][-1] >= T and np.all(ev["t"][:-1] < T)
    assert np.array_equal(np.bincount(ev["i"] - 1, minlength=G.n), a0)
    # first-moment sums recomputed from the trace (trace.jl:182-200, unscaled) for a sample of coordinates
    xs, ts, s1chk = x0.copy(), np.zeros(G.n), np.zeros(G.n)
    sel = ev[ev["i"] <= 2000]
    for t2, i, xi, _ in sel:
        s1chk[i - 1] += (xs[i - 1] + xi) * (t2 - ts[i - 1])
        ts[i - 1], xs[i - 1] = t2, xi
    assert np.array_equal(s1chk[:2000].view(np.uint64), s0[0][:2000].view(np.uint64))
    assert 0.05 < a0.sum() / n0 < 0.5  # acceptance rate in the plausible range (~0.18 at stationarity)


@pytest.mark.parametrize("cval", [1.0, 0.05])
def test_local_bound_variant(gpu, cval):
    """spdmp(grad, t0, x0, th0, T, LocalBound(c), Z) (src/local.jl:95-149) through the C-ABI flag ZZB_FLAG_LOCAL_BOUND."""
    LB = O.PARITY_MODE | O.LOCAL_BOUND
    G, x0, th0, _ = gpu.gmrf_config(24)
    c = np.full(G.n, cval)
    ref = O.spdmp(G, G, 0.0, x0, th0, 5.0, c, mode=LB, adapt=True)
    Xi, (t, x, th), (acc, num), C = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 5.0, gpu.LocalBound(c),
                                               gpu.ZigZag(G, np.zeros(G.n)), seed=(1, 2), adapt=True)
    got = R()
    got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = Xi.events, t, x, th, C.c, acc, num
    O.assert_same_run(ref, got)
    # general sparse target with a linear term
    d = 60
    Gt = gpu.random_sparse_spd(d, deg=3, seed=4)
    rng = np.random.default_rng(4)
    x0, th0, h = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d), 0.3 * rng.standard_normal(d)
    ref = O.spdmp(Gt, Gt, 0.0, x0, th0, 15.0, np.full(d, 0.3), h=h, mode=LB, adapt=True)
    Xi, (t, x, th), (acc, num), C = gpu.spdmp(gpu.GaussianPotential(Gt, h), 0.0, x0, th0, 15.0, gpu.LocalBound(np.full(d, 0.3)),
                                               gpu.ZigZag(Gt, np.zeros(d)), seed=(1, 2), adapt=True)
    got = R()
    got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = Xi.events, t, x, th, C.c, acc, num
    O.assert_same_run(ref, got)


@pytest.mark.parametrize("n,T,kap", [(8, 8.0, 1.0), (32, 3.0, 0.5)])
def test_sticky_zigzag(gpu, n, T, kap):
    """sspdmp(grad, t0, x0, th0, T, c, Z, kappa) (src/ss_fact.jl:159-217) through zzb_run_upload_kappa / ZZB_FLAG_STICKY."""
    G, x0, th0, c = gpu.gmrf_config(n)
    kappa = np.full(G.n, kap)
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c, kappa=kappa)
    Xi, (t, x, th), (acc, num), cc = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, gpu.ZigZag(G, np.zeros(G.n)), kappa,
                                                seed=(1, 2))
    got = R()
    got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = Xi.events, t, x, th, cc, Xi.acc_per_coordinate, num
    got.s1, got.s2 = Xi.sums
    O.assert_same_run(ref, got)
    assert acc == ref.acc.sum() and (Xi.events["theta"] == 0).sum() > 10
    # freeze records carry x = -0*theta: a signed zero, bit-identical to the oracle (checked by assert_same_run above)
    assert np.all(Xi.events["x"][Xi.events["theta"] == 0] == 0)


def test_sticky_1d_closed_form_on_gpu(gpu):
    """test/sticky.jl:7-36 on the device path."""
    import math
    sigma, mu, kap, T = math.sqrt(0.5), 0.9, 1.5, 2000.0
    G = gpu.CSC.from_dense(np.array([[1 / sigma ** 2]]))
    Z = gpu.ZigZag(gpu.CSC.from_dense(np.array([[1.0]])), np.zeros(1))
    Xi, _, (acc, num), _ = gpu.sspdmp(gpu.GaussianPotential(G, np.array([mu / sigma ** 2])), 0.0, np.array([1.0]), np.array([0.8]), T,
                                      np.array([20.0]), Z, np.array([kap]), seed=(3, 4))
    ts, xs = gpu.discretize(Xi, 0.2)
    xs = xs[:, 0]
    w = math.sqrt(2 * math.pi) * sigma / (math.sqrt(2 * math.pi) * sigma + math.exp(-0.5 * mu ** 2 / sigma ** 2) / kap)
    assert abs(np.mean(xs != 0) - w) < 2.5 / math.sqrt(T)
    assert abs(xs.mean() - w * mu) < 5.0 / math.sqrt(T)
    assert abs((xs ** 2).mean() - w * (sigma ** 2 + mu ** 2)) < 5.0 / math.sqrt(T)


def test_fact_sampler_iterator(gpu):
    """FactSampler / trace(FS, T) (src/sfactiter.jl): pulled events equal the one-shot trace (prefix), in time order."""
    import itertools
    G, x0, th0, c = gpu.gmrf_config(16)
    Z = gpu.ZigZag(G, np.zeros(G.n))
    ref = O.spdmp(G, G, 0.0, x0, th0, 6.0, c, seed=(1, 2))
    FS = gpu.FactSampler(gpu.GaussianPotential(G), (0.0, (x0, th0)), c, Z, seed=(1, 2), windows_per_pull=3)
    got = list(itertools.islice(iter(FS), 500))
    assert [e[1][1] for e in got] == ref.events["i"][:500].tolist()
    assert [e[0] for e in got] == ref.events["t"][:500].tolist()
    tr = gpu.trace(FS, 3.0)      # drops the first event (upstream quirk) and stops at the first event with t > T
    n = int(np.searchsorted(ref.events["t"], 3.0, side="right"))
    assert np.array_equal(tr.events["t"], ref.events["t"][1:n]) and np.array_equal(tr.events["i"], ref.events["i"][1:n])


def test_config1_two_dimensional_pdmp_and_one_dimensional(gpu):
    """BASELINE configs[0] (2-d Gaussian through pdmp) and the smallest possible problem (d = 1) on the device path."""
    G = gpu.CSC.from_dense(np.array([[2.0, -1.0], [-1.0, 2.0]]))
    x0, th0 = np.array([0.3, -0.2]), np.array([1.0, -1.0])
    c = 0.7 * G.colnorms()
    ref = O.spdmp(G, G.scaled(0.9), 0.0, x0, th0, 300.0, c, seed=(5, 6))
    Xi, (t, x, th), (acc, num), cc = gpu.pdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 300.0, c, gpu.ZigZag(G.scaled(0.9), np.zeros(2)), seed=(5, 6))
    got = R()
    got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = Xi.events, t, x, th, cc, acc, num
    O.assert_same_run(ref, got)
    G1 = gpu.CSC.from_dense(np.array([[1.5]]))
    ref = O.spdmp(G1, G1, 0.0, np.array([0.4]), np.array([1.0]), 200.0, np.array([0.5]), seed=(7, 8))
    got, _ = run_gpu(gpu, G1, G1, 0.0, np.array([0.4]), np.array([1.0]), 200.0, np.array([0.5]), seed=(7, 8))
    O.assert_same_run(ref, got)
    assert len(ref.events) > 50


def chain_precision(zzb, p):
    """Gamma = tridiag(-1, 2, -1) + 0.1 I with Gamma[1,1] = Gamma[p,p] = 1.1 (the chain of test/sparsesticky.jl:17-30)."""
    import scipy.sparse as sp
    main = np.full(p, 2.1); main[0] = main[-1] = 1.1
    return zzb.CSC.from_scipy(sp.diags([main, -np.ones(p - 1), -np.ones(p - 1)], [0, 1, -1]).tocsc())


def test_config4_sticky_chain_full_size_properties(gpu):
    """BASELINE configs[3] at its full size p = 10^5 (spike-and-slab chain started at x0 = 0, kappa = 2000/p)
    through size-independent properties: the result does not depend on the window length, the trace is time-sorted, every
    coordinate alternates thaw / (reflections) / freeze consistently, counters agree with the trace."""
    p, T = 100000, 3.0
    G = chain_precision(gpu, p)
    rng = np.random.default_rng(0)
    x0, th0 = np.zeros(p), rng.choice(np.array([-1.0, 1.0]), p)
    c, kappa = np.full(p, 2.5), np.full(p, 2000.0 / p)
    out = []
    for frac in (0.05, 0.4):
        Xi, (t, x, th), (acc, num), _ = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, gpu.ZigZag(G, np.zeros(p)), kappa,
                                                    seed=(5, 6), tune=dict(target_frac=frac))
        out.append((Xi.events, t, x, th, acc, num))
    (ev, t, x, th, acc, num), (ev2, t2, x2, th2, acc2, num2) = out
    assert num == num2 and acc == acc2 and np.array_equal(ev, ev2)
    for u, v in ((t, t2), (x, x2), (th, th2)):
        assert np.array_equal(u.view(np.uint64), v.view(np.uint64))
    assert np.all(np.diff(ev["t"]) >= 0) and ev["t"][-1] >= T and np.all(ev["t"][:-1] < T)
    # per coordinate: events alternate between "moving" (theta != 0) and "frozen" (theta == 0, x == 0) consistently
    order = np.lexsort((ev["t"], ev["i"]))
    e = ev[order]
    first = np.r_[True, e["i"][1:] != e["i"][:-1]]
    frozen_after = e["theta"] == 0
    assert np.all(e["x"][frozen_after] == 0)                    # a freeze records x = -0 * theta
    prev_frozen = np.r_[False, frozen_after[:-1]] & ~first
    assert not np.any(prev_frozen & frozen_after)               # never two freezes in a row
    thaw = prev_frozen & ~frozen_after
    assert thaw.sum() > 0 and np.all(e["x"][thaw] == 0)         # a thaw leaves from 0
    refl = ~frozen_after & ~prev_frozen                         # moving -> moving: accepted reflections
    assert refl.sum() == acc
    assert 0.0 < np.mean(th != 0) < 1.0 and frozen_after.sum() > 0.2 * p   # x0 = 0 with theta0 = +-1 starts moving; most coordinates have frozen by T


def test_testiter_fact_sampler_moments(gpu):
    """test/testiter.jl:4-31 on the device path: tr = trace(FactSampler(grad, t0 => (x0, theta0), c, Z), T) with d = 8,
    Z = ZigZag(0.9 Gamma, 0), T = 1000; mean(tr) and the discretised moments within the reference's tolerances."""
    import math
    d, T = 8, 1000.0
    G = gpu.random_spd(d, seed=2)
    rng = np.random.default_rng(8)
    x0, th0 = rng.random(d), rng.choice(np.array([-1.0, 1.0]), d)
    c = 0.7 * G.colnorms()
    FS = gpu.FactSampler(gpu.GaussianPotential(G), (0.0, (x0, th0)), c, gpu.ZigZag(G.scaled(0.9), np.zeros(d)), seed=(21, 22))
    tr = gpu.trace(FS, T)
    assert 0.1 / math.sqrt(T) < np.mean(np.abs(gpu.mean(tr))) < 2 / math.sqrt(T)          # testiter.jl:26
    ts, xs = gpu.discretize(tr, 0.5)
    assert np.mean(np.abs(xs.mean(axis=0))) < 2 / math.sqrt(T)                            # :30
    assert np.mean(np.abs(np.cov(xs.T) - np.linalg.inv(G.to_scipy().toarray()))) < 2.5 / math.sqrt(T)   # :31


@pytest.mark.parametrize("lattice", [True, False])
def test_zigzag_refreshments_on_device(gpu, lattice):
    """spdmp with Z = ZigZag(Gamma, mu, sigma; lambdaref > 0) (hasrefresh: src/fact_samplers.jl:19; refresh branch src/sfact.jl:78-114,
    188-190) on the device (zz_run_kernel_*_refresh): events (reflections AND refreshments), counters, state against the oracle's
    per-coordinate-clock contract zzo_spdmp_refresh, bit for bit; refreshments are trace events but not acceptances."""
    rng = np.random.default_rng(6)
    if lattice:
        G = gpu.grid_precision(30, 20, shift=0.3)
        Gb, mu, h = G, None, None
    else:
        G = next(g for g in (gpu.random_sparse_spd(80, deg=2, seed=sd) for sd in range(3, 60)) if np.diff(g.colptr).max() <= 8)
        Gb, mu, h = G.scaled(0.9), 0.1 * rng.standard_normal(80), 0.2 * rng.standard_normal(80)
    d = G.n
    sigma = 0.5 + rng.random(d)
    x0, th0 = rng.standard_normal(d), sigma * rng.choice(np.array([-1.0, 1.0]), d)
    c = 3.0 * G.colnorms() * sigma.max()
    lam = 0.8 * d
    ref = O.spdmp(G, Gb, 0.0, x0, th0, 5.0, c, h=h, mu=mu, seed=(5, 7), refresh=(sigma, lam), adapt=True)
    Z = gpu.ZigZag(Gb, np.zeros(d) if mu is None else mu, sigma, lambdaref=lam)
    Xi, (t, x, th), (acc, num), cc = gpu.spdmp(gpu.GaussianPotential(G, h), 0.0, x0, th0, 5.0, c.copy(), Z, seed=(5, 7), adapt=True)
    got = R()
    got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = Xi.events, t, x, th, cc, acc, num
    O.assert_same_run(ref, got)
    assert len(ref.events) - int(ref.acc.sum()) > 100 and int(ref.acc.sum()) > 100
    assert set(np.round(np.abs(Xi.events["theta"]) / sigma[Xi.events["i"] - 1], 12)) == {1.0}    # |theta_i| = sigma_i after every event
