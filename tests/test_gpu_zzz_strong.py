"""The strong-bound sparse sticky kernel (src/sparsestickyzz.jl as the reference runs config 4; csrc/zz_strong.h,
zz_run_kernel_csr_strong) against its contract zzo_sparsestickyzz_ctr, bit for bit.  First green on a B200 in round 2
(profiles/r02a_strong.log); on the CPU the same per-coordinate code equals the contract inside the schedule emulation
(tests/test_sparse_sticky.py)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from sticky_stats import chain_precision

pytestmark = pytest.mark.gpu


def run_strong(zzb, G, x0, th0, T, c, kappa, rule, seed, h=None, tune=None):
    p = G.n
    prob = zzb.Problem(zzb.GaussianPotential(G, h), zzb.ZigZag(G, np.zeros(p)))
    run = zzb.Run(prob, kappa=np.full(p, kappa))
    try:
        run.set(strong_c=c, strong_rule={"sticky": 0, "reversible": 1}[rule], **(tune or {}))
        run.upload(0.0, x0, th0, np.full(p, c), seed=seed)
        run.execute(T)

        class R:
            pass
        r = R()
        r.events = run.events()
        r.t, r.x, r.theta, r.c = run.final_state()
        r.acc, r.num = run.counts()
        r.stats, r.device_ms = run.stats(), run.device_ms
        return r
    finally:
        run.close()
        prob.close()


@pytest.mark.parametrize("p,kappa,T,rule", [(12, 0.5, 50.0, "sticky"), (30, 0.3, 80.0, "reversible"), (2000, 0.5, 30.0, "sticky")])
def test_chain_bit_exact(gpu, p, kappa, T, rule):
    G = chain_precision(gpu, p)
    rng = np.random.default_rng(p)
    x0 = np.where(rng.random(p) < 0.3, rng.standard_normal(p), 0.0)
    th0 = rng.choice(np.array([-1.0, 1.0]), p)
    ref = O.sparsestickyzz(G, x0, th0, T, 2.5, kappa, rule=rule, seed=(3, 4), ctr=True)
    got = run_strong(gpu, G, x0, th0, T, 2.5, kappa, rule, (3, 4))
    O.assert_same_run(ref, got)


def test_config4_full_size_timing(gpu):
    """p = 10^5 chain, x0 = 0 (all frozen), kappa = 2000/p, c = 2.5: parity at T = 5 and the passes per window."""
    p, T = 100000, 5.0
    G = chain_precision(gpu, p)
    x0, th0 = np.zeros(p), np.ones(p)
    ref = O.sparsestickyzz(G, x0, th0, T, 2.5, 2000.0 / p, seed=(5, 6), ctr=True)
    got = run_strong(gpu, G, x0, th0, T, 2.5, 2000.0 / p, "sticky", (5, 6))
    O.assert_same_run(ref, got)
    st = got.stats
    print(f"strong-bound config 4: {len(got.events)} events in {got.device_ms:.2f} ms, {st['windows']} windows, {st['passes']} passes")


def test_sspdmp3_api_matches_the_contract(gpu):
    """The reference-shaped call sspdmp3(grad, u0, T, c, nothing, Z, kappa; rule) (src/sparsestickyzz.jl:405-422) through
    zzb200.sspdmp3: same events as the contract oracle, acc = accepted reflections."""
    p, kappa, T = 300, 0.4, 40.0
    G = chain_precision(gpu, p)
    rng = np.random.default_rng(11)
    x0 = np.where(rng.random(p) < 0.25, rng.standard_normal(p), 0.0)
    th0 = rng.choice(np.array([-1.0, 1.0]), p)
    ref = O.sparsestickyzz(G, x0, th0, T, 2.5, kappa, rule="reversible", seed=(7, 8), ctr=True)
    Xi, (acc, num), (t, x, th) = gpu.sspdmp3(gpu.GaussianPotential(G), (x0, th0), T, 2.5, None, gpu.ZigZag(G, np.zeros(p)), kappa,
                                             rule="reversible", seed=(7, 8))
    assert num == ref.num and acc == int(ref.acc.sum())
    assert np.array_equal(Xi.events["i"], ref.events["i"])
    for f in ("t", "x", "theta"):
        assert np.array_equal(Xi.events[f].view(np.uint64), ref.events[f].view(np.uint64))
    assert np.array_equal(x.view(np.uint64), ref.x.view(np.uint64))


def test_sspdmp4_asynchzz_process_bit_exact(gpu):
    """`sspdmp4` (src/asynchzz.jl:250-265) on the device: per-coordinate c and kappa, shifted start, non-unit speeds, velocity kept
    over a freeze -- bit for bit against the contract zzo_strongsticky_ctr; with equal constants and unit speeds it is sspdmp3."""
    import oracle_lib as O
    from test_sparse_sticky import chain_precision
    rng = np.random.default_rng(11)
    for case in range(6):
        G = chain_precision(gpu, int(rng.integers(20, 400))) if case % 2 else gpu.grid_precision(int(rng.integers(4, 20)), int(rng.integers(4, 20)), shift=0.1)
        p = G.n
        x0 = np.where(rng.random(p) < 0.5, rng.standard_normal(p), 0.0)
        th0 = rng.choice(np.array([-1.5, -1.0, 0.5, 1.0]), p)
        h = 0.5 * rng.standard_normal(p) if case % 3 == 0 else None
        t0 = float(rng.choice([0.0, -2.0, 5.0]))
        T = t0 + 20.0
        base = 3.0 + 4.0 * float(np.abs(G.nzval).max()) + (0.0 if h is None else float(np.abs(h).max()))
        c = base * rng.uniform(1.0, 2.0, p)
        kappa = rng.choice(np.array([0.2, 1.0, 3.0]), p)
        ref = O.strongsticky(G, t0, x0, th0, T, c, kappa, h=h, seed=(7, 8 + case))
        Xi, (acc, num) = gpu.sspdmp4(None, gpu.GaussianPotential(G, h), t0, x0, th0, T, c, None, gpu.ZigZag(G, np.zeros(p)), kappa, seed=(7, 8 + case))
        assert num == ref.num and acc == int(ref.acc.sum()) and np.array_equal(Xi.acc_per_coordinate, ref.acc)
        assert len(Xi.events) == len(ref.events) and np.array_equal(Xi.events["i"], ref.events["i"])
        for f in ("t", "x", "theta"):
            assert np.array_equal(Xi.events[f].view(np.uint64), ref.events[f].view(np.uint64)), f
        t, x, th = Xi.final
        assert np.array_equal(x.view(np.uint64), ref.x.view(np.uint64)) and np.array_equal(th.view(np.uint64), ref.theta.view(np.uint64))
    # equal constants, unit speeds, t0 = 0: the :sticky process of sspdmp3
    G = chain_precision(gpu, 300)
    x0 = np.where(rng.random(300) < 0.5, rng.standard_normal(300), 0.0)
    th0 = np.where(x0 != 0.0, rng.choice(np.array([-1.0, 1.0]), 300), 1.0)
    X3, _, _ = gpu.sspdmp3(gpu.GaussianPotential(G), (x0, th0), 30.0, 6.0, None, gpu.ZigZag(G, np.zeros(300)), 0.5, rule="sticky", seed=(1, 2))
    X4, _ = gpu.sspdmp4(None, gpu.GaussianPotential(G), 0.0, x0, th0, 30.0, 6.0, None, gpu.ZigZag(G, np.zeros(300)), 0.5, seed=(1, 2))
    assert np.array_equal(X3.events, X4.events)


@pytest.mark.parametrize("strong", [False, True])
def test_sspdmp2_dense_sticky_on_device(gpu, strong):
    """`sspdmp2` / stickyzz (src/stickyzz.jl:322-338) on the device: the sticky kernels with ZZB_FLAG_STICKY_ZZ (rate floor 0.01,
    coordinates at 0 start frozen), lattice and general sparse graph, bit for bit against the oracle (ZZO_STICKYZZ)."""
    import oracle_lib as O
    rng = np.random.default_rng(21)
    for G in (gpu.grid_precision(14, 11, shift=0.5), gpu.random_sparse_spd(150, deg=2, seed=2)):   # (columns of at most 8 entries: the cap of the sticky kernels)
        d = G.n
        x0 = np.where(rng.random(d) < 0.6, rng.standard_normal(d), 0.0)
        th0 = rng.choice(np.array([-1.5, -1.0, 1.0, 0.5]), d)
        c, kap = 6.0 * G.colnorms(), rng.choice(np.array([0.3, 0.8, 2.0]), d)
        mode = O.PARITY_MODE | O.STICKYZZ | (O.STICKY_STRONG_UB if strong else 0)
        ref = O.spdmp(G, G, 0.0, x0, th0, 8.0, c, kappa=kap, mode=mode)
        Xi, (acc, num) = gpu.sspdmp2(gpu.GaussianPotential(G), 0.0, x0, th0, 8.0, c, None, gpu.ZigZag(G, np.zeros(d)), kap,
                                     strong_upperbounds=strong, seed=(1, 2))
        assert num == ref.num and np.array_equal(Xi.acc_per_coordinate, ref.acc)
        assert len(Xi.events) == len(ref.events) and np.array_equal(Xi.events["i"], ref.events["i"])
        for f in ("t", "x", "theta"):
            assert np.array_equal(Xi.events[f].view(np.uint64), ref.events[f].view(np.uint64)), f
        t, x, th = Xi.final
        assert np.array_equal(x.view(np.uint64), ref.x.view(np.uint64)) and np.array_equal(th.view(np.uint64), ref.theta.view(np.uint64))


def test_sticky_adapt_on_device(gpu):
    """sspdmp / sspdmp2 with adapt = true (src/ss_fact.jl:132-136, src/stickyzz.jl:305-309): c[i] grows by `factor` at an accepted
    proposal with l > lb; bit for bit against the oracle, adapted bounds included."""
    import oracle_lib as O
    G = gpu.grid_precision(12, 13, shift=0.5)
    Gb = G.scaled(0.5)
    rng = np.random.default_rng(12)
    d = G.n
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c, kap = 0.05 * G.colnorms(), np.full(d, 0.8)
    for stickyzz in (False, True):
        ref = O.spdmp(G, Gb, 0.0, x0, th0, 8.0, c, kappa=kap, mode=O.PARITY_MODE | (O.STICKYZZ if stickyzz else 0), adapt=True, factor=1.5)
        if stickyzz:
            Xi, (acc, num) = gpu.sspdmp2(gpu.GaussianPotential(G), 0.0, x0, th0, 8.0, c, None, gpu.ZigZag(Gb, np.zeros(d)), kap, adapt=True, factor=1.5, seed=(1, 2))
            cc = Xi.c
        else:
            Xi, _, (acc, num), cc = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 8.0, c, gpu.ZigZag(Gb, np.zeros(d)), kap, adapt=True, factor=1.5, seed=(1, 2))
        assert (ref.c != c).sum() > 5 and np.array_equal(cc.view(np.uint64), ref.c.view(np.uint64))
        assert num == ref.num and acc == int(ref.acc.sum()) and len(Xi.events) == len(ref.events)
        for f in ("t", "x", "theta"):
            assert np.array_equal(Xi.events[f].view(np.uint64), ref.events[f].view(np.uint64)), f
    with pytest.raises(gpu.BoundError):
        gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 8.0, c, gpu.ZigZag(Gb, np.zeros(d)), kap, seed=(1, 2))


def test_one_call_entries_of_the_sticky_family(gpu):
    """zzb_sspdmp_adapt_run and zzb_sspdmp4_run called the way a Julia `ccall` would (one call, raw pointers): same trace as the
    staged Python path."""
    import ctypes as C
    from zzb200 import _capi
    from zzb200.api import EVENT_DTYPE, Problem
    L = _capi.lib()
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    G = gpu.grid_precision(9, 10, shift=0.5)
    Gb = G.scaled(0.5)
    rng = np.random.default_rng(3)
    d = G.n
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    kap, sd = np.full(d, 0.8), np.array([1, 2], dtype=np.uint64)

    def events_of(run):
        n = C.c_int64()
        _capi.check(L.zzb_trace_len(run, C.byref(n)))
        ev = np.empty(n.value, dtype=EVENT_DTYPE)
        _capi.check(L.zzb_trace_copy(run, ptr(ev), 0, n.value))
        L.zzb_run_free(run)
        return ev

    prob = Problem(gpu.GaussianPotential(G), gpu.ZigZag(Gb, np.zeros(d)))
    c = 0.05 * G.colnorms()
    cio = c.copy()
    run = C.c_void_p()
    _capi.check(L.zzb_sspdmp_adapt_run(prob._h, 0.0, ptr(x0), ptr(th0), 6.0, ptr(cio), ptr(kap), ptr(sd), 1, 1.5, 0, C.byref(run)))
    ev = events_of(run)
    Xi, _, _, cc = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 6.0, c, gpu.ZigZag(Gb, np.zeros(d)), kap, adapt=True, factor=1.5, seed=(1, 2))
    assert np.array_equal(ev, Xi.events) and np.array_equal(cio, cc) and (cio != c).any()
    prob.close()

    prob = Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(d)))
    c4, k4 = 8.0 * G.colnorms() * rng.uniform(1, 2, d), rng.choice(np.array([0.3, 1.0]), d)
    x4 = np.where(rng.random(d) < 0.5, x0, 0.0)
    run = C.c_void_p()
    _capi.check(L.zzb_sspdmp4_run(prob._h, -1.0, ptr(x4), ptr(th0), 9.0, ptr(c4), ptr(k4), ptr(sd), 0, C.byref(run)))
    ev = events_of(run)
    X4, _ = gpu.sspdmp4(None, gpu.GaussianPotential(G), -1.0, x4, th0, 9.0, c4, None, gpu.ZigZag(G, np.zeros(d)), k4, seed=(1, 2))
    assert len(ev) > 100 and np.array_equal(ev, X4.events)
    prob.close()


def test_wide_columns_on_device(gpu):
    """Sticky, refreshment and Boomerang samplers on a graph with columns of 9 .. 32 entries (the out-of-line wide instantiation,
    zz_process_node_wide; round 1 refused more than 8): bit for bit against the oracle."""
    import oracle_lib as O
    from test_boomerang import boom_inputs
    G = gpu.random_sparse_spd(300, deg=5, seed=2)
    d = G.n
    assert 8 < np.diff(G.colptr).max() <= 32
    rng = np.random.default_rng(31)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c, kap = 4.0 * G.colnorms(), rng.choice(np.array([0.3, 0.8, 2.0]), d)
    Z = gpu.ZigZag(G, np.zeros(d))
    ref = O.spdmp(G, G, 0.0, x0, th0, 5.0, c, kappa=kap, seed=(1, 2))
    Xi, (t, x, th), (acc, num), cc = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 5.0, c, Z, kap, seed=(1, 2))
    assert num == ref.num and np.array_equal(Xi.events, ref.events) and np.array_equal(x.view(np.uint64), ref.x.view(np.uint64))
    sigma = 0.5 + rng.random(d)
    ths = sigma * rng.choice(np.array([-1.0, 1.0]), d)
    cr = 3.0 * G.colnorms() * sigma.max()
    ref = O.spdmp(G, G, 0.0, x0, ths, 4.0, cr, seed=(5, 7), refresh=(sigma, 0.8 * d), adapt=True)
    Xi, (t, x, th), (acc, num), cc = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, ths, 4.0, cr, gpu.ZigZag(G, np.zeros(d), sigma, lambdaref=0.8 * d),
                                               seed=(5, 7), adapt=True)
    assert num == ref.num and np.array_equal(Xi.events, ref.events) and np.array_equal(cc.view(np.uint64), ref.c.view(np.uint64))
    Zg, bsig, xb, thb, cb = boom_inputs(gpu, G, 1.0, rng)
    ref = O.spdmp(G, Zg, 0.0, xb, thb, 3.0, cb, seed=(5, 6), boom=(bsig, 20.0, 0.2))
    F = gpu.FactBoomerang(Zg, np.zeros(d), 20.0, bsig, rho=0.2)
    Xi, (t, x, th), (acc, num), cc = gpu.spdmp(gpu.GaussianPotential(G), 0.0, xb, thb, 3.0, cb, F, seed=(5, 6))
    assert num == ref.num and np.array_equal(Xi.events, ref.events) and np.array_equal(x.view(np.uint64), ref.x.view(np.uint64))
    big = gpu.random_sparse_spd(200, deg=20, seed=1)      # columns over 32 entries are still refused
    assert np.diff(big.colptr).max() > 32
    with pytest.raises(gpu.ZZBError, match="at most 32"):
        gpu.sspdmp(gpu.GaussianPotential(big), 0.0, np.ones(200), np.ones(200), 1.0, big.colnorms(), gpu.ZigZag(big, np.zeros(200)), np.ones(200), seed=(1, 2))
