"""Factorised Boomerang (F::FactBoomerang in src/sfact.jl): the oracle against the reference's own statistical tests
(test/maintest.jl:89-137), the shared sincos / normal primitives, and the windowed schedule (host emulation on the
kernel's per-coordinate code) against the oracle, bit for bit."""
import numpy as np
import pytest

import oracle_lib as O


def boom_inputs(zzb, G, scale, rng, x_scale=1.0, mu=None):
    Zg = G.scaled(scale)
    diag = Zg.to_scipy().diagonal()
    d = G.n
    x0 = x_scale * rng.random(d)
    th0 = rng.standard_normal(d) / np.sqrt(diag)           # theta0 = sqrt(Diagonal(Z.Gamma)) \ randn(d), maintest.jl:101
    return Zg, diag ** -0.5, x0, th0, G.colnorms()


def test_sincos_matches_libm_to_one_ulp():
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-1e5, 1e5, 20000), rng.uniform(-10, 10, 20000), rng.uniform(-1e-3, 1e-3, 5000), [0.0]])
    worst = 0.0
    for x in xs:
        s, c = O.sincos(x)
        for got, ref in ((s, np.sin(x)), (c, np.cos(x))):
            worst = max(worst, abs(got - ref) / np.spacing(abs(ref)) if ref != 0 else abs(got))
    assert worst <= 1.0
    assert O.sincos(0.0) == (0.0, 1.0)


def test_randn_is_standard_normal():
    rng = np.random.default_rng(1)
    u = rng.random((200000, 2)) * (1 - 2e-16) + 1e-16
    z = np.array([O.lib().zzo_randn(a, b) for a, b in u[:50000]])
    assert abs(z.mean()) < 0.02 and abs(z.var() - 1) < 0.03 and abs((z ** 3).mean()) < 0.06 and abs((z ** 4).mean() - 3) < 0.15


@pytest.mark.parametrize("scale,mode,covtol", [
    (0.85, O.RNG_SEQ | O.ARITH_INPLACE | O.GRAPH_ALL, 4.5),   # "FactBoomerang": pdmp, test/maintest.jl:89-112
    (1.2, O.RNG_SEQ | O.ARITH_INPLACE, 4.0),                  # "SFactBoomerang": spdmp, test/maintest.jl:114-137
    (1.2, O.PARITY_MODE, 4.0),                                # the GPU contract (per-coordinate streams and clocks)
    (1.2, O.RNG_CTR | O.ARITH_INPLACE, 4.0),
    (0.85, O.RNG_SEQ | O.ARITH_LAZY, 4.5),
])
def test_reference_moment_tests(zzb, scale, mode, covtol):
    d, T = 8, 3000.0
    G = zzb.random_spd(d, seed=2)
    Gd = G.to_scipy().toarray()
    rng = np.random.default_rng(11)
    Zg, sigma, x0, th0, c = boom_inputs(zzb, G, scale, rng, x_scale=0.2 if mode & O.GRAPH_ALL else 1.0)
    r = O.spdmp(G, Zg, 0.0, x0, th0, T, c, seed=(3, 4), mode=mode, boom=(sigma, 0.3, 0.0))
    ts, xs = O.boom_discretize(r, np.zeros(d), 0.5)
    assert np.abs(xs.mean(0)).mean() < 2 / np.sqrt(T)
    assert np.abs(np.cov(xs.T) - np.linalg.inv(Gd)).mean() < covtol / np.sqrt(T)
    assert 0 < r.acc.sum() < r.num and len(r.events) > r.acc.sum()      # the trace also holds the refreshments


def test_lazy_and_inplace_arithmetic_agree(zzb):
    """Anchored rotation vs the reference's incremental in-place rotation: same event sequence, times equal to rounding."""
    G = zzb.grid_precision(6, 6)
    rng = np.random.default_rng(2)
    Zg, sigma, x0, th0, c = boom_inputs(zzb, G, 1.0, rng)
    a = O.spdmp(G, Zg, 0.0, x0, th0, 15.0, c, seed=(5, 6), mode=O.RNG_CTR | O.ARITH_LAZY, boom=(sigma, 1.0, 0.2))
    b = O.spdmp(G, Zg, 0.0, x0, th0, 15.0, c, seed=(5, 6), mode=O.RNG_CTR | O.ARITH_INPLACE, boom=(sigma, 1.0, 0.2))
    n = min(len(a.events), len(b.events))
    assert n > 100 and abs(len(a.events) - len(b.events)) <= 1
    assert np.array_equal(a.events["i"][:n], b.events["i"][:n])
    assert np.allclose(a.events["t"][:n], b.events["t"][:n], rtol=1e-9, atol=1e-11)
    assert np.allclose(a.events["x"][:n], b.events["x"][:n], rtol=1e-7, atol=1e-9)


BOOM_CASES = [
    # (graph, scale of Z.Gamma, mu?, h?, T, lambdaref, rho, adapt, c scale, sim options)
    ("spd8", 1.2, False, False, 200.0, 0.3, 0.0, False, 1.0, {}),
    ("spd8", 0.85, True, False, 200.0, 0.5, 0.4, True, 0.05, {}),
    ("grid12", 1.0, False, False, 20.0, 2.0, 0.0, False, 1.0, {}),
    ("grid12", 1.0, False, False, 5.0, 20.0, 0.3, False, 1.0, dict(target_frac=0.2, tag_limit=40)),
    ("grid12csr", 1.0, False, False, 5.0, 5.0, 0.0, False, 1.0, {}),
    ("sparse50", 1.1, True, True, 30.0, 3.0, 0.2, True, 1.0, dict(delta0=0.5, target_frac=3.0)),
]


def boom_case(zzb, name):
    graph, scale, with_mu, with_h, T, lref, rho, adapt, cs, opts = next(c for c in BOOM_CASES if c == name)
    rng = np.random.default_rng(sum(map(ord, graph)))
    if graph == "spd8":
        G = zzb.random_spd(8, seed=2)
    elif graph.startswith("grid12"):
        G = zzb.grid_precision(12, 12)
    else:
        G = next(g for g in (zzb.random_sparse_spd(50, deg=2, seed=s) for s in range(4, 40)) if np.diff(g.colptr).max() <= 8)
    d = G.n
    Zg, sigma, x0, th0, c = boom_inputs(zzb, G, scale, rng)
    x0 = rng.standard_normal(d)
    mu = 0.3 * rng.standard_normal(d) if with_mu else None
    h = 0.2 * rng.standard_normal(d) if with_h else None
    opts = dict(opts)
    if graph.endswith("csr"):
        opts["tag_limit"] = opts.get("tag_limit", 0x0F000000) | 0x80000000
    return G, Zg, sigma, x0, th0, cs * c, mu, h, T, lref, rho, adapt, opts


@pytest.mark.parametrize("case", BOOM_CASES, ids=lambda c: f"{c[0]}-{c[5]}-{c[6]}")
def test_window_schedule_equals_oracle(zzb, case):
    G, Zg, sigma, x0, th0, c, mu, h, T, lref, rho, adapt, opts = boom_case(zzb, case)
    boom = (sigma, lref, rho)
    ref = O.spdmp(G, Zg, 0.0, x0, th0, T, c, mu=mu, h=h, seed=(7, 8), mode=O.PARITY_MODE, boom=boom, adapt=adapt)
    got = O.window_sim(G, Zg, 0.0, x0, th0, T, c, mu=mu, h=h, seed=(7, 8), boom=boom, adapt=adapt, **opts)
    O.assert_same_run(ref, got)
    assert len(ref.events) > 50 and ref.acc.sum() > 10 and len(ref.events) > ref.acc.sum()
    if adapt:
        assert np.any(ref.c > c)


def test_bound_violation_is_reported(zzb):
    G = zzb.random_spd(8, seed=2)
    rng = np.random.default_rng(3)
    Zg, sigma, x0, th0, c = boom_inputs(zzb, G, 1.0, rng)
    with pytest.raises(O.BoundError):
        O.spdmp(G, Zg, 0.0, x0, th0, 500.0, 1e-3 * c, seed=(1, 1), mode=O.PARITY_MODE, boom=(sigma, 0.3, 0.0))
    with pytest.raises(O.BoundError):
        O.window_sim(G, Zg, 0.0, x0, th0, 500.0, 1e-3 * c, seed=(1, 1), boom=(sigma, 0.3, 0.0))
