"""Device-side discretisation (zzb_run_discretize / zzb_run_grid; collect(discretize(trace, dt)), src/trace.jl:94-125, produced
while the windows are committed): bit-exact against the same anchored evaluation done in numpy from the returned trace, and
equal to rounding to the reference-style incremental iteration."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def anchored_grid(Xi, mu, dt, n_rows, t_end, boom=False):
    """x_j(t0 + k dt) from the trace with the device's rule: the segment (tf, fs] that contains the grid time is evaluated
    from its anchor; row 0 is x0."""
    d = len(Xi.x0)
    out = np.full((n_rows, d), np.nan)
    out[0] = Xi.x0
    tf, xf, th = np.full(d, Xi.t0), Xi.x0.copy(), Xi.theta0.copy()
    tk = Xi.t0 + dt * np.arange(n_rows)

    def fill(j, fs):
        ks = np.nonzero((tk > tf[j]) & (tk <= fs))[0]
        for k in ks:
            if boom:
                tau = tk[k] - tf[j]
                s, c = O.sincos(tau)
                xm = xf[j] - mu[j]
                out[k, j] = xf[j] if tau == 0.0 else xm * c + th[j] * s + mu[j]
            else:
                out[k, j] = xf[j] + th[j] * (tk[k] - tf[j])

    for t, i, x, thn in Xi.events:
        j = i - 1
        fill(j, t)
        tf[j], xf[j], th[j] = t, x, thn
    for j in range(d):
        fill(j, t_end)
    return out


import oracle_lib as O  # noqa: E402  (only the shared sincos primitive is used here)


@pytest.mark.parametrize("n,T,dt", [(8, 6.0, 0.25), (24, 3.0, 0.1)])
def test_grid_zigzag(gpu, n, T, dt):
    G, x0, th0, c = gpu.gmrf_config(n)
    Xi, _, (acc, num), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, gpu.ZigZag(G, np.zeros(G.n)), seed=(1, 2),
                                     discretize_dt=dt)
    ts, xs = Xi.grid
    t_end = Xi.events["t"][-1]
    assert len(ts) == int(np.floor(T / dt)) + 1 and np.all(ts <= t_end) and not np.isnan(xs).any()
    want = anchored_grid(Xi, None, dt, len(ts), t_end)
    assert np.array_equal(xs.view(np.uint64), want.view(np.uint64))
    rts, rxs = gpu.discretize(Xi, dt)                       # the reference's incremental iteration (host mirror)
    m = min(len(rts), len(ts))
    assert m >= len(ts) - 1 and np.allclose(rxs[:m], xs[:m], rtol=1e-9, atol=1e-9) and np.allclose(rts[:m], ts[:m])
    # the same grid without recording the trace
    Xj, _, (acc2, num2), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, gpu.ZigZag(G, np.zeros(G.n)), seed=(1, 2),
                                       discretize_dt=dt, record_trace=False)
    assert num2 == num and np.array_equal(Xj.grid[1].view(np.uint64), xs.view(np.uint64))


def test_grid_boomerang(gpu):
    G = gpu.grid_precision(10, 10)
    rng = np.random.default_rng(4)
    diag = G.to_scipy().diagonal()
    x0, th0 = rng.standard_normal(G.n), rng.standard_normal(G.n) / np.sqrt(diag)
    F = gpu.FactBoomerang(G, np.zeros(G.n), 4.0, rho=0.1)
    Xi, _, _, _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 5.0, G.colnorms(), F, seed=(3, 4), discretize_dt=0.2)
    ts, xs = Xi.grid
    want = anchored_grid(Xi, F.mu, 0.2, len(ts), Xi.events["t"][-1], boom=True)
    assert np.array_equal(xs.view(np.uint64), want.view(np.uint64))
    rts, rxs = gpu.discretize(Xi, 0.2)
    m = min(len(rts), len(ts))
    assert np.allclose(rxs[:m], xs[:m], rtol=1e-8, atol=1e-8)


def test_grid_full_size_statistics(gpu):
    """d = 10^6 without a trace: the grid rows are produced on the device; their spatial mean / variance stay finite and the
    row count matches T / dt."""
    G, x0, th0, c = gpu.gmrf_config(1000)
    prob = gpu.Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(G.n)))
    run = gpu.Run(prob, record_trace=False)
    run.discretize(0.05, 6)
    run.upload(0.0, x0, th0, c, seed=(1, 2))
    run.execute(0.25)
    ts, xs = run.grid()
    assert len(ts) == 6 and not np.isnan(xs).any() and np.array_equal(xs[0], x0)
    step = np.abs(xs[1:] - xs[:-1]).max()
    assert 0 < step <= 0.05 * (1 + 1e-12)                  # unit speed: no coordinate moves more than dt between rows
    run.close(); prob.close()


def test_subtrace_filter_and_inclusion_prob_on_device(gpu):
    """f1 of SURVEY 8(f): subtrace (src/trace.jl:275-290) applied at the source -- only the selected coordinates' events are
    recorded, renumbered -- equals subtrace() of the full trace; inclusion_prob (:161-178) of a sticky run from the device
    accumulator equals the host mirror evaluated on the full trace; the sspdmp options reversible / strong_upperbounds (src/ss_fact.jl:97-113)
    against the oracle."""
    import oracle_lib as O
    G, x0, th0, c = gpu.gmrf_config(20)
    Z = gpu.ZigZag(G, np.zeros(G.n))
    full, _, (acc, num), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 3.0, c, Z, seed=(1, 2))
    J = np.array([1, 2, 7, 40, 41, 199, 200, 399, 400])
    sub, _, (acc2, num2), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 3.0, c, Z, seed=(1, 2), trace_filter=J)
    want = gpu.subtrace(full, J)
    assert num2 == num and np.array_equal(acc, acc2)
    assert len(want.events) >= 10 and np.array_equal(sub.events, want.events)
    assert np.array_equal(sub.x0, want.x0)
    # sticky: inclusion probabilities
    d = G.n
    kap = np.full(d, 0.7)
    cs = 4.0 * G.colnorms()
    Xs, _, _, _ = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 5.0, cs, Z, kap, seed=(3, 4))
    host = gpu.inclusion_prob(Xs)
    assert np.allclose(Xs.inclusion_prob, host, rtol=1e-12, atol=1e-15) and 0.05 < host.mean() < 0.95
    for rev, strong in ((True, False), (False, True), (True, True)):
        mode = O.PARITY_MODE | (O.STICKY_REVERSIBLE if rev else 0) | (O.STICKY_STRONG_UB if strong else 0)
        ref = O.spdmp(G, G, 0.0, x0, th0, 5.0, cs, kappa=kap, seed=(3, 4), mode=mode)
        Xo, (t, x, th), (a, n), _ = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 5.0, cs, Z, kap, seed=(3, 4), reversible=rev,
                                               strong_upperbounds=strong)
        assert n == ref.num and a == int(ref.acc.sum()) and np.array_equal(Xo.events["i"], ref.events["i"])
        for f in ("t", "x", "theta"):
            assert np.array_equal(Xo.events[f].view(np.uint64), ref.events[f].view(np.uint64))


def test_cummean_on_the_device(gpu):
    """cummean(trace) (src/trace.jl:203-225) computed on the device from the trace in HBM (zzb_trace_cummean: grouping by
    coordinate with the bucket kernels of the trace sort + one thread per coordinate) against the host routine on the copied
    trace, bit for bit; the host fallback (trace drained to the host, many events per coordinate) gives the same."""
    G, x0, th0, c = gpu.gmrf_config(40)
    prob = gpu.Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(G.n)))
    for T, host_sort in ((3.0, 0), (3.0, 1), (150.0, 0)):   # device path; trace ordered on the host; > 64 events per coordinate (host walk)
        run = gpu.Run(prob, record_trace=True)
        run.set(host_sort=host_sort)
        run.upload(0.5, x0, th0, c, seed=(3, 4))
        run.execute(0.5 + T)
        launches0 = run.stats()["launches"]
        off, times, values = run.cummean()
        kernels = run.stats()["launches"] - launches0
        ev = run.events()
        assert off[-1] == len(ev) == len(times)
        ref = gpu.cummean(gpu.FactTrace(None, 0.5, x0, th0, ev))
        got = gpu.cummean_lists(0.5, x0, off, times, values)
        for k in range(0, G.n, 7):
            assert np.array_equal(ref[k][0].view(np.uint64), got[k][0].view(np.uint64)), k
            assert np.array_equal(ref[k][1].view(np.uint64), got[k][1].view(np.uint64)), k
        assert kernels == (0 if host_sort else 4), (T, host_sort, kernels)   # (T = 150: the device tries, a coordinate overflows, the host walks)
        run.close()
    prob.close()
