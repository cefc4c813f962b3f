"""Sequential-chain schedule (zz_seq.cuh, zzb_run_set("schedule", 2)): one warp per connected component runs the event loop of
src/sfact.jl:73-145 as written.  Bit for bit against the CPU oracle (mode ctr|lazy) and against the windowed kernels."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_parity import run_gpu

pytestmark = pytest.mark.gpu

SEQ = dict(schedule=2)


def test_lattice_one_component(gpu):
    """Target and sampler matrix are the same object (the gradient's column sum is reused for the bound)."""
    G, x0, th0, c = gpu.gmrf_config(8)
    ref = O.spdmp(G, G, 0.0, x0, th0, 12.0, c)
    got, Xi = run_gpu(gpu, G, G, 0.0, x0, th0, 12.0, c, tune=SEQ)
    O.assert_same_run(ref, got)
    assert Xi.stats["windows"] == 0                      # no windows: the event loop ran sequentially
    m1, m2 = Xi.moments
    assert np.allclose(m1, ref.m1, rtol=1e-12, atol=1e-300) and np.allclose(m2, ref.m2, rtol=1e-12, atol=1e-300)
    win, _ = run_gpu(gpu, G, G, 0.0, x0, th0, 12.0, c, tune=dict(schedule=1))
    O.assert_same_run(win, got)


@pytest.mark.parametrize("seed", range(3))
def test_general_sparse_with_mu_h_adapt(gpu, seed):
    """Separate target matrix, Z.mu != 0, linear term, adaptation of c; columns of different lengths."""
    d = 60
    Gt = gpu.random_sparse_spd(d, deg=2 + seed % 3, seed=seed)
    Gb = Gt.scaled(0.9)
    rng = np.random.default_rng(seed)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
    mu, h = 0.1 * rng.standard_normal(d), 0.3 * rng.standard_normal(d)
    c = 0.2 * Gt.colnorms()
    ref = O.spdmp(Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True)
    got, _ = run_gpu(gpu, Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True, tune=SEQ)
    O.assert_same_run(ref, got)
    assert (got.c != c).any()


@pytest.mark.parametrize("warps", [1, 2, 4, 8])
def test_speculative_thinning_any_number_of_warps(gpu, warps):
    """The warps of a chain take the earliest queue entries side by side; an entry counts only if every earlier one of the step
    was rejected and re-queued later.  Same bits for 1, 2 and 4 warps: Gaussian chains (several components, tight and loose
    bounds: mostly accepted / mostly rejected proposals) and the logistic target."""
    import logistic_cases as LC
    G, x0, th0, c = gpu.gmrf_config(12)
    Gb = O.block_diagonal(G, 3)
    for scale, T in ((1.0, 8.0), (25.0, 3.0)):
        ref = O.spdmp(Gb, Gb, 0.0, x0, th0, T, scale * c)
        got, _ = run_gpu(gpu, Gb, Gb, 0.0, x0, th0, T, scale * c, tune=dict(schedule=2, seq_warps=warps))
        O.assert_same_run(ref, got)
    ct = np.full(Gb.n, np.sqrt(np.finfo(float).eps))
    ref = O.spdmp(Gb, Gb, 0.0, x0, th0, 4.0, ct)
    got, _ = run_gpu(gpu, Gb, Gb, 0.0, x0, th0, 4.0, ct, tune=dict(schedule=2, seq_warps=warps))
    O.assert_same_run(ref, got)
    *design, T = LC.SMALL[1]
    cfg = LC.make(gpu, *design)
    lref = LC.run_oracle(O, cfg, T)
    lgot, _ = LC.run_device(gpu, cfg, T, tune=dict(schedule=2, seq_warps=warps))
    O.assert_same_run(lref, lgot)


def test_dense_columns(gpu):
    """Columns longer than a warp: the cooperative column sums run over several chunks."""
    d = 70
    G = gpu.random_spd(d, density=0.9, seed=4)
    rng = np.random.default_rng(1)
    x0, th0 = rng.random(d), rng.choice(np.array([-1.0, 1.0]), d)
    c = 2.0 * G.colnorms()
    ref = O.spdmp(G, G, 0.0, x0, th0, 6.0, c, adapt=True)
    got, _ = run_gpu(gpu, G, G, 0.0, x0, th0, 6.0, c, adapt=True, tune=SEQ)
    O.assert_same_run(ref, got)


def test_block_diagonal_components_end_at_the_global_last_event(gpu):
    """Four independent chains: the run ends at the first accepted flip at or after T over ALL chains (sfact.jl:199); every chain
    has processed exactly the proposals before that time; the merged trace is time-ordered."""
    G, x0, th0, c = gpu.gmrf_config(12)
    Gb = O.block_diagonal(G, 4)
    for T in (0.7, 5.0):
        ref = O.spdmp(Gb, Gb, 0.0, x0, th0, T, c)
        got, Xi = run_gpu(gpu, Gb, Gb, 0.0, x0, th0, T, c, tune=SEQ)
        O.assert_same_run(ref, got)
        assert got.events["t"][-1] >= T and np.all(got.events["t"][:-1] < T)
        m1, m2 = Xi.moments
        assert np.allclose(m1, ref.m1, rtol=1e-12, atol=1e-300)


@pytest.mark.auto_schedule
def test_scattered_components_and_automatic_choice(gpu):
    """Components that are NOT contiguous index ranges (a symmetric permutation of a block-diagonal matrix): the host renumbers
    the coordinates chain by chain; draw streams, trace ids and the returned arrays keep the caller's numbering.  Small
    components make the sequential schedule the automatic choice."""
    from zzb200.problems import CSC
    G, x0, th0, c = gpu.gmrf_config(12)
    Gb = O.block_diagonal(G, 6)
    perm = np.random.default_rng(7).permutation(Gb.n)
    D = Gb.to_scipy().toarray()[np.ix_(perm, perm)]
    Gp = CSC.from_dense(D)
    x0, th0, c = x0[perm], th0[perm], c[perm]
    ref = O.spdmp(Gp, Gp, 0.0, x0, th0, 6.0, c)
    got, Xi = run_gpu(gpu, Gp, Gp, 0.0, x0, th0, 6.0, c, tune=SEQ)
    O.assert_same_run(ref, got)
    auto, Xa = run_gpu(gpu, Gp, Gp, 0.0, x0, th0, 6.0, c)           # components of 24 coordinates: sequential by default
    O.assert_same_run(ref, auto)
    assert Xa.stats["windows"] == 0
    win, Xw = run_gpu(gpu, Gp, Gp, 0.0, x0, th0, 6.0, c, tune=dict(schedule=1))
    O.assert_same_run(ref, win)
    assert Xw.stats["windows"] > 0
    # one large sparse component: the windowed relaxation stays the automatic choice
    G2, x2, t2, c2 = gpu.gmrf_config(16)
    _, X2 = run_gpu(gpu, G2, G2, 0.0, x2, t2, 1.0, c2)
    assert X2.stats["windows"] > 0


def test_trace_drain_continue_and_filter(gpu):
    """A trace buffer that holds a fraction of the events (drain and relaunch), two execute calls on one run (the chains resume
    from their saved state), the trace filter (subtrace at the source) and the host-side ordering."""
    G, x0, th0, c = gpu.gmrf_config(10)
    Gb = O.block_diagonal(G, 2)
    ref = O.spdmp(Gb, Gb, 0.0, x0, th0, 30.0, c)
    prob = gpu.Problem(gpu.GaussianPotential(Gb), gpu.ZigZag(Gb, np.zeros(Gb.n)))
    run = gpu.Run(prob, record_trace=True, trace_capacity=1)   # clamped to the minimum: 6 d + 2 records
    run.set(schedule=2)
    run.upload(0.0, x0, th0, c, seed=(1, 2))
    run.execute(11.0)
    n1 = len(run.events())
    run.execute(30.0)
    ev = run.events()
    assert 0 < n1 < len(ev) and run.stats()["launches"] > 8
    assert np.array_equal(ev["i"], ref.events["i"]) and np.array_equal(ev["t"].view(np.uint64), ref.events["t"].view(np.uint64))
    acc, num = run.counts()
    assert num == ref.num and np.array_equal(acc, ref.acc)
    t, x, th, cc = run.final_state()
    assert np.array_equal(x.view(np.uint64), ref.x.view(np.uint64)) and np.array_equal(t.view(np.uint64), ref.t.view(np.uint64))
    run.close()
    J = np.array([3, 4, 50, 99], dtype=np.int64)
    Xi, _, (acc2, num2), _ = gpu.spdmp(gpu.GaussianPotential(Gb), 0.0, x0, th0, 30.0, c, gpu.ZigZag(Gb, np.zeros(Gb.n)), seed=(1, 2),
                                      tune=SEQ, trace_filter=J)
    sel = ref.events[np.isin(ref.events["i"], J)]
    assert num2 == ref.num and len(Xi.events) == len(sel)
    assert np.array_equal(Xi.events["t"].view(np.uint64), sel["t"].view(np.uint64))
    assert np.array_equal(Xi.events["i"], np.searchsorted(J, sel["i"]) + 1)
    # ordering on the host instead of the device: same trace
    run = gpu.Run(prob, record_trace=True)
    run.set(schedule=2, host_sort=1)
    run.upload(0.0, x0, th0, c, seed=(1, 2))
    run.execute(30.0)
    ev2 = run.events()
    assert np.array_equal(ev2["t"].view(np.uint64), ref.events["t"].view(np.uint64)) and np.array_equal(ev2["i"], ref.events["i"])
    run.close()
    prob.close()


def test_bound_violation_and_refusals(gpu):
    G, x0, th0, c = gpu.gmrf_config(6)
    Gb = G.scaled(0.5)   # a sampler matrix that underestimates the rate, adapt = false: error of sfact.jl:124 on both sides
    with pytest.raises(O.BoundError):
        O.spdmp(G, Gb, 0.0, x0, th0, 50.0, 1e-3 * c)
    with pytest.raises(gpu.BoundError, match="Tuning parameter `c` too small"):
        run_gpu(gpu, G, Gb, 0.0, x0, th0, 50.0, 1e-3 * c, tune=SEQ)
    # a connected component larger than the shared memory of an SM; samplers other than the plain ZigZag
    Gbig, xb, tb, cb = gpu.gmrf_config(64)
    with pytest.raises(gpu.ZZBError, match="sequential"):
        run_gpu(gpu, Gbig, Gbig, 0.0, xb, tb, 0.1, cb, tune=SEQ)
    prob = gpu.Problem(gpu.GaussianPotential(G), gpu.ZigZag(G, np.zeros(G.n)))
    for kw in (dict(local_bound=True), dict(kappa=np.ones(G.n))):
        run = gpu.Run(prob, **kw)
        with pytest.raises(gpu.ZZBError, match="sequential"):
            run.set(schedule=2)
        run.close()
    prob.close()
