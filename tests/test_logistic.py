"""Config 3 (sparse logistic regression, subsampled ZigZag; scripts/logistic.jl) on the CPU: the oracle's restatement, the
shared exponential, and the host emulation of the device schedule (same per-coordinate code as the kernel, zz_logit.h)
against the sequential oracle, bit for bit."""
import math

import numpy as np
import pytest

import logistic_cases as LC
import oracle_lib as O


def test_exp_within_one_ulp_of_libm():
    L = O.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-40, 40, 20000), rng.uniform(-1, 1, 5000), rng.uniform(-700, 700, 3000),
                         [0.0, 1e-10, -1e-10, 0.34657359027997264, -0.35, 709.0, -745.0, 800.0, -800.0]])
    worst = 0.0
    for x in xs:
        a, b = L.zzo_exp(float(x)), math.exp(x) if x < 709.78 else math.inf
        if b == 0.0 or math.isinf(b) or b < 2.3e-308:
            assert a == pytest.approx(b, abs=1e-307)
            continue
        worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0


def test_design_matches_the_reference_construction(zzb):
    """scripts/sparsedesign.jl:1-25 and scripts/logistic.jl:21-31: p = sum(d) + pairwise interactions + r, n = m p, at most
    one level per factor switched on, interaction = 0.3 * (both levels on), At = A'."""
    A = zzb.sparse_design((3, 4), 2, 6, np.random.default_rng(1))
    n, p = A.shape
    assert p == 3 + 4 + 12 + 2 and n == 6 * p
    assert np.all((A[:, :3] != 0).sum(1) <= 1) and np.all((A[:, 3:7] != 0).sum(1) <= 1)
    for c2 in range(3):          # CartesianIndices((d[2], d[1])): the level of the later factor runs fastest
        for c1 in range(4):
            col = 7 + c2 * 4 + c1
            assert np.array_equal(A[:, col], 0.3 * ((A[:, 3 + c1] == 1) & (A[:, c2] == 1)))
    S = zzb.RectCSC.from_dense(A)
    assert np.array_equal(S.to_dense(), A) and np.array_equal(S.transpose().to_dense(), A.T)
    assert np.all(np.diff(S.transpose().rowval[: S.transpose().colptr[1] - 1]) > 0)


def test_full_config_shape_and_mode(zzb):
    """README.md:50 / BASELINE config 3: n = 8840, p = 442; mu is the mode (full gradient ~ 0 there, scripts/logistic.jl:102,120-125)."""
    cfg = LC.make(zzb, *LC.FULL)
    assert (cfg["n"], cfg["p"]) == (8840, 442)
    lg = cfg["logistic"]
    grad = zzb.LogisticSubsampled(lg["A"], lg["At"], lg["y"], lg["ny"], lg["mu"], lg["gamma0"], lg["k"])
    for i in (1, 40, 41, 300, 441, 442):
        assert abs(grad(cfg["mu"], i)) < 1e-6
    Gd = cfg["Gamma_drop"]
    assert np.array_equal(Gd.to_scipy().toarray(), Gd.to_scipy().toarray().T) and Gd.nnz < cfg["Gamma"].nnz


@pytest.mark.parametrize("case", LC.SMALL)
def test_schedule_emulation_equals_sequential_oracle(zzb, case):
    """The windowed relaxation with the device's per-coordinate code (zz_process_node_logit) reproduces the sequential event
    loop bit for bit, whatever the window policy."""
    *design, T = case
    cfg = LC.make(zzb, *design)
    ref = LC.run_oracle(O, cfg, T)
    assert len(ref.events) > 50 and (ref.c != cfg["c"]).any()
    for kw in (dict(), dict(delta0=1e-3, target_frac=0.1), dict(delta0=2.0, target_frac=4.0, tag_limit=40)):
        sim = O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], T, cfg["c"], mu=cfg["mu"], adapt=True, factor=5.0,
                           logistic=cfg["logistic"], seed=(5, 6), **kw)
        O.assert_same_run(ref, sim)


def test_full_size_emulation_equals_oracle(zzb):
    cfg = LC.make(zzb, *LC.FULL)
    ref = LC.run_oracle(O, cfg, 6.0)
    sim = O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], 6.0, cfg["c"], mu=cfg["mu"], adapt=True, factor=5.0,
                       logistic=cfg["logistic"], seed=(5, 6))
    assert len(ref.events) > 1000
    O.assert_same_run(ref, sim)


def test_bound_violation_without_adapt(zzb):
    """c = 0.01 is far too small at the start (the script relies on adapt = true): adapt = false must raise (sfact.jl:124)."""
    cfg = LC.make(zzb, *LC.SMALL[0][:4])
    with pytest.raises(O.BoundError):
        LC.run_oracle(O, cfg, 20.0, adapt=False)
    with pytest.raises(O.BoundError):
        O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], 20.0, cfg["c"], mu=cfg["mu"], adapt=False,
                     logistic=cfg["logistic"], seed=(5, 6))


def test_modes_agree_in_law_and_sample_the_posterior(zzb):
    """Faithful draw order / in-place moves (seq|inplace) and the parity contract (ctr|lazy) are the same process in law; the
    time averages sit near the mode and the marginal variances agree with the Laplace approximation sigma^2 = diag(inv(Gamma)) (scripts/logistic.jl:150)."""
    cfg = LC.make(zzb, (4, 4), 2, 40, 11)
    T = 3000.0
    out = []
    for mode in (O.PARITY_MODE, O.RNG_SEQ | O.ARITH_INPLACE):
        r = LC.run_oracle(O, cfg, T, mode=mode)
        z = (r.m1 - cfg["mu"]) / cfg["sigma"]
        v = (r.m2 - r.m1 ** 2) / cfg["sigma"] ** 2
        out.append((z, v, len(r.events) / r.num))
        # the posterior is skewed for sparsely observed columns: mean and mode differ by up to ~1.5 sigma there, the same in
        # every mode and for every T (measured: max 1.5, mean 0.28, variance ratio 1.03)
        assert np.abs(z).max() < 2.5 and np.abs(z).mean() < 0.5
        assert 0.8 < np.median(v) < 1.3
    assert abs(out[0][2] - out[1][2]) < 0.01            # acceptance ratios
    assert np.abs(out[0][0] - out[1][0]).max() < 0.35   # Monte Carlo error between two independent runs (measured 0.13)


def test_lazy_and_inplace_arithmetic_agree(zzb):
    """Same counter streams, positions advanced in place (reference arithmetic, idot_moving!) vs flip-anchored: same events,
    times equal to rounding."""
    cfg = LC.make(zzb, *LC.SMALL[1][:4])
    a = LC.run_oracle(O, cfg, 6.0, mode=O.RNG_CTR | O.ARITH_LAZY)
    b = LC.run_oracle(O, cfg, 6.0, mode=O.RNG_CTR | O.ARITH_INPLACE)
    n = min(len(a.events), len(b.events), 60)
    assert n >= 40 and np.array_equal(a.events["i"][:n], b.events["i"][:n])
    assert np.allclose(a.events["t"][:n], b.events["t"][:n], rtol=1e-9)


def test_replicas_are_independent_chains_in_one_problem(zzb):
    """Block-diagonal replication (config 3 is "plumbing + replicas"): the schedule emulation equals the oracle on the joint
    problem, the replicas do not interact (each one's event count is that of a chain of its own) and they differ from
    each other (own counter streams)."""
    *design, T = LC.SMALL[0]
    cfg = LC.make(zzb, *design)
    R, p = 3, cfg["p"]
    big = LC.replicas(zzb, cfg, R)
    assert big["A"].ncols == R * p and big["Gamma_drop"].n == R * p
    ref = LC.run_oracle(O, big, T)
    sim = O.window_sim(None, big["Gamma_drop"], 0.0, big["x0"], big["theta0"], T, big["c"], mu=big["mu"], adapt=True, factor=5.0,
                       logistic=big["logistic"], seed=(5, 6))
    O.assert_same_run(ref, sim)
    one = LC.run_oracle(O, cfg, T)
    block = (ref.events["i"] - 1) // p
    first = ref.events[block == 0]
    # replica 0 has the coordinate ids (hence the streams) of the single chain: identical events up to the stopping rule
    n = min(len(first), len(one.events)) - 1
    assert n > 50 and np.array_equal(first["i"][:n], one.events["i"][:n]) and np.array_equal(first["t"][:n], one.events["t"][:n])
    counts = np.bincount(block, minlength=R)
    assert counts.min() > 0.6 * counts.max() and len({tuple(ref.events["t"][block == r][:5]) for r in range(R)}) == R


def test_random_designs_emulation_equals_oracle(zzb):
    """A slice of tools/fuzz_logistic.py: random designs / subsample counts / bounds / window policies."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_logistic", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "fuzz_logistic.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    n_ok, n_err = fz.run(zzb, 7, 40, verbose=False)
    assert n_ok >= 20 and n_err >= 1


def test_host_preparation_rejects_malformed_designs(zzb):
    """zz_build_logit (shared by libzzb200.so and the schedule emulation): At must be the transpose of A, every column of A
    needs an entry (upstream: rand over an empty range throws), k >= 1."""
    cfg = LC.make(zzb, *LC.SMALL[0][:4])
    lg = dict(cfg["logistic"])

    def sim(**over):
        l2 = dict(lg)
        l2.update(over)
        return O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], 1.0, cfg["c"], mu=cfg["mu"], adapt=True,
                            logistic=l2, seed=(1, 2))
    sim()
    At = lg["At"]
    bad_vals = zzb.RectCSC(At.nrows, At.ncols, At.colptr, At.rowval, At.nzval * 1.5)
    with pytest.raises(RuntimeError, match="status 4"):
        sim(At=bad_vals)
    with pytest.raises(RuntimeError, match="status 4"):
        sim(k=0)
    A = lg["A"]
    dense = A.to_dense()
    dense[:, 3] = 0.0                                  # an empty column
    A0 = zzb.RectCSC.from_dense(dense)
    with pytest.raises(RuntimeError, match="status 4"):
        sim(A=A0, At=A0.transpose())
    with pytest.raises(RuntimeError):                  # the oracle refuses it as well
        O.spdmp(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], 1.0, cfg["c"], mu=cfg["mu"], adapt=True,
                logistic=dict(lg, A=A0, At=A0.transpose()))


def test_logistic_golden_fixture(zzb):
    """tests/golden/logistic26.json pins our RNG / arithmetic contract for the logistic target across rounds."""
    import json
    import os

    import golden_cases as GC
    for name in GC.LOGISTIC_CASES:
        g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", name + ".json")))
        assert GC.run_oracle(O, GC.case_inputs(zzb, name)) == g


def test_columns_longer_than_the_gather_scratch_fall_back_to_list_walking(zzb):
    """No sparsification (droptol = 0): the dense regressors couple to all p = 257 > ZZ_LNB = 192 coordinates, so their
    timelines run in the list-walking version next to gather-first ones for the short columns -- same bits."""
    cfg = zzb.logistic_config(levels=(15, 15), r=2, m=12, seed=4, droptol=0.0)
    assert np.diff(cfg["A"].colptr).min() >= 1 and np.diff(cfg["Gamma_drop"].colptr).max() > 192
    cfg["logistic"] = dict(A=cfg["A"], At=cfg["At"], y=cfg["y"], ny=cfg["ny"], mu=cfg["mu"], gamma0=cfg["gamma0"], k=10)
    ref = LC.run_oracle(O, cfg, 2.0)
    sim = O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], 2.0, cfg["c"], mu=cfg["mu"], adapt=True, factor=5.0,
                       logistic=cfg["logistic"], seed=(5, 6))
    assert len(ref.events) > 100
    O.assert_same_run(ref, sim)
