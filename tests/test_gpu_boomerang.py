"""GPU parity for the factorised Boomerang (spdmp / pdmp with F::FactBoomerang, src/sfact.jl:29-48,73-145): the CUDA event
loop through the C-ABI (ZZB_FLAG_BOOMERANG, zzb_run_upload_boomerang) against the CPU oracle, bit for bit, plus the
reference's statistical acceptance test (test/maintest.jl:114-137) on the device path."""
import math

import numpy as np
import pytest

import oracle_lib as O
from test_boomerang import BOOM_CASES, boom_case, boom_inputs

pytestmark = pytest.mark.gpu


class R:
    pass


def run_gpu(zzb, G, Zg, sigma, mu, h, x0, th0, T, c, lref, rho, *, seed, adapt=False, tune=None, all_graph=False):
    F = zzb.FactBoomerang(Zg, np.zeros(G.n) if mu is None else mu, lref, sigma, rho=rho)
    f = zzb.pdmp if all_graph else zzb.spdmp
    Xi, (t, x, th), (acc, num), cc = f(zzb.GaussianPotential(G, h), 0.0, x0, th0, T, c, F, seed=seed, adapt=adapt, tune=tune)
    r = R()
    r.events, r.t, r.x, r.theta, r.c, r.acc, r.num = Xi.events, t, x, th, cc, acc, num
    return r, Xi


@pytest.mark.parametrize("case", BOOM_CASES, ids=lambda c: f"{c[0]}-{c[5]}-{c[6]}")
def test_boomerang_bit_exact(gpu, case):
    G, Zg, sigma, x0, th0, c, mu, h, T, lref, rho, adapt, opts = boom_case(gpu, case)
    ref = O.spdmp(G, Zg, 0.0, x0, th0, T, c, mu=mu, h=h, seed=(7, 8), mode=O.PARITY_MODE, boom=(sigma, lref, rho), adapt=adapt)
    tune = {k: v for k, v in opts.items() if k in ("delta0", "target_frac")}
    if "tag_limit" in opts and not (opts["tag_limit"] & 0x80000000):
        tune["tag_limit"] = opts["tag_limit"]
    got, Xi = run_gpu(gpu, G, Zg, sigma, mu, h, x0, th0, T, c, lref, rho, seed=(7, 8), adapt=adapt, tune=tune or None)
    O.assert_same_run(ref, got)
    assert Xi.moments is None


def test_boomerang_lattice_32(gpu):
    G = gpu.grid_precision(32, 32)
    rng = np.random.default_rng(9)
    Zg, sigma, x0, th0, c = boom_inputs(gpu, G, 1.0, rng)
    ref = O.spdmp(G, Zg, 0.0, x0, th0, 3.0, c, seed=(2, 3), mode=O.PARITY_MODE, boom=(sigma, 50.0, 0.1))
    got, _ = run_gpu(gpu, G, Zg, sigma, None, None, x0, th0, 3.0, c, 50.0, 0.1, seed=(2, 3))
    O.assert_same_run(ref, got)
    assert len(ref.events) > 1000


def test_sfactboomerang_moments_on_gpu(gpu):
    """test/maintest.jl:114-137 ("SFactBoomerang") on the device path: d = 8, Z = FactBoomerang(1.2 Gamma, 0, 0.3), T = 3000."""
    d, T = 8, 3000.0
    G = gpu.random_spd(d, seed=2)
    rng = np.random.default_rng(21)
    Zg, sigma, x0, th0, c = boom_inputs(gpu, G, 1.2, rng)
    F = gpu.FactBoomerang(Zg, np.zeros(d), 0.3)
    assert np.allclose(F.sigma, sigma)
    Xi, _, (acc, num), _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, F, seed=(5, 6))
    ts, xs = gpu.discretize(Xi, 0.5)
    Sigma = np.linalg.inv(G.to_scipy().toarray())
    assert np.mean(np.abs(xs.mean(axis=0))) < 2 / math.sqrt(T)
    assert np.mean(np.abs(np.cov(xs.T) - Sigma)) < 4 / math.sqrt(T)
    ref = O.spdmp(G, Zg, 0.0, x0, th0, T, c, seed=(5, 6), mode=O.PARITY_MODE, boom=(sigma, 0.3, 0.0))
    assert num == ref.num and np.array_equal(Xi.events["t"].view(np.uint64), ref.events["t"].view(np.uint64))


def test_boomerang_bound_violation_and_argument_errors(gpu):
    G = gpu.random_spd(8, seed=2)
    rng = np.random.default_rng(3)
    Zg, sigma, x0, th0, c = boom_inputs(gpu, G, 1.0, rng)
    with pytest.raises(gpu.BoundError, match="Tuning parameter `c` too small"):
        run_gpu(gpu, G, Zg, sigma, None, None, x0, th0, 500.0, 1e-3 * c, 0.3, 0.0, seed=(1, 1))
    with pytest.raises(gpu.ZZBError, match="lambdaref > 0"):
        run_gpu(gpu, G, Zg, sigma, None, None, x0, th0, 5.0, c, 0.0, 0.0, seed=(1, 1))


def test_fact_sampler_dispatches_on_boomerang_and_local_bound(gpu):
    """FactSampler(grad, u0, c, F) takes F::Union{ZigZag,FactBoomerang} (src/sfactiter.jl:5-40): with a FactBoomerang the pulled
    events are the Boomerang's (not a ZigZag run with F.Gamma as the bound); with LocalBound(c) they are those of
    spdmp(..., LocalBound(c), Z).  Same seed => the same events as the one-shot calls."""
    d, T = 8, 40.0
    G = gpu.random_spd(d, seed=2)
    rng = np.random.default_rng(5)
    Zg, sigma, x0, th0, c = boom_inputs(gpu, G, 1.2, rng)
    F = gpu.FactBoomerang(Zg, np.zeros(d), 0.7)
    Xi, _, _, _ = gpu.spdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, c, F, seed=(3, 9))
    FS = gpu.FactSampler(gpu.GaussianPotential(G), (0.0, (x0, th0)), c, F, seed=(3, 9), windows_per_pull=3)
    tr = gpu.trace(FS, T)
    k = len(tr.events)
    assert k > 50
    one = Xi.events[1:1 + k]                 # trace(FS, T) drops the first event (sfactiter.jl:68-70)
    assert np.array_equal(tr.events["i"], one["i"]) and np.array_equal(tr.events["t"].view(np.uint64), one["t"].view(np.uint64))
    assert np.array_equal(tr.events["theta"].view(np.uint64), one["theta"].view(np.uint64))
    # LocalBound through the iterator == LocalBound through spdmp
    Gl, xl, tl, _ = gpu.gmrf_config(12)
    Z = gpu.ZigZag(Gl.scaled(0.5), np.zeros(Gl.n))   # (F.Gamma is ignored by LocalBound: the bound comes from the target)
    cl = np.full(Gl.n, 1.0)
    Xl, _, _, _ = gpu.spdmp(gpu.GaussianPotential(Gl), 0.0, xl, tl, 3.0, gpu.LocalBound(cl.copy()), Z, seed=(4, 4))
    trl = gpu.trace(gpu.FactSampler(gpu.GaussianPotential(Gl), (0.0, (xl, tl)), gpu.LocalBound(cl.copy()), Z, seed=(4, 4)), 3.0)
    kl = len(trl.events)
    assert kl > 50 and np.array_equal(trl.events["t"].view(np.uint64), Xl.events["t"][1:1 + kl].view(np.uint64))
    with pytest.raises(TypeError, match="F::ZigZag only"):
        gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, 1.0, c, F, 1.0)
