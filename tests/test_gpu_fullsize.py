"""Full-size parity: the device against the CPU oracle (contract mode ctr | lazy), bit for bit, at the sizes BASELINE.json names
-- not only through size-independent properties.  The oracle needs a few seconds per case on one host thread."""
import numpy as np
import pytest

import oracle_lib as O
from sticky_stats import chain_precision

pytestmark = pytest.mark.gpu


class R:
    pass


def device_run(zzb, G, x0, th0, T, c, seed, *, kappa=None, tune=None):
    Z = zzb.ZigZag(G, np.zeros(G.n))
    if kappa is None:
        Xi, (t, x, th), (acc, num), cc = zzb.spdmp(zzb.GaussianPotential(G), 0.0, x0, th0, T, c, Z, seed=seed, tune=tune)
    else:
        Xi, (t, x, th), (acc, num), cc = zzb.sspdmp(zzb.GaussianPotential(G), 0.0, x0, th0, T, c, Z, kappa, seed=seed, tune=tune)
    r = R()
    r.events, r.t, r.x, r.theta, r.c, r.acc, r.num = Xi.events, t, x, th, cc, acc, num
    if kappa is not None:
        r.acc = Xi.acc_per_coordinate   # (sspdmp returns the total like the reference; the oracle counts per coordinate)
    return r


def test_config5_d1e6_bit_exact(gpu):
    """BASELINE configs[4] on one GPU: d = 10^6 lattice GMRF, c = ||Gamma[:,i]||, T = 0.5 (3.4e5 switches, 1.9e6 proposals)."""
    G, x0, th0, c = gpu.gmrf_config(1000)
    ref = O.spdmp(G, G, 0.0, x0, th0, 0.5, c, seed=(3, 4))
    got = device_run(gpu, G, x0, th0, 0.5, c, (3, 4))
    O.assert_same_run(ref, got)
    assert ref.num > 1_500_000 and len(ref.events) > 250_000


def test_config5_d1e6_tight_bound_bit_exact(gpu):
    """The same lattice with the tight bound c = sqrt(eps) of scripts/example.jl:39 (acceptance ~ 1)."""
    G, x0, th0, c = gpu.gmrf_config(1000, tight=True)
    ref = O.spdmp(G, G, 0.0, x0, th0, 0.5, c, seed=(5, 6))
    got = device_run(gpu, G, x0, th0, 0.5, c, (5, 6))
    O.assert_same_run(ref, got)


@pytest.mark.parametrize("tight", [False, True])
def test_config2_d1e4_T20_bit_exact(gpu, tight):
    """BASELINE configs[1] as SURVEY 8(d) specifies it: n = 100 (d = 10^4), T = 20 (~1.5e5 switches, 8e5 proposals), both bounds."""
    G, x0, th0, c = gpu.gmrf_config(100, tight=tight)
    ref = O.spdmp(G, G, 0.0, x0, th0, 20.0, c, seed=(1, 2))
    got = device_run(gpu, G, x0, th0, 20.0, c, (1, 2))
    O.assert_same_run(ref, got)
    assert len(ref.events) > 100_000


def test_config4_sticky_chain_p1e5_bit_exact(gpu):
    """BASELINE configs[3] at full size on the ss_fact.jl kernel: p = 10^5 chain, x0 = 0, kappa = 2000/p, c = 2.5, T = 3."""
    p, T = 100000, 3.0
    G = chain_precision(gpu, p)
    rng = np.random.default_rng(0)
    x0, th0 = np.zeros(p), rng.choice(np.array([-1.0, 1.0]), p)
    c, kappa = np.full(p, 2.5), np.full(p, 2000.0 / p)
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c, kappa=kappa, seed=(5, 6))
    got = device_run(gpu, G, x0, th0, T, c, (5, 6), kappa=kappa)
    O.assert_same_run(ref, got)
    assert len(ref.events) > 50_000
