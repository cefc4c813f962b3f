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
