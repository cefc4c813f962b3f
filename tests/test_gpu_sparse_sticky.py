"""Config 4 on the device against the reference's own algorithm for it: the sticky kernel (contract: src/ss_fact.jl, bit-exact
against zzo_sspdmp elsewhere) must show the occupancy statistics of the CPU restatement of src/sparsestickyzz.jl."""
import numpy as np
import pytest

import oracle_lib as O
from sticky_stats import chain_precision, occupancy

pytestmark = pytest.mark.gpu


def test_device_sticky_run_has_the_law_of_sparsestickyzz(gpu):
    p, kappa, T = 2000, 0.5, 150.0
    G = chain_precision(gpu, p)
    x0 = np.zeros(p)
    ref = O.sparsestickyzz(G, x0, np.ones(p), T, 2.5, kappa, rule="sticky", seed=(3, 4))
    th0 = np.random.default_rng(0).choice(np.array([-1.0, 1.0]), p)
    Xi, _, (acc, num), _ = gpu.sspdmp(gpu.GaussianPotential(G), 0.0, x0, th0, T, G.colnorms(), gpu.ZigZag(G, np.zeros(p)),
                                      np.full(p, kappa), seed=(5, 6))
    oa, ma = occupancy(ref.events, p, x0, T)
    ob, mb = occupancy(Xi.events, p, x0, T)
    # measured on the CPU with the device's contract (zzo_sspdmp, ctr|lazy, same seed): 0.5100 vs 0.5107, 0.379 vs 0.389
    assert abs(oa.mean() - ob.mean()) < 0.01
    assert abs(ma.mean() - mb.mean()) < 0.06 * mb.mean()
    assert abs(len(ref.events) - len(Xi.events)) < 0.03 * len(ref.events)
