"""The restatement of the reference's multithreaded parallel_spdmp (src/parallel.jl; oracle/zz_oracle.c
zzo_parallel_spdmp) -- the CPU baseline bench.py runs on all host cores -- against the reference's own test
(test/testparallel.jl:21-73) and a few structural properties."""
import numpy as np
import pytest

import oracle_lib as O


def tridiag(zzb, d):
    import scipy.sparse as sp
    M = sp.diags([np.ones(d), -0.4 * np.ones(d - 1), -0.4 * np.ones(d - 1)], [0, 1, -1]).tocsc()   # testparallel.jl:30
    return zzb.CSC.from_scipy(M), M.toarray()


def test_partition_blocks(zzb):
    """Partition(2, 6) puts 1:3 / 4:6 into chunks 1 / 2 (test/testparallel.jl:4-12): block_diagonal keeps exactly the
    entries inside those chunks."""
    G, M = tridiag(zzb, 6)
    G2 = O.block_diagonal(G, 2).to_scipy().toarray()
    want = M.copy()
    want[:3, 3:] = 0
    want[3:, :3] = 0
    assert np.array_equal(G2, want)


@pytest.mark.parametrize("K", [1, 2, 4])
def test_reference_parallel_zigzag_test(zzb, K):
    """test/testparallel.jl:21-64: d = 20, tridiagonal precision, bound matrix without the cross-chunk entries,
    c = 5 ||Gamma[:, i]||, Delta = 0.05, T = 1000."""
    d, T = 20, 1000.0
    G, M = tridiag(zzb, d)
    G2 = O.block_diagonal(G, K)
    rng = np.random.default_rng(1)
    x0, th0 = 0.1 * rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c = 5 * G.colnorms()
    r = O.spdmp(G, G2, 0.0, x0, th0, T, c, seed=(1, 2), parallel=(K, 0.05))
    assert 0.1 / np.sqrt(T) < np.abs(r.m1).mean() < 4 / np.sqrt(T)                    # :59
    ts, xs = zzb.discretize(zzb.FactTrace(None, 0.0, x0, th0, r.events), 0.5)
    assert np.abs(np.cov(xs.T) - np.linalg.inv(M)).mean() < 4 / np.sqrt(T)            # :63-64
    assert np.all(np.diff(r.events["t"]) >= 0) and len(r.events) == r.acc.sum() and r.num > len(r.events)
    assert r.loop_seconds > 0


def test_bound_across_chunks_is_refused(zzb):
    """error("Upper bounds may not depend across chunks.") (src/parallel.jl:126-129)."""
    G, _ = tridiag(zzb, 20)
    with pytest.raises(RuntimeError, match="status 4"):
        O.spdmp(G, G, 0.0, np.zeros(20), np.ones(20), 1.0, G.colnorms(), parallel=(2, 0.05))


def test_lattice_same_law_as_sequential(zzb):
    """On the lattice GMRF the multithreaded sampler and the sequential spdmp agree in their event and proposal rates."""
    G, x0, th0, c = zzb.gmrf_config(60)
    a = O.spdmp(G, G, 0.0, x0, th0, 3.0, c, seed=(1, 2), mode=O.RNG_SEQ | O.ARITH_INPLACE)
    b = O.spdmp(G, O.block_diagonal(G, 4), 0.0, x0, th0, 3.0, c, seed=(1, 2), parallel=(4, 0.02), adapt=True)
    assert abs(len(a.events) - len(b.events)) < 0.03 * len(a.events)
    assert abs(a.num - b.num) < 0.03 * a.num
