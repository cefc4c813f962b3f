"""The windowed relaxation schedule (host emulation built on the kernel's own per-coordinate code, zz_core.h /
zz_fast.h / zz_ctl.h) against the sequential oracle: bit-exact for every graph kind the kernels specialise on."""
import numpy as np
import pytest

import oracle_lib as O

FORCE_CSR = 0x80000000  # top bit of tag_limit: disable the lattice fast path in the emulation


@pytest.mark.parametrize("n,T", [(3, 10.0), (9, 5.0), (32, 3.0)])
@pytest.mark.parametrize("tag_limit", [0x0F000000, 0x0F000000 | FORCE_CSR, 40])
def test_grid(zzb, n, T, tag_limit):
    G, x0, th0, c = zzb.gmrf_config(n)
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c)
    got = O.window_sim(G, G, 0.0, x0, th0, T, c, tag_limit=tag_limit)
    O.assert_same_run(ref, got)
    assert np.allclose(got.s1, ref.s1, rtol=0, atol=0) and got.stats["windows"] > 1


@pytest.mark.parametrize("delta0,frac", [(1e-4, 0.02), (0.01, 0.4), (2.0, 5.0)])
def test_window_length_is_irrelevant(zzb, delta0, frac):
    G, x0, th0, c = zzb.gmrf_config(20)
    ref = O.spdmp(G, G, 0.0, x0, th0, 4.0, c)
    got = O.window_sim(G, G, 0.0, x0, th0, 4.0, c, delta0=delta0, target_frac=frac)
    O.assert_same_run(ref, got)


def test_rectangular_grid_and_tight_bound(zzb):
    G = zzb.grid_precision(5, 9)
    rng = np.random.default_rng(3)
    x0, th0 = rng.standard_normal(45), rng.choice(np.array([-1.0, 1.0]), 45)
    for c in (G.colnorms(), np.full(45, np.sqrt(np.finfo(float).eps))):
        ref = O.spdmp(G, G, 0.0, x0, th0, 8.0, c)
        O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, 8.0, c))


@pytest.mark.parametrize("seed", range(5))
def test_random_sparse_with_mu_h_adapt(zzb, seed):
    d = 40
    Gt = zzb.random_sparse_spd(d, deg=2 + seed % 3, seed=seed)
    Gb = Gt.scaled(0.9)
    rng = np.random.default_rng(seed)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
    mu, h = 0.1 * rng.standard_normal(d), 0.3 * rng.standard_normal(d)
    c = 0.2 * Gt.colnorms()
    ref = O.spdmp(Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True)
    got = O.window_sim(Gt, Gb, 0.0, x0, th0, 20.0, c, h=h, mu=mu, adapt=True)
    O.assert_same_run(ref, got)


def test_dense_columns_take_the_slow_path(zzb):
    G = zzb.random_spd(12, density=0.5)
    assert np.diff(G.colptr).max() > 8
    rng = np.random.default_rng(0)
    x0, th0 = rng.random(12), rng.choice(np.array([-1.0, 1.0]), 12)
    ref = O.spdmp(G, G, 0.0, x0, th0, 50.0, 2.0 * G.colnorms(), adapt=True)
    O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, 50.0, 2.0 * G.colnorms(), adapt=True))


def test_bound_violation_and_t0(zzb):
    G, x0, th0, c = zzb.gmrf_config(8)
    with pytest.raises(O.BoundError):
        O.window_sim(G, G.scaled(0.5), 0.0, x0, th0, 5.0, np.full(G.n, 1e-9))
    ref = O.spdmp(G, G, 0.0, x0, th0, 0.0, c)     # T <= t0
    got = O.window_sim(G, G, 0.0, x0, th0, 0.0, c)
    assert len(got.events) == 0 and got.num == ref.num == 0


@pytest.mark.parametrize("cval", [1.0, 0.05])
def test_local_bound(zzb, cval):
    """spdmp(..., C::LocalBound, ...) (src/local.jl): bound expiry / renew items in the timelines."""
    LB = O.PARITY_MODE | O.LOCAL_BOUND
    G, x0, th0, _ = zzb.gmrf_config(16)
    c = np.full(G.n, cval)
    ref = O.spdmp(G, G, 0.0, x0, th0, 5.0, c, mode=LB, adapt=True)
    for tl in (0x0F000000, 0x0F000000 | FORCE_CSR):
        O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, 5.0, c, tag_limit=tl, local_bound=True, adapt=True))
    d = 40
    Gt = zzb.random_sparse_spd(d, deg=4, seed=2)
    rng = np.random.default_rng(2)
    x0, th0, h = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d), 0.3 * rng.standard_normal(d)
    ref = O.spdmp(Gt, Gt, 0.0, x0, th0, 15.0, np.full(d, 0.3), h=h, mode=LB, adapt=True)
    O.assert_same_run(ref, O.window_sim(Gt, Gt, 0.0, x0, th0, 15.0, np.full(d, 0.3), h=h, local_bound=True, adapt=True))


@pytest.mark.parametrize("n,T,kap", [(6, 8.0, 1.0), (16, 5.0, 0.5), (24, 3.0, 5.0)])
def test_sticky(zzb, n, T, kap):
    """sspdmp (src/ss_fact.jl): freeze / thaw items in the timelines, (time, velocity) lists."""
    G, x0, th0, c = zzb.gmrf_config(n)
    kappa = np.full(G.n, kap)
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c, kappa=kappa)
    assert (ref.events["theta"] == 0).sum() > 10
    for tl in (0x0F000000, 0x0F000000 | FORCE_CSR, 60):
        O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, T, c, tag_limit=tl, kappa=kappa))


def test_sticky_general_graph(zzb):
    d = 40
    Gt = zzb.random_sparse_spd(d, deg=2, seed=0)
    assert np.diff(Gt.colptr).max() <= 8
    rng = np.random.default_rng(0)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
    h, mu, kappa = 0.3 * rng.standard_normal(d), 0.1 * rng.standard_normal(d), rng.uniform(0.2, 3.0, d)
    c = 2.0 * Gt.colnorms() + 1
    ref = O.spdmp(Gt, Gt.scaled(0.9), 0.0, x0, th0, 15.0, c, h=h, mu=mu, kappa=kappa)
    O.assert_same_run(ref, O.window_sim(Gt, Gt.scaled(0.9), 0.0, x0, th0, 15.0, c, h=h, mu=mu, kappa=kappa))


def test_lattice_column_multiply_shift_is_exact():
    """zz_grid_col (j * magic >> shift) equals j // M for every coordinate the formula may see."""
    import ctypes as C
    L = O.wlib()
    L.zzw_check_grid_col.restype = C.c_int64
    L.zzw_check_grid_col.argtypes = [C.c_int32, C.c_int64]
    for M in (1, 2, 3, 5, 7, 12, 100, 255, 256, 257, 1000, 1023, 1024, 1025, 4097, 46341, 65535, 65536, 1 << 20, (1 << 31) - 1):
        assert L.zzw_check_grid_col(M, 3_000_000) == 0, M


def test_random_campaign_emulation_equals_oracle(zzb):
    """The randomised campaign of tests/fuzz_cases.py (random lattices / sparse graphs, ZigZag / LocalBound / sticky / Boomerang,
    mu, h, scaled sampler matrix, window policies, bounds that are violated) with the host emulation of the device schedule in
    place of the CUDA path: the kernels' per-coordinate code against the sequential oracle, bit for bit, without a GPU."""
    import fuzz_cases
    bad, nbound = fuzz_cases.run_cases(zzb, 150, 4, sim=True)
    assert bad == 0


def test_bound_matrix_without_stored_diagonal_is_refused(zzb):
    """The reference reschedules i after its own flip only because i is in G1[i] (src/sfact.jl:131-135,170); the graph builder
    shared by the product library and this emulation (csrc/zz_host_graph.h) refuses a sampler matrix whose column lacks the
    diagonal entry instead of silently sampling a different process (the C-ABI maps the message to ZZB_E_GRAPH)."""
    import scipy.sparse as sp
    d = 6
    G = zzb.CSC.from_scipy(sp.diags([np.full(d, 2.0), -np.ones(d - 1), -np.ones(d - 1)], [0, 1, -1]).tocsc())
    offdiag = sp.diags([-np.ones(d - 1), -np.ones(d - 1)], [1, -1]).tocsc()
    Gb = zzb.CSC.from_scipy(offdiag)
    x0, th0, c = np.zeros(d), np.ones(d), np.ones(d)
    with pytest.raises(RuntimeError, match="status 4"):
        O.window_sim(G, Gb, 0.0, x0, th0, 1.0, c)
    O.window_sim(G, G, 0.0, x0, th0, 1.0, c)   # with the diagonal stored it runs


# ---------------------------------------------------------------------------------------------------------------------
# The ASYNCHRONOUS tile-local relaxation of round 2 (zz_run_body_async), emulated on the host with the device's own
# per-coordinate code: per-tile queues (lattice: checkerboard colours), dedupe bits, inboxes with delayed delivery, publisher-owned
# list tags, every evaluation with the freshest lists.  Any interleaving must give the oracle's bits.
@pytest.mark.parametrize("tiles,order_seed", [(1, 1), (3, 2), (7, 3), (16, 4), (16, 5), (37, 6)])
def test_async_schedule_lattice(zzb, tiles, order_seed):
    G, x0, th0, c = zzb.gmrf_config(16)
    ref = O.spdmp(G, G, 0.0, x0, th0, 4.0, c)
    got = O.window_sim(G, G, 0.0, x0, th0, 4.0, c, async_tiles=tiles, order_seed=order_seed)
    O.assert_same_run(ref, got)


@pytest.mark.parametrize("seed", range(6))
def test_async_schedule_random_sparse_adapt(zzb, seed):
    d = 50
    Gt = zzb.random_sparse_spd(d, deg=2 + seed % 3, seed=seed)
    Gb = Gt.scaled(0.9)
    rng = np.random.default_rng(seed)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
    h, mu = 0.3 * rng.standard_normal(d), 0.1 * rng.standard_normal(d)
    c = 0.2 * Gt.colnorms()
    ref = O.spdmp(Gt, Gb, 0.0, x0, th0, 10.0, c, h=h, mu=mu, adapt=True)
    got = O.window_sim(Gt, Gb, 0.0, x0, th0, 10.0, c, h=h, mu=mu, adapt=True, async_tiles=1 + seed * 3, order_seed=100 + seed,
                       target_frac=[0.05, 0.4, 2.0][seed % 3])
    O.assert_same_run(ref, got)


def test_async_schedule_other_samplers(zzb):
    """LocalBound, sticky and Boomerang timelines under the asynchronous schedule."""
    G, x0, th0, _ = zzb.gmrf_config(10)
    c = np.full(G.n, 1.0)
    LB = O.PARITY_MODE | O.LOCAL_BOUND
    ref = O.spdmp(G, G, 0.0, x0, th0, 3.0, c, mode=LB)
    O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, 3.0, c, local_bound=True, async_tiles=5, order_seed=7))
    kap = np.full(G.n, 0.7)
    c2 = np.full(G.n, 5.0)
    ref = O.spdmp(G, G, 0.0, x0, th0, 3.0, c2, kappa=kap)
    O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, th0, 3.0, c2, kappa=kap, async_tiles=4, order_seed=8))
    rng = np.random.default_rng(3)
    sigma = 1.0 / np.sqrt(np.array([G.to_scipy()[i, i] for i in range(G.n)]))
    thb = sigma * rng.standard_normal(G.n)
    cb = np.full(G.n, 8.0)
    ref = O.spdmp(G, G, 0.0, x0, thb, 2.0, cb, boom=(sigma, 20.0, 0.2), mode=O.PARITY_MODE)
    O.assert_same_run(ref, O.window_sim(G, G, 0.0, x0, thb, 2.0, cb, boom=(sigma, 20.0, 0.2), async_tiles=6, order_seed=9))


def test_async_schedule_randomised_campaign(zzb):
    """The randomised campaign of tests/fuzz_cases.py under the asynchronous schedule (random tile counts and interleavings)."""
    import fuzz_cases
    bad, nbound = fuzz_cases.run_cases(zzb, 120, 11, sim=True, async_sim=True)
    assert bad == 0


@pytest.mark.parametrize("reversible,strong", [(True, False), (False, True), (True, True)])
def test_sticky_options_reversible_and_strong_upperbounds(zzb, reversible, strong):
    """sspdmp(...; reversible, strong_upperbounds) (src/ss_fact.jl:97-113): the device's sticky timeline against the oracle, under
    both schedules of the emulation."""
    G = zzb.grid_precision(7, 9, shift=0.5)
    rng = np.random.default_rng(4)
    d = G.n
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c, kap = 4.0 * G.colnorms(), np.full(d, 0.8)
    mode = O.PARITY_MODE | (O.STICKY_REVERSIBLE if reversible else 0) | (O.STICKY_STRONG_UB if strong else 0)
    ref = O.spdmp(G, G, 0.0, x0, th0, 6.0, c, kappa=kap, mode=mode)
    plain = O.spdmp(G, G, 0.0, x0, th0, 6.0, c, kappa=kap)
    assert len(ref.events) > 200 and not np.array_equal(ref.events["t"][:len(plain.events)], plain.events["t"][:len(ref.events)])
    for tiles in (0, 5):
        got = O.window_sim(G, G, 0.0, x0, th0, 6.0, c, kappa=kap, reversible=reversible, strong_upperbounds=strong, async_tiles=tiles)
        O.assert_same_run(ref, got)


@pytest.mark.parametrize("stickyzz", [False, True])
def test_sticky_adaptation_of_the_bounds(zzb, stickyzz):
    """sspdmp / sspdmp2 (...; adapt = true, factor) (src/ss_fact.jl:132-136, src/stickyzz.jl:305-309): too small a c is multiplied by
    `factor` when a proposal is accepted with l > lb; device timeline against the oracle under both schedules."""
    G = zzb.grid_precision(7, 8, shift=0.5)
    rng = np.random.default_rng(12)
    d = G.n
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c, kap = 0.05 * G.colnorms(), np.full(d, 0.8)
    mode = O.PARITY_MODE | (O.STICKYZZ if stickyzz else 0)
    with pytest.raises(O.BoundError):
        O.spdmp(G, G.scaled(0.5), 0.0, x0, th0, 8.0, c, kappa=kap, mode=mode)
    ref = O.spdmp(G, G.scaled(0.5), 0.0, x0, th0, 8.0, c, kappa=kap, mode=mode, adapt=True, factor=1.5)
    assert (ref.c != c).sum() > 5 and len(ref.events) > 200
    for tiles in (0, 4):
        got = O.window_sim(G, G.scaled(0.5), 0.0, x0, th0, 8.0, c, kappa=kap, stickyzz=stickyzz, adapt=True, factor=1.5, async_tiles=tiles)
        O.assert_same_run(ref, got)


@pytest.mark.parametrize("strong", [False, True])
def test_dense_sticky_sampler_stickyzz(zzb, strong):
    """stickyzz / sspdmp2 (src/stickyzz.jl:176-338): the sspdmp loop with proposal times at rate 0.01 + (a + b t)^+ and coordinates
    that start at 0 starting frozen -- the device's sticky timeline with ZZ_STICKY_ZZ against the oracle (ZZO_STICKYZZ), under both
    schedules of the emulation; the floor and the frozen start both change the run."""
    G = zzb.grid_precision(8, 7, shift=0.5)
    rng = np.random.default_rng(9)
    d = G.n
    x0 = np.where(rng.random(d) < 0.6, rng.standard_normal(d), 0.0)
    th0 = rng.choice(np.array([-1.5, -1.0, 1.0, 0.5]), d)
    c, kap = 6.0 * G.colnorms(), rng.choice(np.array([0.3, 0.8, 2.0]), d)
    mode = O.PARITY_MODE | O.STICKYZZ | (O.STICKY_STRONG_UB if strong else 0)
    ref = O.spdmp(G, G, 0.0, x0, th0, 8.0, c, kappa=kap, mode=mode)
    assert len(ref.events) > 200
    started = {int(i) for i in ref.events["i"][ref.events["t"] < 1e-9]}
    first = {}
    for t, i, x, th in ref.events:
        first.setdefault(int(i), (t, x, th))
    frozen0 = [j + 1 for j in range(d) if x0[j] == 0.0 and (j + 1) in first]
    assert frozen0 and all(first[i][1] == 0.0 and first[i][2] == th0[i - 1] for i in frozen0)   # first event = thaw at 0 with the kept velocity
    xs = np.where(x0 == 0.0, 1e-300, x0)   # (no coordinate AT 0: the plain sspdmp start, for comparison)
    plain = O.spdmp(G, G, 0.0, xs, th0, 8.0, c, kappa=kap, mode=O.PARITY_MODE | (O.STICKY_STRONG_UB if strong else 0))
    assert not np.array_equal(ref.events["t"][:50], plain.events["t"][:50])
    for tiles in (0, 5):
        got = O.window_sim(G, G, 0.0, x0, th0, 8.0, c, kappa=kap, strong_upperbounds=strong, stickyzz=True, async_tiles=tiles)
        O.assert_same_run(ref, got)


@pytest.mark.parametrize("lattice,tiles", [(True, 0), (True, 5), (False, 0), (False, 4)])
def test_zigzag_refreshments(zzb, lattice, tiles):
    """ZigZag with velocity refreshments (Z.lambdaref > 0: src/sfact.jl:78-114,188-190): the device timeline
    (zz_timeline_refresh) inside the schedule emulation against the oracle's per-coordinate-clock contract, both schedules."""
    rng = np.random.default_rng(6)
    if lattice:
        G = zzb.grid_precision(9, 7, shift=0.3)
        Gb, mu, h = G, None, None
    else:
        G = zzb.random_sparse_spd(40, deg=2, seed=3)
        Gb, mu, h = G.scaled(0.9), 0.1 * rng.standard_normal(40), 0.2 * rng.standard_normal(40)
    d = G.n
    sigma = 0.5 + rng.random(d)
    x0, th0 = rng.standard_normal(d), sigma * rng.choice(np.array([-1.0, 1.0]), d)
    c = 3.0 * G.colnorms() * sigma.max()
    ref = O.spdmp(G, Gb, 0.0, x0, th0, 6.0, c, h=h, mu=mu, seed=(5, 7), refresh=(sigma, 0.8 * d), adapt=True)
    nrefresh = len(ref.events) - int(ref.acc.sum())
    assert nrefresh > 0.5 * 0.8 * d * 6.0 * 0.5 and int(ref.acc.sum()) > 50          # both kinds of event occur
    got = O.window_sim(G, Gb, 0.0, x0, th0, 6.0, c, h=h, mu=mu, seed=(5, 7), refresh=(sigma, 0.8 * d), adapt=True, async_tiles=tiles)
    O.assert_same_run(ref, got)


@pytest.mark.parametrize("tiles", [0, 4])
def test_wide_columns_of_the_velocity_list_samplers(zzb, tiles):
    """Columns of 9 .. 32 entries for the samplers without a list-walking fallback (sticky, Boomerang, refreshments): the wider
    out-of-line instantiation of gather + timeline (zz_process_node_wide) against the oracle, both schedules."""
    from test_boomerang import boom_inputs
    G = zzb.random_sparse_spd(70, deg=5, seed=2)
    d = G.n
    assert 8 < np.diff(G.colptr).max() <= 32
    rng = np.random.default_rng(31)
    x0, th0 = rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d)
    c, kap = 4.0 * G.colnorms(), rng.choice(np.array([0.3, 0.8, 2.0]), d)
    ref = O.spdmp(G, G, 0.0, x0, th0, 5.0, c, kappa=kap)
    got = O.window_sim(G, G, 0.0, x0, th0, 5.0, c, kappa=kap, async_tiles=tiles)
    O.assert_same_run(ref, got)
    assert len(ref.events) > 150
    sigma = 0.5 + rng.random(d)
    ths = sigma * rng.choice(np.array([-1.0, 1.0]), d)
    cr = 3.0 * G.colnorms() * sigma.max()
    ref = O.spdmp(G, G, 0.0, x0, ths, 4.0, cr, seed=(5, 7), refresh=(sigma, 0.8 * d), adapt=True)
    got = O.window_sim(G, G, 0.0, x0, ths, 4.0, cr, seed=(5, 7), refresh=(sigma, 0.8 * d), adapt=True, async_tiles=tiles)
    O.assert_same_run(ref, got)
    Zg, bsig, xb, thb, cb = boom_inputs(zzb, G, 1.0, rng)
    ref = O.spdmp(G, Zg, 0.0, xb, thb, 3.0, cb, seed=(5, 6), boom=(bsig, 20.0, 0.2))
    got = O.window_sim(G, Zg, 0.0, xb, thb, 3.0, cb, seed=(5, 6), boom=(bsig, 20.0, 0.2), async_tiles=tiles)
    O.assert_same_run(ref, got)
    assert len(ref.events) > 50


def test_sticky_family_randomised_campaign(zzb):
    """Random sticky problems through the device timeline inside the schedule emulation against the oracle: sspdmp and sspdmp2
    (stickyzz: rate floor, frozen starts), with and without reversible / strong_upperbounds / adapt, per-coordinate thaw rates,
    non-unit speeds, linear term, lattices and general sparse graphs with columns up to 32 entries, both schedules."""
    rng = np.random.default_rng(77)
    nbound = 0
    for case in range(36):
        if rng.random() < 0.5:
            G = zzb.grid_precision(int(rng.integers(2, 12)), int(rng.integers(2, 12)), shift=float(rng.choice([0.1, 0.5])))
        else:
            G = zzb.random_sparse_spd(int(rng.integers(4, 90)), deg=int(rng.integers(1, 6)), seed=int(rng.integers(1 << 30)))
            if np.diff(G.colptr).max() > 32:
                continue
        d = G.n
        stickyzz = bool(rng.random() < 0.5)
        x0 = np.where(rng.random(d) < (0.7 if stickyzz else 1.0), rng.standard_normal(d), 0.0 if stickyzz else 1e-3)
        th0 = rng.choice(np.array([-1.5, -1.0, -0.5, 0.5, 1.0, 1.5]), d)
        h = 0.3 * rng.standard_normal(d) if rng.random() < 0.3 else None
        adapt = bool(rng.random() < 0.5)
        cs = float(rng.choice([0.05, 0.5, 3.0])) if adapt else float(rng.choice([0.5, 3.0, 6.0]))
        c = cs * G.colnorms() + (np.abs(h) if h is not None else 0.0)
        kap = rng.choice(np.array([0.2, 0.8, 3.0]), d)
        rev, sub = bool(rng.random() < 0.3), bool(rng.random() < 0.3)
        mode = O.PARITY_MODE | (O.STICKYZZ if stickyzz else 0) | (O.STICKY_REVERSIBLE if rev else 0) | (O.STICKY_STRONG_UB if sub else 0)
        T = float(rng.choice([1.0, 4.0, 9.0]))
        sd = (int(rng.integers(1 << 40)), int(rng.integers(1 << 40)))
        Gb = G.scaled(float(rng.choice([0.7, 1.0, 1.2]))) if rng.random() < 0.5 else G
        kw = dict(delta0=float(10 ** rng.uniform(-3, 0.3)), target_frac=float(10 ** rng.uniform(-1.3, 0.7)))
        if case % 2:
            kw.update(async_tiles=int(rng.integers(1, 7)), order_seed=int(rng.integers(1, 1 << 30)))
        try:
            ref = O.spdmp(G, Gb, 0.0, x0, th0, T, c, h=h, kappa=kap, mode=mode, adapt=adapt, factor=1.5, seed=sd)
        except O.BoundError:
            ref = None
        try:
            got = O.window_sim(G, Gb, 0.0, x0, th0, T, c, h=h, kappa=kap, stickyzz=stickyzz, reversible=rev, strong_upperbounds=sub,
                               adapt=adapt, factor=1.5, seed=sd, **kw)
        except O.BoundError:
            got = None
        assert (ref is None) == (got is None), case
        if ref is None:
            nbound += 1
        else:
            O.assert_same_run(ref, got)
    assert nbound < 30
