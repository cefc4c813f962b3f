"""SURVEY 8(a) row a18 / BASELINE config 4: the reference's sparse sticky ZigZag (src/sparsestickyzz.jl, `sspdmp3`) restated
for the CPU, and the statistical check behind DESIGN.md's claim that the sticky sampler the device runs (src/ss_fact.jl,
per-coordinate thaw clocks and affine bounds) is the same process in law (one thaw clock of rate kappa * #frozen with a
uniform pick, constant expiring bounds)."""
import numpy as np
import pytest

import oracle_lib as O
from sticky_stats import chain_precision, occupancy


def test_event_records_follow_the_reference(zzb):
    """Thaw: (t', i, 0, +-1); hit: the record of a deleted coordinate (t', i, 0, 0) (sparsestickyzz.jl:20-26,326,363);
    reflection: velocity flipped (:392).  A coordinate alternates thaw -> reflections -> hit."""
    p = 12
    G = chain_precision(zzb, p)
    r = O.sparsestickyzz(G, np.zeros(p), np.ones(p), 300.0, 2.5, 0.3, seed=(1, 2))
    ev = r.events
    assert len(ev) > 500 and np.all(np.diff(ev["t"]) >= 0) and ev["t"][-1] >= 300.0 and ev["t"][-2] < 300.0
    for j in range(1, p + 1):
        e = ev[ev["i"] == j]
        frozen = True
        for t, _, x, th in e:
            if frozen:
                assert x == 0.0 and abs(th) == 1.0      # thaw
                frozen = False
            elif th == 0.0:
                assert x == 0.0                          # hit: frozen again
                frozen = True
            else:
                assert abs(th) == 1.0                    # reflection
    assert r.num >= r.acc.sum() > 0


def test_sticky_rule_remembers_the_direction(zzb):
    """rule :sticky re-enters with the sign the coordinate had when it hit 0 (-1 + 2p[i], :316,386-388): the velocity of a
    thaw equals the velocity before the preceding hit."""
    p = 8
    G = chain_precision(zzb, p)
    ev = O.sparsestickyzz(G, np.zeros(p), np.ones(p), 400.0, 2.5, 0.5, rule="sticky", seed=(3, 4)).events
    checked = 0
    for j in range(1, p + 1):
        th = ev[ev["i"] == j]["theta"]
        for k in range(2, len(th)):
            if th[k - 1] == 0.0 and th[k] != 0.0:       # ... v, hit, thaw
                assert th[k] == th[k - 2]
                checked += 1
    assert checked > 20


@pytest.mark.parametrize("p,kappa,T", [(30, 0.5, 12000.0), (12, 0.1, 30000.0)])
def test_same_law_as_the_sticky_sampler_of_the_device_contract(zzb, p, kappa, T):
    """Occupancy (fraction of time away from 0) and second moments per coordinate agree between the reference's sparse
    sticky algorithm and sspdmp in the device's parity mode (ctr|lazy) -- two independent runs, Monte Carlo tolerance."""
    G = chain_precision(zzb, p)
    x0 = np.zeros(p)
    a = O.sparsestickyzz(G, x0, np.ones(p), T, 2.5, kappa, rule="sticky", seed=(3, 4))
    th0 = np.random.default_rng(0).choice(np.array([-1.0, 1.0]), p)
    b = O.spdmp(G, G, 0.0, x0, th0, T, G.colnorms(), kappa=np.full(p, kappa), seed=(5, 6), mode=O.PARITY_MODE)
    oa, ma = occupancy(a.events, p, x0, T)
    ob, mb = occupancy(b.events, p, x0, T)
    assert abs(oa.mean() - ob.mean()) < 0.01 and np.abs(oa - ob).max() < 0.04
    assert abs(ma.mean() - mb.mean()) < 0.03 * mb.mean() and np.abs(ma - mb).max() < 0.15 * mb.max()
    assert abs(len(a.events) - len(b.events)) < 0.02 * len(b.events)   # thaws + hits + reflections per unit time


def test_reversible_rule_and_adaptation(zzb):
    p = 16
    G = chain_precision(zzb, p)
    a = O.sparsestickyzz(G, np.zeros(p), np.ones(p), 4000.0, 2.5, 0.4, rule="reversible", seed=(7, 8))
    b = O.sparsestickyzz(G, np.zeros(p), np.ones(p), 4000.0, 2.5, 0.4, rule="sticky", seed=(9, 10))
    oa, _ = occupancy(a.events, p, np.zeros(p), 4000.0)
    ob, _ = occupancy(b.events, p, np.zeros(p), 4000.0)
    assert 0.2 < oa.mean() < 0.9 and 0.2 < ob.mean() < 0.9      # both mix between frozen and moving
    # c far too small: the strong bound is violated -> error without adapt (:334,381), c multiplied with adapt (:336,383)
    with pytest.raises(O.BoundError):
        O.sparsestickyzz(G, np.zeros(p), np.ones(p), 500.0, 0.05, 0.4, seed=(1, 2))
    r = O.sparsestickyzz(G, np.zeros(p), np.ones(p), 500.0, 0.05, 0.4, adapt=True, multiplier=1.5, seed=(1, 2))
    assert r.c[0] > 0.05 and np.all(r.c == r.c[0])


def test_nonzero_start_and_linear_term(zzb):
    """Active coordinates at t = 0 (x0 != 0, velocities given) and a target with a linear term (the heart target's data term,
    research/sticky/heart/heart_sparse2.jl:52, is of this form)."""
    p = 10
    G = chain_precision(zzb, p)
    rng = np.random.default_rng(2)
    x0 = np.where(rng.random(p) < 0.5, rng.standard_normal(p), 0.0)
    r = O.sparsestickyzz(G, x0, rng.choice(np.array([-1.0, 1.0]), p), 3000.0, 4.0, 0.3, h=np.full(p, 0.8), seed=(1, 2))
    occ, _ = occupancy(r.events, p, x0, 3000.0)
    r0 = O.sparsestickyzz(G, x0, np.ones(p), 3000.0, 4.0, 0.3, seed=(1, 2))
    occ0, _ = occupancy(r0.events, p, x0, 3000.0)
    assert occ.mean() > occ0.mean() + 0.05     # the pull of h keeps coordinates away from 0 longer


@pytest.mark.parametrize("rule", ["sticky", "reversible"])
def test_parity_arithmetic_has_the_same_law(zzb, rule):
    """zzo_sparsestickyzz_ctr (per-coordinate streams and thaw clocks, flip-anchored positions -- the contract of a future
    strong-bound device kernel) against the faithful restatement: occupancy, second moments, event and proposal rates."""
    p, kappa, T = 30, 0.5, 12000.0
    G = chain_precision(zzb, p)
    x0 = np.zeros(p)
    a = O.sparsestickyzz(G, x0, np.ones(p), T, 2.5, kappa, rule=rule, seed=(3, 4))
    b = O.sparsestickyzz(G, x0, np.ones(p), T, 2.5, kappa, rule=rule, seed=(3, 4), ctr=True)
    oa, ma = occupancy(a.events, p, x0, T)
    ob, mb = occupancy(b.events, p, x0, T)
    assert abs(oa.mean() - ob.mean()) < 0.01 and np.abs(oa - ob).max() < 0.04
    assert abs(ma.mean() - mb.mean()) < 0.04 * mb.mean()
    assert abs(len(a.events) - len(b.events)) < 0.02 * len(b.events) and abs(a.num - b.num) < 0.02 * b.num
    # deterministic in the seed, different across seeds; a non-zero start is honoured
    b2 = O.sparsestickyzz(G, x0, np.ones(p), 200.0, 2.5, kappa, rule=rule, seed=(3, 4), ctr=True)
    b3 = O.sparsestickyzz(G, x0, np.ones(p), 200.0, 2.5, kappa, rule=rule, seed=(3, 5), ctr=True)
    assert np.array_equal(b2.events, b.events[: len(b2.events)]) and not np.array_equal(b3.events["t"][:20], b2.events["t"][:20])
    x1 = np.where(np.arange(p) % 3 == 0, 0.7, 0.0)
    b4 = O.sparsestickyzz(G, x1, -np.ones(p), 50.0, 2.5, kappa, rule=rule, seed=(1, 2), ctr=True)
    first = {int(i): (x, th) for t, i, x, th in b4.events[::-1]}
    assert all(first[j + 1][0] != 0.0 or first[j + 1][1] == 0.0 for j in range(p) if x1[j] != 0 and (j + 1) in first)


def test_windowed_schedule_with_the_strong_bound_timeline_equals_its_contract(zzb):
    """zz_strong.h (the per-coordinate timeline a strong-bound device kernel would run: own items only, neighbours enter
    through positions) inside the host emulation of the device schedule against zzo_sparsestickyzz_ctr, bit for bit: random
    chains and lattices, both rules, frozen and moving starts, linear term, window policies, tag rebases.  The relaxation needs
    ~3 passes per window here against ~12 for the samplers in which a flip reschedules its neighbours."""
    rng = np.random.default_rng(0)
    max_iters = 0
    for case in range(40):
        if rng.random() < 0.5:
            p = int(rng.integers(2, 120))
            G = chain_precision(zzb, max(p, 2))
        else:
            G = zzb.grid_precision(int(rng.integers(2, 9)), int(rng.integers(2, 9)), shift=0.1)
        p = G.n
        x0 = np.where(rng.random(p) < rng.choice([0.0, 0.3, 1.0]), rng.standard_normal(p), 0.0)
        th0 = rng.choice(np.array([-1.0, 1.0]), p)
        h = 0.5 * rng.standard_normal(p) if rng.random() < 0.3 else None
        rule = str(rng.choice(["sticky", "reversible"]))
        kappa, T = float(rng.choice([0.1, 0.5, 3.0])), float(rng.uniform(5, 60))
        c = float(rng.choice([3.0, 6.0])) + 2.0 * float(np.abs(G.nzval).max()) + (0.0 if h is None else float(np.abs(h).max()))
        sd = (int(rng.integers(1 << 40)), int(rng.integers(1 << 40)))
        kw = dict(delta0=float(10 ** rng.uniform(-3, 0.3)), target_frac=float(10 ** rng.uniform(-1.3, 0.7)))
        if rng.random() < 0.3:
            kw["tag_limit"] = int(rng.integers(30, 200))
        try:
            ref = O.sparsestickyzz(G, x0, th0, T, c, kappa, h=h, rule=rule, seed=sd, ctr=True)
        except O.BoundError:
            ref = None
        try:
            sim = O.window_sim(None, G, 0.0, x0, th0, T, np.full(p, c), h=h, kappa=kappa, strong=rule, seed=sd, **kw)
        except O.BoundError:
            sim = None
        assert (ref is None) == (sim is None), case
        if ref is not None:
            O.assert_same_run(ref, sim)
            max_iters = max(max_iters, sim.stats["max_iters"])
    assert 2 <= max_iters <= 8


def test_asynchzz_process_contract_and_schedule_emulation(zzb):
    """asynchzz / sspdmp4 (src/asynchzz.jl): the strong-bound sticky process with c[i], kappa[i] per coordinate, a start time and
    the velocity kept over a freeze.  (1) With equal constants, unit speeds and t0 = 0 it is sspdmp3's :sticky process, bit for
    bit; (2) the device timeline inside the host emulation of both schedules equals the contract on random problems (per-coordinate
    constants, non-unit speeds, shifted start); (3) a coordinate continues with exactly the velocity it froze with."""
    rng = np.random.default_rng(5)
    p = 40
    G = chain_precision(zzb, p)
    x0 = np.where(rng.random(p) < 0.5, rng.standard_normal(p), 0.0)
    th0 = rng.choice(np.array([-1.0, 1.0]), p)
    a = O.sparsestickyzz(G, x0, th0, 40.0, 6.0, 0.5, rule="sticky", seed=(3, 4), ctr=True)
    th0a = np.where(x0 != 0.0, th0, 1.0)    # sspdmp3 re-enters a coordinate that starts frozen with +1
    b = O.strongsticky(G, 0.0, x0, th0a, 40.0, np.full(p, 6.0), np.full(p, 0.5), rule="keep", seed=(3, 4))
    O.assert_same_run(a, b)
    for case in range(24):
        if rng.random() < 0.5:
            G = chain_precision(zzb, int(rng.integers(2, 90)))
        else:
            G = zzb.grid_precision(int(rng.integers(2, 8)), int(rng.integers(2, 8)), shift=0.1)
        p = G.n
        x0 = np.where(rng.random(p) < rng.choice([0.0, 0.4, 1.0]), rng.standard_normal(p), 0.0)
        th0 = rng.choice(np.array([-1.5, -1.0, -0.5, 0.5, 1.0, 2.0]), p)
        h = 0.5 * rng.standard_normal(p) if rng.random() < 0.3 else None
        t0 = float(rng.choice([0.0, -3.0, 7.5]))
        T = t0 + float(rng.uniform(5, 40))
        base = 3.0 + 4.0 * float(np.abs(G.nzval).max()) + (0.0 if h is None else float(np.abs(h).max()))
        c = base * rng.uniform(1.0, 2.0, p)
        kappa = rng.choice(np.array([0.1, 0.5, 3.0]), p)
        sd = (int(rng.integers(1 << 40)), int(rng.integers(1 << 40)))
        kw = dict(delta0=float(10 ** rng.uniform(-3, 0.3)), target_frac=float(10 ** rng.uniform(-1.3, 0.7)))
        if case % 2:
            kw.update(async_tiles=int(rng.integers(1, 6)), order_seed=int(rng.integers(1, 1 << 30)))
        try:
            ref = O.strongsticky(G, t0, x0, th0, T, c, kappa, h=h, seed=sd)
        except O.BoundError:
            ref = None
        try:
            sim = O.window_sim(None, G, t0, x0, th0, T, c, h=h, kappa=kappa, strong="keep", seed=sd, **kw)
        except O.BoundError:
            sim = None
        assert (ref is None) == (sim is None), case
        if ref is None:
            continue
        O.assert_same_run(ref, sim)
        # velocity before a freeze == velocity after the following thaw, per coordinate
        vel = {j + 1: (th0[j] if x0[j] != 0.0 else 0.0) for j in range(p)}
        saved = {j + 1: th0[j] for j in range(p)}
        for t, i, x, th in ref.events:
            i = int(i)
            if th == 0.0:
                saved[i] = vel[i]
            elif vel[i] == 0.0:
                assert th == saved[i], (case, i)
            vel[i] = th
