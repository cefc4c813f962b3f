"""Randomised parity campaign: the CUDA path against the CPU oracle, bit for bit, over random graphs (lattices of random
shape, random sparse graphs), samplers (ZigZag, LocalBound, sticky, Boomerang), targets (mu, h, scaled sampler matrix) and
tuning knobs (window target, first window, tag budget, grid size).  Used by tests/test_gpu_fuzz.py (a short run) and by
tools/fuzz_parity.py (long runs)."""
import numpy as np

import oracle_lib as O


def run_cases(z, ncases, seed, verbose=False, sim=False, async_sim=False, seq=False):
    """Returns (failures, bound_errors); prints failing cases.  sim = True: instead of the CUDA path, the host emulation of the
    device schedule (oracle/zz_window_sim.cpp: the kernels' own per-coordinate code) is compared with the oracle -- no GPU.
    seq = True: the sequential-chain schedule of the device (zz_seq.cuh: plain ZigZag, components of at most 2700 coordinates,
    1 / 2 / 4 warps per chain, block-diagonal problems with scattered components among the cases)."""
    rng = np.random.default_rng(seed)
    nb = [0]
    class R:
        pass


    def rand_tune():
        t = {}
        if rng.random() < 0.5: t["target_frac"] = float(rng.choice([0.02, 0.1, 0.25, 0.6, 2.0]))
        if rng.random() < 0.3: t["delta0"] = float(rng.choice([1e-4, 1e-2, 0.5]))
        if rng.random() < 0.3: t["tag_limit"] = int(rng.choice([40, 200, 5000]))
        if rng.random() < 0.3: t["grid"] = int(rng.choice([1, 3, 37, 148]))
        if rng.random() < 0.3: t["target_flip_frac"] = float(rng.choice([0.01, 0.045, 0.2]))
        return t or None


    bad = 0
    for case in range(ncases):
        kind = rng.choice(["zigzag", "zigzag", "localbound", "sticky", "boomerang"])
        if seq:
            kind = "zigzag"
        if rng.random() < 0.5:
            big = rng.random() < 0.1 and not seq
            m, n = (int(rng.integers(60, 200)), int(rng.integers(60, 200))) if big else (int(rng.integers(2, 40)), int(rng.integers(2, 40)))
            G = z.grid_precision(m, n, shift=float(rng.choice([0.01, 0.5])))
            Zg = G
        else:
            d = int(rng.integers(3, 120))
            G = next(gg for gg in (z.random_sparse_spd(d, deg=int(rng.integers(1, 3)), seed=int(rng.integers(1 << 30))) for _ in range(50))
                     if np.diff(gg.colptr).max() <= 8 or kind in ("zigzag", "localbound"))
            Zg = G.scaled(float(rng.choice([0.8, 1.0, 1.3]))) if kind in ("zigzag", "sticky", "boomerang") and rng.random() < 0.5 else G
        d = G.n
        x0 = rng.standard_normal(d)
        th0 = rng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
        T = float(rng.choice([0.5, 2.0, 6.0]))
        sd = (int(rng.integers(1 << 40)), int(rng.integers(1 << 40)))
        tune = rand_tune()
        if seq:
            tune = dict(schedule=2, seq_warps=int(rng.choice([0, 1, 2, 4, 8])))
            if rng.random() < 0.4 and d >= 6:   # several chains: drop the couplings between K chunks, then scatter the coordinates
                from zzb200.problems import CSC
                K = int(rng.integers(2, 5))
                perm = rng.permutation(d)
                dense = O.block_diagonal(G, K).to_scipy().toarray() if d % K == 0 else None
                if dense is not None:
                    G = CSC.from_dense(dense[np.ix_(perm, perm)])
                    Zg = G
        adapt = bool(rng.random() < 0.6)
        cs = float(rng.choice([1e-6, 0.3, 1.0, 3.0])) if adapt else float(rng.choice([0.5, 2.0, 4.0]))   # 0.5: provokes bound errors
        c = cs * G.colnorms()
        mu = 0.2 * rng.standard_normal(d) if (rng.random() < 0.3 and kind in ("zigzag", "boomerang")) else None
        h = 0.3 * rng.standard_normal(d) if rng.random() < 0.3 else None
        if kind == "localbound":
            Zg, mu = G, None
            c = np.full(d, float(rng.choice([0.05, 0.5, 2.0])))
        desc = f"case {case}: {kind} d={d} T={T} adapt={adapt} cs={cs} mu={mu is not None} h={h is not None} tune={tune}"
        tgt = z.GaussianPotential(G, h)

        def sim_fn(**kw):
            tk = {k: v for k, v in (tune or {}).items() if k in ("target_frac", "delta0", "tag_limit")}
            if async_sim:   # the asynchronous tile-local schedule: the `grid` knob is the number of tiles; a fresh interleaving per case
                tk["async_tiles"] = int((tune or {}).get("grid", 1 + case % 9))
                tk["order_seed"] = 1000 + case
            r = O.window_sim(G, Zg, 0.0, x0, th0, T, c, h=h, mu=mu, seed=sd, **tk, **kw)
            return r.events, r.t, r.x, r.theta, r.c, r.acc, r.num
        if kind == "sticky":
            kappa = np.full(d, float(rng.choice([0.3, 1.0, 5.0])))
            c = 3.0 * G.colnorms() + (np.abs(h) if h is not None else 0.0)
            ref_fn = lambda: O.spdmp(G, Zg, 0.0, x0, th0, T, c, h=h, seed=sd, kappa=kappa)

            def dev_fn():
                if sim:
                    return sim_fn(kappa=kappa)
                Xi, (t, x, th), (acc, num), cc = z.sspdmp(tgt, 0.0, x0, th0, T, c, z.ZigZag(Zg, np.zeros(d)), kappa, seed=sd, tune=tune)
                return Xi.events, t, x, th, cc, Xi.acc_per_coordinate, num
        elif kind == "boomerang":
            diag = Zg.to_scipy().diagonal()
            boom = (diag ** -0.5, float(rng.choice([0.5, 5.0, 50.0])), float(rng.choice([0.0, 0.3])))
            th0 = rng.standard_normal(d) / np.sqrt(diag)
            ref_fn = lambda: O.spdmp(G, Zg, 0.0, x0, th0, T, c, h=h, mu=mu, seed=sd, adapt=adapt, boom=boom)

            def dev_fn():
                if sim:
                    return sim_fn(boom=boom, adapt=adapt)
                F = z.FactBoomerang(Zg, np.zeros(d) if mu is None else mu, boom[1], boom[0], rho=boom[2])
                Xi, (t, x, th), (acc, num), cc = z.spdmp(tgt, 0.0, x0, th0, T, c, F, seed=sd, adapt=adapt, tune=tune)
                return Xi.events, t, x, th, cc, acc, num
        else:
            mode = O.PARITY_MODE | (O.LOCAL_BOUND if kind == "localbound" else 0)
            ref_fn = lambda: O.spdmp(G, Zg, 0.0, x0, th0, T, c, h=h, mu=mu, seed=sd, adapt=adapt, mode=mode)

            def dev_fn():
                if sim:
                    return sim_fn(adapt=adapt, local_bound=(kind == "localbound"))
                cc_in = z.LocalBound(c) if kind == "localbound" else c
                Xi, (t, x, th), (acc, num), cc = z.spdmp(tgt, 0.0, x0, th0, T, cc_in, z.ZigZag(Zg, np.zeros(d) if mu is None else mu),
                                                         seed=sd, adapt=adapt, tune=tune)
                return Xi.events, t, x, th, (cc.c if kind == "localbound" else cc), acc, num
        # error("Tuning parameter `c` too small.") must be raised by both sides or by neither
        try:
            ref = ref_fn()
        except O.BoundError:
            ref = None
        try:
            out = dev_fn()
        except (z.BoundError, O.BoundError):
            out = None
        try:
            assert (ref is None) == (out is None), "bound violation reported by one side only"
            if ref is None:
                nb[0] += 1
            else:
                got = R()
                got.events, got.t, got.x, got.theta, got.c, got.acc, got.num = out
                O.assert_same_run(ref, got)
            if verbose and case % 50 == 0:
                print("ok  ", desc, "bound error on both sides" if ref is None else f"events={len(ref.events)} num={ref.num}", flush=True)
        except AssertionError as e:
            bad += 1
            print("FAIL", desc, repr(e)[:200], flush=True)

    return bad, nb[0]
