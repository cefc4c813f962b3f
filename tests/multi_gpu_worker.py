"""Worker of tests/test_gpu_multi.py: run under torchrun with N ranks, one GPU each; compares the sharded GPU run with
the CPU oracle on rank 0 and prints MULTI_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    zzb = graft.load_package()
    zzb.init(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = []
    G, x0, th0, c = zzb.gmrf_config(24)
    cases.append(("grid24", G, G, x0, th0, c, None, None, 4.0, False))
    G2, x2, t2, c2 = zzb.gmrf_config(64)
    cases.append(("grid64", G2, G2, x2, t2, c2, None, None, 2.0, False))
    d = 60
    Gt = zzb.random_sparse_spd(d, deg=3, seed=1)
    rng = np.random.default_rng(1)
    cases.append(("sparse60", Gt, Gt.scaled(0.9), rng.standard_normal(d), rng.choice(np.array([-1.0, 1.0]), d),
                  0.2 * Gt.colnorms(), 0.3 * rng.standard_normal(d), 0.1 * rng.standard_normal(d), 15.0, True))
    for name, Gt_, Gb_, x0_, th0_, c_, h_, mu_, T, adapt in cases:
        Z = zzb.ZigZag(Gb_, np.zeros(Gt_.n) if mu_ is None else mu_)
        res, stats, ms = zzb.spdmp_sharded(zzb, zzb.GaussianPotential(Gt_, h_), Z, 0.0, x0_, th0_, T, c_, seed=(1, 2), adapt=adapt)
        if rank == 0:
            ref = O.spdmp(Gt_, Gb_, 0.0, x0_, th0_, T, c_, h=h_, mu=mu_, seed=(1, 2), adapt=adapt)

            class R:
                pass
            got = R()
            got.events, got.num, got.acc = res["events"], res["num"], res["acc"]
            got.t, got.x, got.theta, got.c, got.s1, got.s2 = res["t"], res["x"], res["theta"], res["c"], res["s1"], res["s2"]
            O.assert_same_run(ref, got)
            print(f"case {name}: world {world}, {len(ref.events)} events bit-exact, stats {stats['windows']} windows "
                  f"{stats['passes']} passes, {ms:.2f} ms", flush=True)
        dist.barrier()
    # randomised cases (the same generator state on every rank): random lattices / sparse graphs, bounds, tuning knobs
    frng = np.random.default_rng(int(os.environ.get("ZZB_MULTI_FUZZ_SEED", "5")))
    for case in range(int(os.environ.get("ZZB_MULTI_FUZZ", "20"))):
        if frng.random() < 0.6:
            m, n = int(frng.integers(2, 60)), int(frng.integers(world, 60))
            Gt_ = zzb.grid_precision(m, n, shift=float(frng.choice([0.01, 0.5])))
            Gb_ = Gt_
        else:
            d = int(frng.integers(2 * world, 150))
            Gt_ = zzb.random_sparse_spd(d, deg=int(frng.integers(1, 4)), seed=int(frng.integers(1 << 30)))
            Gb_ = Gt_.scaled(float(frng.choice([0.8, 1.0, 1.3]))) if frng.random() < 0.5 else Gt_
        d = Gt_.n
        x0_, th0_ = frng.standard_normal(d), frng.choice(np.array([-1.0, -0.5, 0.5, 1.0]), d)
        T = float(frng.choice([0.5, 2.0, 5.0]))
        cs = float(frng.choice([1e-6, 0.3, 1.0, 3.0]))
        c_ = cs * Gt_.colnorms()
        h_ = 0.3 * frng.standard_normal(d) if frng.random() < 0.3 else None
        mu_ = 0.2 * frng.standard_normal(d) if frng.random() < 0.3 else None
        tune = {}
        if frng.random() < 0.5: tune["target_frac"] = float(frng.choice([0.02, 0.1, 0.25, 0.6, 2.0]))
        if frng.random() < 0.3: tune["delta0"] = float(frng.choice([1e-4, 1e-2, 0.5]))
        if frng.random() < 0.3: tune["tag_limit"] = int(frng.choice([40, 200, 5000]))
        sd = (int(frng.integers(1 << 40)), int(frng.integers(1 << 40)))
        Z = zzb.ZigZag(Gb_, np.zeros(d) if mu_ is None else mu_)
        res, stats, ms = zzb.spdmp_sharded(zzb, zzb.GaussianPotential(Gt_, h_), Z, 0.0, x0_, th0_, T, c_, seed=sd, adapt=True, tune=tune or None)
        if rank == 0:
            ref = O.spdmp(Gt_, Gb_, 0.0, x0_, th0_, T, c_, h=h_, mu=mu_, seed=sd, adapt=True)

            class R2:
                pass
            got = R2()
            got.events, got.num, got.acc = res["events"], res["num"], res["acc"]
            got.t, got.x, got.theta, got.c, got.s1, got.s2 = res["t"], res["x"], res["theta"], res["c"], res["s1"], res["s2"]
            O.assert_same_run(ref, got)
            if case % 5 == 0:
                print(f"random case {case}: d={d} T={T} cs={cs} tune={tune}: {len(ref.events)} events bit-exact", flush=True)
        dist.barrier()
    # sticky ZigZag (src/ss_fact.jl) and factorised Boomerang sharded: the lists of these samplers carry the velocity after each
    # event, pushed into the neighbour ranks' replicas together with the event times
    def compare(name, ref, res, stats, ms, check_sums=True):
        class R3:
            pass
        got = R3()
        got.events, got.num, got.acc = res["events"], res["num"], res["acc"]
        got.t, got.x, got.theta, got.c = res["t"], res["x"], res["theta"], res["c"]
        if check_sums:
            got.s1, got.s2 = res["s1"], res["s2"]
        else:
            for f in ("s1", "s2"):
                if hasattr(ref, f):
                    delattr(ref, f)
        O.assert_same_run(ref, got)
        print(f"case {name}: world {world}, {len(ref.events)} events bit-exact, {stats['windows']} windows, {ms:.2f} ms", flush=True)

    srng = np.random.default_rng(17)
    for name, Gs in (("sticky grid 20x14", zzb.grid_precision(20, 14, shift=0.5)), ("sticky sparse 80", zzb.random_sparse_spd(80, deg=2, seed=10))):
        d = Gs.n
        x0_, th0_ = srng.standard_normal(d), srng.choice(np.array([-1.0, 1.0]), d)
        c_, kap = 4.0 * Gs.colnorms(), srng.choice(np.array([0.3, 0.8, 2.0]), d)
        for opts in (dict(), dict(reversible=True, strong_upperbounds=True)):
            res, stats, ms = zzb.spdmp_sharded(zzb, zzb.GaussianPotential(Gs), zzb.ZigZag(Gs, np.zeros(d)), 0.0, x0_, th0_, 6.0, c_, seed=(3, 4),
                                               run_kwargs=dict(kappa=kap, **opts))
            if rank == 0:
                mode = O.PARITY_MODE | (O.STICKY_REVERSIBLE if opts else 0) | (O.STICKY_STRONG_UB if opts else 0)
                compare(name + (" reversible strong_ub" if opts else ""), O.spdmp(Gs, Gs, 0.0, x0_, th0_, 6.0, c_, kappa=kap, seed=(3, 4), mode=mode),
                        res, stats, ms, check_sums=False)
            dist.barrier()
    from test_boomerang import boom_inputs
    for name, Gb2 in (("boomerang grid 18x16", zzb.grid_precision(18, 16)), ("boomerang sparse 80", zzb.random_sparse_spd(80, deg=2, seed=5))):
        Zg, sigma, x0_, th0_, c_ = boom_inputs(zzb, Gb2, 1.0, srng)
        F = zzb.FactBoomerang(Zg, np.zeros(Gb2.n), 20.0, sigma, rho=0.2)
        res, stats, ms = zzb.spdmp_sharded(zzb, zzb.GaussianPotential(Gb2), F, 0.0, x0_, th0_, 3.0, c_, seed=(5, 6), run_kwargs=dict(boomerang=F))
        if rank == 0:
            compare(name, O.spdmp(Gb2, Zg, 0.0, x0_, th0_, 3.0, c_, seed=(5, 6), mode=O.PARITY_MODE, boom=(sigma, 20.0, 0.2)), res, stats, ms,
                    check_sums=False)
        dist.barrier()
    if rank == 0:
        print("MULTI_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
