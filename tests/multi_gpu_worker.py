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
    if rank == 0:
        print("MULTI_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
