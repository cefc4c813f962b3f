"""Randomised parity campaign for the logistic target (config 3): random designs, subsample counts, bounds (including ones
that are violated), window policies.  Default: the host emulation of the device schedule (the kernel's own per-coordinate
code, oracle/zz_window_sim.cpp) against the sequential oracle on the CPU; with --gpu the CUDA path through the C-ABI.
    python tools/fuzz_logistic.py [--gpu] SEED N"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import logistic_cases as LC  # noqa: E402
import oracle_lib as O  # noqa: E402


def run(z, seed, n, gpu=False, verbose=True, schedule=None):
    rng = np.random.default_rng(seed)
    n_ok = n_err = 0
    t0 = time.time()
    for it in range(n):
        K = rng.integers(1, 4)
        levels = tuple(int(v) for v in rng.integers(2, 6, size=K))
        r, m, dseed, k = int(rng.integers(0, 3)), int(rng.integers(6, 14)), int(rng.integers(0, 10 ** 6)), int(rng.integers(1, 13))
        try:
            cfg = LC.make(z, levels, r, m, dseed, k=k)
        except AssertionError:     # a design column without entries: refused by both sides
            continue
        T = float(rng.uniform(2, 25))
        sd = (int(rng.integers(1, 2 ** 40)), int(rng.integers(1, 2 ** 40)))
        adapt, factor = bool(rng.random() < 0.8), float(rng.choice([1.5, 2.0, 5.0]))
        c = cfg["c"] * float(rng.choice([1.0, 10.0, 100.0, 1000.0]))
        kw = dict(delta0=float(10 ** rng.uniform(-3, 0.5)), target_frac=float(10 ** rng.uniform(-1.3, 0.7)))
        if rng.random() < 0.3:
            kw["tag_limit"] = int(rng.integers(30, 200))
        if gpu and schedule is not None:   # 1: windowed relaxation, 2: sequential chains (the default for this target)
            kw["schedule"] = schedule
        try:
            ref = LC.run_oracle(O, cfg, T, seed=sd, adapt=adapt, factor=factor, c=c)
        except O.BoundError:
            ref = None
        try:
            if gpu:
                out, _ = LC.run_device(z, cfg, T, seed=sd, adapt=adapt, factor=factor, c=c, tune=kw)
            else:
                out = O.window_sim(None, cfg["Gamma_drop"], 0.0, cfg["x0"], cfg["theta0"], T, c, mu=cfg["mu"], adapt=adapt, factor=factor,
                                   logistic=cfg["logistic"], seed=sd, **kw)
        except (O.BoundError, z.BoundError):
            out = None
        # error("Tuning parameter `c` too small.") must be raised by both sides or by neither
        assert (ref is None) == (out is None), (it, levels, r, m, dseed, k)
        if ref is None:
            n_err += 1
        else:
            O.assert_same_run(ref, out)
            n_ok += 1
    if verbose:
        print(f"logistic fuzz seed {seed}: {n_ok} cases bit-exact, {n_err} bound errors on both sides, {time.time() - t0:.1f} s")
    return n_ok, n_err


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--gpu"]
    z = graft.load_package()
    if "--gpu" in sys.argv:
        z.init(0)
    run(z, int(args[0]) if args else 1, int(args[1]) if len(args) > 1 else 200, gpu="--gpu" in sys.argv)
