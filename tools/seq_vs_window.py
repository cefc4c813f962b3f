"""Small Gaussian problems: event-loop time of the sequential-chain schedule (one warp, zz_seq_kernel) next to the windowed
relaxation (development tool; decides where the automatic schedule should switch).  python tools/seq_vs_window.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

z = graft.load_package()
z.init(0)


def timed(G, x0, th0, c, T, schedule):
    prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
    run = z.Run(prob, record_trace=False)
    run.set(schedule=schedule)
    run.upload(0.0, x0, th0, c, seed=(1, 2), adapt=True)
    run.execute(T)
    run.reset()
    run.execute(T)
    acc, num = run.counts()
    ms = run.device_ms
    run.close(); prob.close()
    return ms, int(acc.sum()), int(num)


cases = []
G2 = z.random_spd(2, density=1.0, seed=3)
cases.append(("dense d=2", G2, 2000.0))
G8 = z.random_spd(8, density=1.0, seed=2)
cases.append(("dense d=8", G8, 1000.0))
cases.append(("dense d=32", z.random_spd(32, density=1.0, seed=2), 200.0))
for n, T in ((4, 400.0), (8, 200.0), (16, 100.0), (32, 50.0), (48, 30.0)):
    cases.append((f"lattice d={n * n}", z.grid_precision(n, n), T))
for name, G, T in cases:
    rng = np.random.default_rng(1)
    x0, th0 = rng.standard_normal(G.n), rng.choice(np.array([-1.0, 1.0]), G.n)
    c = G.colnorms()
    out = []
    for s in (1, 2):
        ms, na, num = timed(G, x0, th0, c, T, s)
        out.append((ms, na, num))
    assert out[0][1:] == out[1][1:]
    print(f"{name:18s} T={T:6g}: {out[0][2]:8d} proposals, {out[0][1]:7d} events; windowed {out[0][0]:8.2f} ms ({out[0][0] * 1e3 / out[0][2]:6.2f} us/proposal), "
          f"sequential {out[1][0]:8.2f} ms ({out[1][0] * 1e3 / out[1][2]:6.2f} us/proposal)", flush=True)
