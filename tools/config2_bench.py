import sys, os, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
z = g.load_package(); z.init(0)
G, x0, th0, c = z.gmrf_config(100)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
run = z.Run(prob, record_trace=False); run.set(schedule=1)
run.upload(0.0, x0, th0, c, seed=(1, 2)); run.execute(20.0); run.reset()
ms = run.execute(20.0); acc, num = run.counts()
print(f"config 2 (d = 10^4, T = 20): {ms:.2f} ms, {int(acc.sum())} switches -> {acc.sum() / ms * 1e3:.3e} switches/s, stats windows {run.stats()['windows']}")
