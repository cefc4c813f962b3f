import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
for n, T in ((100, 20.0), (316, 5.0)):
    G, x0, th0, c = z.gmrf_config(n)
    prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
    for grid in (148, 74, 37, 16, 8):
        for frac in (0.25, 0.5):
            run = z.Run(prob, record_trace=False); run.set(target_frac=frac, target_flip_frac=0.045 * frac / 0.25, grid=grid)
            run.upload(0.0, x0, th0, c, seed=(1, 2)); run.execute(T)
            run.reset(); ms = run.execute(T)
            acc, num = run.counts(); st = run.stats()
            print(f"n={n} grid={grid} frac={frac}: {ms:.2f} ms {acc.sum()/ms*1e3:.3e} sw/s windows {st['windows']} passes {st['passes']}", flush=True)
            run.close()
