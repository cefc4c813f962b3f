import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
G, x0, th0, c = z.gmrf_config(1000)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
for et in (0, 288, 192, 96):
    run = z.Run(prob, record_trace=False); run.set(eval_threads=et); run.upload(0.0, x0, th0, c, seed=(1, 2))
    best = 1e9
    for rep in range(3):
        run.reset(); ms = run.execute(2.0); best = min(best, ms)
    st = run.stats()
    print(f"eval_threads {et}: {best:.3f} ms relax {st['ns_relax']/1e6:.2f} idle {st['ns_tail']/1e6:.2f} commit {st['ns_commit']/1e6:.2f} scan {st['ns_scan']/1e6:.2f} rounds {st['passes']}", flush=True)
    run.close()
