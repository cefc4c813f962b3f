import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
G, x0, th0, c = z.gmrf_config(1000)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
run = z.Run(prob, record_trace=False); run.upload(0.0, x0, th0, c, seed=(1, 2))
for rep in range(8):
    run.reset(); ms = run.execute(2.0); st = run.stats()
    d2 = st['dbg2']
    print(f"{ms:.3f} ms windows {st['windows']} retries {st['retries']} rounds0 {st['passes']} evals {st['node_evals']} tag>=12: {st['dbg0']} tagovf {st['dbg1']} j={d2 & 0xffffffff} rounds={((d2>>32)&0xffff)} nflip={d2>>48} idle {st['ns_tail']/1e6:.2f}", flush=True)
