import os, sys
sys.path.insert(0, "/root/repo")
os.environ["ZZB200_DBG_WINDOW"] = "9999"
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
G, x0, th0, c = z.gmrf_config(1000)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
run = z.Run(prob, record_trace=False); run.upload(0.0, x0, th0, c, seed=(1, 2))
for rep in range(12):
    os.environ["ZZB200_DBG_FILE"] = "gpurun_out/osc_%d.bin" % rep
    run.reset(); ms = run.execute(2.0); st = run.stats()
    print(rep, f"{ms:.3f} ms retries {st['retries']} tagovf {st['dbg1']} logged {st['dbg7']}", flush=True)
    if st['dbg1'] == 0: os.remove("gpurun_out/osc_%d.bin" % rep)
    elif st['dbg1'] >= 1 and rep >= 2: break
