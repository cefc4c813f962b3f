import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
tight = len(sys.argv) > 3 and sys.argv[3] == "tight"
fracs = [float(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0.03]
ffracs = [float(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [None]
G, x0, th0, c = z.gmrf_config(n, tight=tight)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
# warm the clocks up
w = z.Run(prob, record_trace=False); w.upload(0.0, x0, th0, c, seed=(1, 2)); w.execute(min(T, 2.0)); w.close()
for frac in fracs:
    for ff in ffracs:
        rec = False
        run = z.Run(prob, record_trace=rec); run.set(target_frac=frac)
        if ff is not None: run.set(target_flip_frac=ff)
        t0 = time.time(); run.upload(0.0, x0, th0, c, seed=(1, 2)); t1 = time.time()
        ms = run.execute(T); t2 = time.time()
        acc, num = run.counts(); nacc = int(acc.sum()); st = run.stats()
        print(f"n={n} T={T} tight={tight} frac={frac} ff={ff} trace={rec}: kernel {ms:.2f} ms, upload {1e3*(t1-t0):.1f} ms, exec wall {1e3*(t2-t1):.1f} ms; "
              f"{nacc} switches {num} proposals -> {nacc/ms*1e3:.3e} switches/s {num/ms*1e3:.3e} proposals/s; {st}; "
              f"us/pass {1e3*ms/max(st['passes'],1):.2f}, evals/prop {st['node_evals']/max(num,1):.2f}")
        run.close()
