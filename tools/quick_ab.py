"""A/B of the two schedules on the bench workload (d = n^2 lattice GMRF): python tools/quick_ab.py [n] [T] [fracs] [schedules]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
fracs = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0.25]
scheds = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1]
tight = len(sys.argv) > 5 and sys.argv[5] == "tight"
G, x0, th0, c = z.gmrf_config(n, tight=tight)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
w = z.Run(prob, record_trace=False); w.upload(0.0, x0, th0, c, seed=(1, 2)); w.execute(min(T, 2.0)); w.close()
ref = None
for sched in scheds:
    for frac in fracs:
        run = z.Run(prob, record_trace=False); run.set(target_frac=frac, schedule=sched)
        if os.environ.get("FLIPSCALE"): run.set(target_flip_frac=float(os.environ["FLIPSCALE"]) * frac)
        run.upload(0.0, x0, th0, c, seed=(1, 2))
        best = 1e30
        for rep in range(3):
            run.reset(); ms = run.execute(T); best = min(best, ms)
        acc, num = run.counts(); nacc = int(acc.sum()); st = run.stats()
        t, x, th, cc = run.final_state()
        sig = (num, nacc, float(np.sum(x)), float(np.sum(t)))
        if ref is None: ref = sig
        keep = {k: st[k] for k in ("windows", "retries", "passes", "node_evals", "ns_scan", "ns_relax", "ns_tail", "ns_commit", "ns_barrier", "ns_phaseb", "n_barriers", "n_tail_passes", "dbg0", "dbg1", "dbg2", "dbg3", "dbg4", "dbg5", "dbg6", "dbg7")}
        print(f"sched={sched} n={n} T={T} frac={frac}: best {best:.3f} ms; {nacc} switches {num} proposals -> {nacc/best*1e3:.3e} switches/s; same={sig == ref}; {keep}", flush=True)
        run.close()
