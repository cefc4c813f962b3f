"""Context measurement for BASELINE.md section 1 (the only timings upstream publishes are for the sparse sticky ZigZag on the
n x n "heart" image-denoising GMRF, research/sticky/heart/heart_sparse2.jl): the same target family on the device's
dense sticky sampler `sspdmp`.  NOT the same sampler variant (upstream: sparsestickyzz with constant "strong" bounds and
an aggregated thaw clock), so the number is context, not a like-for-like comparison.

    python tools/sticky_bench.py [n] [T]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

z = g.load_package(); z.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 316
T = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
c1, c2, sigma2, kappa = 2.0, 0.1, 0.5, 0.15           # heart_sparse2.jl:33-38,91
L = z.grid_precision(n, n, shift=0.0)                  # graph Laplacian
G = z.CSC(L.n, L.colptr, L.rowval, L.nzval * c1)       # c1 * L ...
d = G.n
cols = np.repeat(np.arange(d), np.diff(G.colptr)); on = (G.rowval - 1) == cols
G.nzval = G.nzval.copy(); G.nzval[on] += c2 + 1 / sigma2   # ... + (c2 + 1/sigma^2) I
r1 = np.linspace(-4.5, 4.5, n); r2 = np.linspace(-4.1, 4.9, n)
X, Y = np.meshgrid(r1, r2, indexing="ij")
heart = 5 * np.maximum(1 - (X ** 2 + (5 * Y / 4 - np.sqrt(np.abs(X))) ** 2), 0)          # heart_sparse2.jl:55,66
rng = np.random.default_rng(1)
h = (heart + np.sqrt(sigma2) * rng.standard_normal((n, n))).reshape(-1, order="F") / sigma2   # linear term of grad phi
x0 = np.zeros(d); th0 = rng.choice(np.array([-1.0, 1.0]), d)
c = G.colnorms() + np.abs(h)            # Z.mu = 0 ignores the linear term: cover it with the constant (|theta| = 1)
prob = z.Problem(z.GaussianPotential(G, h), z.ZigZag(G, np.zeros(d)))
for rep in range(2):
    run = z.Run(prob, record_trace=False, kappa=np.full(d, kappa))
    run.upload(0.0, x0, th0, c, seed=(1, 2))
    ms = run.execute(T)
    acc, num = run.counts(); nev = run.n_events(); st = run.stats()
    t, x, th, _ = run.final_state()
    print(f"sticky heart-like GMRF n={n} (d={d}) T={T}: kernel {ms:.1f} ms, {nev} trace events (reflections {int(acc.sum())}), {num} proposals "
          f"-> {nev / ms * 1e3:.3e} events/s; active fraction at T {np.mean(th != 0):.3f}; windows {st['windows']} passes {st['passes']}")
    run.close()
# the reference's own sampler for this workload: the STRONG-BOUND sparse sticky ZigZag (src/sparsestickyzz.jl, sspdmp3; upstream runs
# it with c = 4.26, heart_sparse2.jl:127): zz_run_kernel_csr_strong.  The constant bound must dominate |grad_i| on the path.
cs = float(sys.argv[3]) if len(sys.argv) > 3 else 4.26
for rep in range(2):
    run = z.Run(prob, record_trace=False, kappa=np.full(d, kappa))
    run.set(strong_c=cs, strong_rule=0)
    run.upload(0.0, x0, th0, np.full(d, cs), seed=(1, 2))
    try:
        ms = run.execute(T)
    except z.BoundError as e:
        print("strong-bound sampler: bound too small:", str(e)[:100]); run.close(); break
    acc, num = run.counts(); nev = run.n_events(); st = run.stats()
    t, x, th, _ = run.final_state()
    print(f"STRONG-BOUND sparse sticky (sspdmp3) n={n} (d={d}) T={T} c={cs}: kernel {ms:.1f} ms, {nev} trace events (reflections {int(acc.sum())}), {num} proposals "
          f"-> {nev / ms * 1e3:.3e} events/s; active fraction at T {np.mean(th != 0):.3f}; windows {st['windows']} rounds {st['passes']}")
    run.close()
# BASELINE config 4 itself: the p = 10^5 chain of test/sparsesticky.jl:17-54 (x0 = 0, kappa = 2000/p, c = 2.5, T = 500)
import scipy.sparse as sp
p = 100000
main = np.full(p, 2.1); main[0] = main[-1] = 1.1
Gc = z.CSC.from_scipy(sp.diags([main, -np.ones(p - 1), -np.ones(p - 1)], [0, 1, -1]).tocsc())
probc = z.Problem(z.GaussianPotential(Gc), z.ZigZag(Gc, np.zeros(p)))
Tc = float(sys.argv[4]) if len(sys.argv) > 4 else 100.0
for name, strong in (("ss_fact kernel (sspdmp)", False), ("strong-bound kernel (sspdmp3)", True)):
    for rep in range(2):
        run = z.Run(probc, record_trace=False, kappa=np.full(p, 2000.0 / p))
        if strong:
            run.set(strong_c=2.5, strong_rule=0)
        run.upload(0.0, np.zeros(p), np.ones(p), np.full(p, 2.5), seed=(5, 6))
        ms = run.execute(Tc)
        acc, num = run.counts(); nev = run.n_events(); st = run.stats()
    print(f"config 4 chain p={p} T={Tc} {name}: kernel {ms:.1f} ms, {nev} trace events, {num} proposals -> {nev / ms * 1e3:.3e} events/s; "
          f"windows {st['windows']} rounds {st['passes']} evals {st['node_evals']}")
    run.close()
