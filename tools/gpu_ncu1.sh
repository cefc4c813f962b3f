cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as g
z = g.load_package(); z.init(0)
n = int(sys.argv[1]); T = float(sys.argv[2]); frac = float(sys.argv[3])
G, x0, th0, c = z.gmrf_config(n)
prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
run = z.Run(prob, record_trace=False); run.set(target_frac=frac)
run.upload(0.0, x0, th0, c, seed=(1, 2)); ms = run.execute(T); print(ms, run.stats())
PY
ncu --set full --clock-control none --import-source on -k regex:zz_run_kernel -c 1 -o gpurun_out/prof_small python /tmp/one.py 8 40.0 0.4 > gpurun_out/ncu_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zz_run_kernel -c 1 -o gpurun_out/prof_big python /tmp/one.py 1000 1.0 0.1 > gpurun_out/ncu_big.log 2>&1
tail -3 gpurun_out/ncu_small.log gpurun_out/ncu_big.log
