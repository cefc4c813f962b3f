# compute-sanitizer memcheck + synccheck over small runs of every kernel family (slow: the persistent kernel spins on barriers)
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
import golden_cases as GC
z = g.load_package(); z.init(0)
for name in GC.CASES:
    k = GC.case_inputs(z, name)
    k["T"] = min(k["T"], 1.0 if name != "spd8_adapt" else 5.0)
    r = GC.run_device(z, k)
    print(name, r["num"], r["n_events"], flush=True)
G, x0, th0, c = z.gmrf_config(40)
Xi, _, (acc, num), _ = z.spdmp(z.GaussianPotential(G), 0.0, x0, th0, 0.5, c, z.ZigZag(G, np.zeros(G.n)), seed=(1, 2), discretize_dt=0.1)
print("grid", num, Xi.grid[1].shape, flush=True)
PY
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -25
echo "exit $?"
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -4
# (racecheck is not usable here: it does not keep the CTAs of the cooperative launch co-resident, so the software grid
#  barrier of the persistent kernel hangs or the run goes wrong under the tool)
