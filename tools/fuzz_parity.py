"""Long randomised parity campaign (development aid): python tools/fuzz_parity.py [n_cases] [seed] [--sim]
--sim: the host emulation of the device schedule instead of the CUDA path (no GPU needed)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import fuzz_cases

SIM = "--sim" in sys.argv
sys.argv = [a for a in sys.argv if a != "--sim"]
z = g.load_package()
if not SIM:
    z.init(0)
ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time()
bad, nbound = fuzz_cases.run_cases(z, ncases, seed, verbose=True, sim=SIM)
print(f"{ncases} cases, {bad} failures, {nbound} bound errors, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
