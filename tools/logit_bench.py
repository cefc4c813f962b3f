"""Config 3 (scripts/logistic.jl, n = 8840, p = 442) on the device: time of the event loop for R independent replicas run as one
block-diagonal problem, next to the single-threaded oracle.  python tools/logit_bench.py [T] [R ...]
SCHEDULE=1 in the environment times the windowed relaxation instead of the sequential chains (the default for this target)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
import logistic_cases as LC  # noqa: E402
import oracle_lib as O  # noqa: E402

z = graft.load_package()
z.init(0)
T = float(sys.argv[1]) if len(sys.argv) > 1 else 25.0
Rs = [int(a) for a in sys.argv[2:]] or [1, 8, 64]
cfg = LC.make(z, *LC.FULL)
ref = LC.run_oracle(O, cfg, T)
print(f"oracle (1 thread, 1 chain): {len(ref.events)} events, {ref.num} proposals in {ref.loop_seconds * 1e3:.1f} ms "
      f"= {len(ref.events) / ref.loop_seconds:.3g} events/s", flush=True)
for R in Rs:
    t0 = time.time()
    big = cfg if R == 1 else LC.replicas(z, cfg, R)
    t1 = time.time()
    got, Xi = LC.run_device(z, big, T, tune=dict(schedule=int(os.environ.get("SCHEDULE", "2")), seq_warps=int(os.environ.get("SEQ_WARPS", "0"))))
    st = Xi.stats
    print(f"R = {R}: d = {big['p']}, {len(got.events)} events, {got.num} proposals in {Xi.device_ms:.1f} ms on the device = "
          f"{len(got.events) / (Xi.device_ms * 1e-3):.3g} events/s; windows {st['windows']}, passes {st['passes']}, evaluations {st['node_evals']}, "
          f"tail passes {st['n_tail_passes']} ({st['ns_tail'] * 1e-6:.1f} ms), scan {st['ns_scan'] * 1e-6:.1f} ms, relax {st['ns_relax'] * 1e-6:.1f} ms; "
          f"host setup {t1 - t0:.1f} s, call {time.time() - t1:.1f} s", flush=True)
