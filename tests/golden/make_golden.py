"""Writes tests/golden/*.json from the CPU oracle (mode ctr|lazy) for the cases of tests/golden_cases.py.  These are
REGRESSION fixtures for our own RNG/arithmetic contract, not reference vectors: the reference holds no golden vector for
this path and cannot be executed here (no Julia) -- see oracle/zz_oracle.c."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as graft  # noqa: E402
import golden_cases as GC  # noqa: E402
import oracle_lib as O  # noqa: E402

zzb = graft.load_package()
for name in GC.CASES + GC.LOGISTIC_CASES:
    out = GC.run_oracle(O, GC.case_inputs(zzb, name))
    json.dump(out, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    print(name, out["num"], out["n_events"], out["acc_sum"])
