"""Writes tests/golden/gmrf16_T3.json from the CPU oracle (mode ctr|lazy).  These are REGRESSION fixtures for our own
RNG/arithmetic contract, not reference vectors: the reference holds no golden vector for this path and cannot be
executed here (no Julia) -- see oracle/zz_oracle.c."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import __graft_entry__ as graft  # noqa: E402
import oracle_lib as O  # noqa: E402

zzb = graft.load_package()
n, T, seed = 16, 3.0, (1, 2)
G, x0, th0, c = zzb.gmrf_config(n)
r = O.spdmp(G, G, 0.0, x0, th0, T, c, seed=seed)
out = dict(n=n, T=T, seed=list(seed), num=r.num, n_events=len(r.events), first_i=r.events["i"][:32].tolist(),
           first_t_hex=[float(t).hex() for t in r.events["t"][:8]], last_t_hex=float(r.events["t"][-1]).hex(),
           xor_t=int(np.bitwise_xor.reduce(r.events["t"].view(np.uint64))))
json.dump(out, open(os.path.join(HERE, "gmrf16_T3.json"), "w"), indent=1)
print(out)
