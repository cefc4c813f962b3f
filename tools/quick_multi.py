"""torchrun helper: sharded run of the d = n^2 lattice, prints per-rank phase timers (development tool)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = g.load_package(); z.init(local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
G, x0, th0, c = z.gmrf_config(n)
for rep in range(2):
    part, st, ms = z.spdmp_sharded(z, z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)), 0.0, x0, th0, T, c, seed=(1, 2),
                                   record_trace=False, gather=False, tune=({"target_frac": float(os.environ["FRAC"]), "target_flip_frac": 0.18 * float(os.environ["FRAC"])} if os.environ.get("FRAC") else None))
    dist.barrier()
    if rep == 1:
        keys = ("windows", "passes", "ns_scan", "ns_relax", "ns_tail", "ns_commit", "ns_barrier", "ns_phaseb", "n_barriers")
        print(f"rank {dist.get_rank()}: kernel {ms:.2f} ms, owned [{part['lo']},{part['hi']}) acc {int(part['acc'][part['lo']:part['hi']].sum())} "
              + " ".join(f"{k}={st[k]/1e6:.2f}ms" if k.startswith("ns_") else f"{k}={st[k]}" for k in keys), flush=True)
dist.destroy_process_group()
