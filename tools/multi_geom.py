"""torchrun helper: lattice of m rows x n columns sharded over the ranks (per-rank geometry of a bigger run on fewer GPUs)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
import __graft_entry__ as g
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = g.load_package(); z.init(local)
m, n, T = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
G = z.grid_precision(m, n, shift=0.01)
rng = np.random.default_rng(1)
d = G.n
x0 = rng.standard_normal(d); th0 = rng.choice(np.array([-1.0, 1.0]), d); c = G.colnorms()
for rep in range(2):
    part, st, ms = z.spdmp_sharded(z, z.GaussianPotential(G), z.ZigZag(G, np.zeros(d)), 0.0, x0, th0, T, c, seed=(1, 2), record_trace=False, gather=False, tune=({"target_frac": float(os.environ["FRAC"]), "target_flip_frac": 0.18 * float(os.environ["FRAC"])} if os.environ.get("FRAC") else None))
    dist.barrier()
    print(f"rep {rep} rank {dist.get_rank()}: kernel {ms:.2f} ms windows {st['windows']} retries {st['retries']} rounds {st['passes']} relax {st['ns_relax']/1e6:.2f} xchg {st['ns_phaseb']/1e6:.2f} idle {st['ns_tail']/1e6:.2f} dbg {[st['dbg%d' % q] for q in range(4)]}", flush=True)
dist.destroy_process_group()
