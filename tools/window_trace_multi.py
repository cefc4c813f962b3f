"""torchrun helper: per-CTA round log of ONE window of a sharded run; writes gpurun_out/wtm_<rank>.bin (see window_trace.py show)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
os.environ["ZZB200_DBG_WINDOW"] = sys.argv[1] if len(sys.argv) > 1 else "12"
os.environ["ZZB200_DBG_FILE"] = "gpurun_out/wtm_%d.bin" % local
import __graft_entry__ as g
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
z = g.load_package(); z.init(local)
G, x0, th0, c = z.gmrf_config(1000)
for rep in range(2):
    part, st, ms = z.spdmp_sharded(z, z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)), 0.0, x0, th0, 2.0, c, seed=(1, 2), record_trace=False, gather=False)
    dist.barrier()
print("rank", dist.get_rank(), "kernel ms", ms, flush=True)
dist.destroy_process_group()
