"""Host-side logic of the sharded path on CPU: shard rule, gloo all-gather of handle blobs, merging of shard results
(world_size 2, gloo)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds(zzb):
    assert zzb.shard_bounds(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert zzb.shard_bounds(1000 * 1000, 8, 1000) == [(125000 * r, 125000 * (r + 1)) for r in range(8)]
    b = zzb.shard_bounds(24 * 24, 5, 24)   # 576 coordinates, blocks of whole columns (5 columns = 120)
    assert all((lo % 24 == 0) for lo, hi in b) and b[-1][1] == 576 and b[0] == (0, 120)
    assert zzb.shard_bounds(7, 8) [-1] == (7, 7)  # more ranks than coordinates: empty tail shards


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as graft
    zzb = graft.load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    blobs = zzb.multigpu.exchange_blobs(bytes([rank]) * 512)
    assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 512 for b in blobs)
    d = 10
    lo, hi = zzb.shard_bounds(d, world)[rank]
    ev = np.zeros(3, dtype=zzb._capi.EVENT_DTYPE)
    ev["t"] = [0.1 + rank, 0.5 + rank * 0.01, 2.0 - rank]
    ev["i"] = [lo + 1, lo + 2, lo + 1]
    ev = ev[np.argsort(ev["t"])]
    full = lambda v: np.full(d, float(v))
    part = dict(lo=lo, hi=hi, acc=np.full(d, rank + 1, np.int64), num=10 * (rank + 1), t=full(rank), x=full(rank), theta=full(rank),
                c=full(rank), s1=full(rank), s2=full(rank), events=ev)
    parts = [None] * world
    dist.all_gather_object(parts, part)
    res = zzb.merge_shards(parts, d)
    assert res["num"] == 30 and np.all(np.diff(res["events"]["t"]) >= 0) and len(res["events"]) == 6
    assert res["acc"][:5].tolist() == [1] * 5 and res["acc"][5:].tolist() == [2] * 5 and res["x"][7] == 1.0
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_gloo_two_ranks_exchange_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]
