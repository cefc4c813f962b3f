"""Coordinate sharding of one sampler run over the GPUs of a node: one process per GPU, `torch.distributed` for the
plumbing (exchange of CUDA-IPC handles, host barriers, gathering of results), the data path entirely inside the
persistent kernels (peer loads / atomics over NVLink, mailbox all-reduce at pass boundaries).

Every rank calls :func:`spdmp_sharded` with the SAME global inputs; the complete result is returned on every rank.
"""
from __future__ import annotations

import numpy as np

from ._capi import EVENT_DTYPE


def shard_bounds(d: int, nranks: int, grid_m: int = 0):
    """[lo, hi) of every rank -- the rule of zzb_run_shard (csrc/zzb200.cpp): equal contiguous blocks, rounded up to
    whole lattice columns when the problem is a lattice."""
    shard = -(-d // nranks)
    if grid_m:
        shard = -(-shard // grid_m) * grid_m
    return [(min(d, shard * r), min(d, shard * (r + 1))) for r in range(nranks)]


def exchange_blobs(blob: bytes, group=None):
    """All-gather one opaque byte string per rank (the IPC handle block of zzb_run_ipc_export)."""
    import torch.distributed as dist

    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, blob, group=group)
    return out


def merge_shards(parts, d: int):
    """Combine per-rank results (each valid on its owned range) into the global result.
    parts: list over ranks of dicts with keys lo, hi, acc, num, t, x, theta, c, s1, s2, events."""
    out = {k: np.empty(d, dtype=parts[0][k].dtype) for k in ("acc", "t", "x", "theta", "c", "s1", "s2")}
    for p in parts:
        lo, hi = p["lo"], p["hi"]
        for k in out:
            out[k][lo:hi] = p[k][lo:hi]
    out["num"] = int(sum(p["num"] for p in parts))
    ev = np.concatenate([p["events"] for p in parts]) if parts else np.empty(0, dtype=EVENT_DTYPE)
    if len(ev):
        ev = ev[np.lexsort((ev["i"], ev["t"]))]  # per-rank traces are sorted; merge by (time, coordinate)
    out["events"] = ev
    return out


def spdmp_sharded(zzb, target, Z, t0, x0, theta0, T, c, *, seed=(1, 2), adapt=False, factor=1.8, record_trace=True,
                  tune=None, group=None, gather=True, run_kwargs=None):
    """Sharded `spdmp`: returns (result dict, run statistics, kernel milliseconds of this rank).  Must be called by every
    rank of `group` (default: the world), each after `zzb.init(local_rank)`."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    prob = zzb.Problem(target, Z)
    run = zzb.Run(prob, record_trace=record_trace, **(run_kwargs or {}))   # run_kwargs: kappa=... (sticky), boomerang=F (FactBoomerang)
    try:
        if tune:
            run.set(**tune)
        run.shard(rank, world)
        blobs = exchange_blobs(run.ipc_export(), group)
        for p, b in enumerate(blobs):
            if p != rank:
                run.ipc_import(p, b)
        dist.barrier(group)
        run.upload(t0, x0, theta0, c, seed=seed, adapt=adapt, factor=factor)
        dist.barrier(group)  # nobody may write into a peer's mailbox before that peer has reset it
        ms = run.execute(T)
        dist.barrier(group)
        lo, hi = run.owned_range()
        acc, num = run.counts()
        t, x, th, cc = run.final_state()
        s1, s2 = run.sums()
        part = dict(lo=lo, hi=hi, acc=acc, num=num, t=t, x=x, theta=th, c=cc, s1=s1, s2=s2,
                    events=run.events() if record_trace else np.empty(0, dtype=EVENT_DTYPE))
        stats = run.stats()
        if not gather:
            return part, stats, ms
        parts = [None] * world
        dist.all_gather_object(parts, part, group=group)
        return merge_shards(parts, prob.d), stats, ms
    finally:
        dist.barrier(group)  # peers may still be reading this rank's buffers
        run.close()
        prob.close()
