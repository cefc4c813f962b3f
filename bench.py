#!/usr/bin/env python
"""bench.py -- switching events/sec of the local ZigZag on the d = 10^6 grid-Laplacian GMRF (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (B200 kernels through the C-ABI)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle's restatement of the reference

A "step" is one complete `spdmp` run over [0, T_step] from the same synthetic initial state (Gamma = 0.01 I +
gridlaplacian(n, n), x0 ~ N(0,1), theta0 = +-1, c = ||Gamma[:, i]||_2 as in scripts/gaussianrandomfield.jl:14-42):
  value   steps with the inputs resident in HBM: re-initialisation kernels + the persistent event-loop kernel,
          timed with CUDA events on the launching stream;
  e2e     the same step through the host-buffer call path: H2D of (x0, theta0, c) from pinned memory, the kernels,
          D2H of the counters / final state / moment sums, all inside the timed region.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as graft  # noqa: E402

METRIC = "switching events/sec, d=10^6 sparse-GMRF local ZigZag"
# algorithmic bytes of SURVEY.md 8(d) (5-point grid, implicit Gamma): 208 B per proposal, +456 B per accepted one
B_PROPOSAL, B_ACCEPT = 208.0, 456.0


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev=0):
        self.dev, self.rows, self.proc = dev, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args):
    zzb = graft.load_package()
    G, x0, th0, c = zzb.gmrf_config(args.n, seed=1, tight=args.tight)
    return zzb, G, x0, th0, c


def workload_string(args, d):
    """The ONE description of the bench workload, identical in both arms (the driver compares the strings)."""
    return (f"local ZigZag spdmp, {args.n}x{args.n} grid GMRF (d={d}), c={'sqrt(eps)' if args.tight else '||Gamma[:,i]||'}, "
            f"T_step={args.T}, one step = full run from (x0, theta0)")


def run_signature(acc, num, t, x, theta, c=None, s1=None, s2=None):
    """Order-independent fingerprint of a finished run (bit patterns, not rounded values)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.int64(num).tobytes())
    for a in (acc, t, x, theta, c, s1, s2):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def oracle_parity(O, G, x0, th0, c, T, got):
    """The bench workload once more on the CPU oracle in the device's contract mode (ctr | lazy), compared bit for bit with
    what the device produced (`got`: dict with acc, num, t, x, theta, c, s1, s2 and optionally events).  Outside every timed region."""
    ref = O.spdmp(G, G, 0.0, x0, th0, T, c, seed=(1, 2))
    out = {"oracle": "zzo_spdmp mode ctr|lazy (oracle/zz_oracle.c), same inputs, T = %g" % T,
           "num_equal": int(ref.num) == int(got["num"]),
           "acc_equal": bool(np.array_equal(ref.acc, got["acc"])),
           "final_state_equal": all(bool(np.array_equal(getattr(ref, f).view(np.uint64), got[f].view(np.uint64))) for f in ("t", "x", "theta", "c")),
           "moment_sums_equal": bool(np.array_equal(ref.s1.view(np.uint64), got["s1"].view(np.uint64)) and
                                     np.array_equal(ref.s2.view(np.uint64), got["s2"].view(np.uint64))),
           "switches": int(ref.acc.sum()), "proposals": int(ref.num),
           "signature_oracle": run_signature(ref.acc, ref.num, ref.t, ref.x, ref.theta, ref.c, ref.s1, ref.s2),
           "signature_device": run_signature(got["acc"], got["num"], got["t"], got["x"], got["theta"], got["c"], got["s1"], got["s2"])}
    if got.get("events") is not None:
        ev = got["events"]
        out["events_equal"] = bool(len(ev) == len(ref.events) and np.array_equal(ev["i"], ref.events["i"]) and
                                   all(np.array_equal(ev[f].view(np.uint64), ref.events[f].view(np.uint64)) for f in ("t", "x", "theta")))
    out["ok"] = all(v for k, v in out.items() if k.endswith("_equal"))
    return out


def cpu_reference_step(O, G, x0, th0, c, T, seed):
    """One step on the host, single thread: the oracle's faithful restatement of spdmp (binary heap, single xoroshiro
    stream, in-place moves; src/sfact.jl:162-212).  Returns (switches, proposals, seconds of the event loop, setup excluded)."""
    r = O.spdmp(G, G, 0.0, x0, th0, T, c, seed=seed, mode=O.RNG_SEQ | O.ARITH_INPLACE)
    return int(r.acc.sum()), int(r.num), r.loop_seconds


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def parallel_chunks(d, cores):
    """Chunk counts to try for parallel_spdmp: Partition(nt, n) needs nt | n (src/parallel.jl:27); one host thread runs the
    outer task, so prefer nt <= cores - 1 but also try nt <= cores."""
    divs = [k for k in range(1, cores + 1) if d % k == 0]
    cand = {max(divs)}
    lower = [k for k in divs if k <= max(1, cores - 1)]
    if lower:
        cand.add(max(lower))
    return sorted(cand)


def cpu_parallel_step(O, G, G2, K, x0, th0, c, T, seed, delta):
    """One step on the host, K worker threads + the outer task: the restatement of the reference's multithreaded
    parallel_spdmp (src/parallel.jl) with the bound matrix restricted to the chunks (G2)."""
    r = O.spdmp(G, G2, 0.0, x0, th0, T, c, seed=seed, parallel=(K, delta), adapt=True)
    return int(r.acc.sum()), int(r.num), r.loop_seconds


def cpu_baselines(O, G, x0, th0, c, T, delta=0.02):
    """Single-thread spdmp and multithreaded parallel_spdmp on the same workload; returns the cpu_baseline object (the better
    of the two as `value`, both stated)."""
    d = G.n
    cores = host_cores()
    a, b, s = cpu_reference_step(O, G, x0, th0, c, T, (1, 2))
    single = dict(value=a / s, proposals_per_s=b / s, seconds=s, switches=a)
    best = None
    for K in parallel_chunks(d, cores):
        if K == 1:
            continue
        G2 = O.block_diagonal(G, K)
        a2, b2, s2 = cpu_parallel_step(O, G, G2, K, x0, th0, c, T, (1, 2), delta)
        if best is None or a2 / s2 > best["value"]:
            best = dict(value=a2 / s2, proposals_per_s=b2 / s2, seconds=s2, switches=a2, threads=K + 1, chunks=K)
    use_par = best is not None and best["value"] > single["value"]
    top = best if use_par else single
    return {
        "value": top["value"], "unit": "events/s", "cores": top.get("threads", 1), "kind": "port",
        "sample": (f"spdmp over [0,{T}] on the same d={d} GMRF, event loop only (setup excluded): "
                   + (f"multithreaded parallel_spdmp restatement (src/parallel.jl), {best['chunks']} chunk threads + outer task, Delta={delta}, "
                      f"{best['switches']} switches in {best['seconds']:.2f} s; " if best else "")
                   + f"single-thread spdmp restatement (src/sfact.jl) {single['switches']} switches in {single['seconds']:.2f} s; host has {cores} cores"),
        "proposals_per_s": top["proposals_per_s"], "seconds": top["seconds"],
        "single_thread_events_per_s": single["value"],
        "multithread_events_per_s": best["value"] if best else None,
        "multithread_threads": best["threads"] if best else None,
        "host_cores": cores,
    }


def run_reference(args):
    """CPU arm: the restatement of the reference's own CPU implementation of the path on the box's host cores.  Exactly
    `--warmup` untimed and `--steps` timed steps; a step is one run of the faster of (single-thread spdmp, multithreaded
    parallel_spdmp on all host threads) over the bounded sample [0, cpu_T] of the workload; `value` = events of the timed steps /
    their event-loop seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O
    zzb, G, x0, th0, c = workload(args)
    d = G.n
    T = args.cpu_T if args.cpu_T else min(args.T, 1.0)
    cores = host_cores()
    delta = 0.02
    # which variant is faster on this box?  (one probe each, untimed)
    a1, b1, s1 = cpu_reference_step(O, G, x0, th0, c, min(T, 0.25), (1, 2))
    cand = [("single", None, None, a1 / s1)]
    for K in parallel_chunks(d, cores):
        if K > 1:
            G2 = O.block_diagonal(G, K)
            a2, b2, s2 = cpu_parallel_step(O, G, G2, K, x0, th0, c, min(T, 0.25), (1, 2), delta)
            cand.append(("parallel", K, G2, a2 / s2))
    kind, K, G2, _ = max(cand, key=lambda r: r[3])

    def step():
        if kind == "single":
            return cpu_reference_step(O, G, x0, th0, c, T, (1, 2))
        return cpu_parallel_step(O, G, G2, K, x0, th0, c, T, (1, 2), delta)

    for _ in range(args.warmup):
        step()
    ev = pr = 0
    sec = 0.0
    for _ in range(args.steps):
        a, b, sdt = step()
        ev += a; pr += b; sec += sdt
    val = ev / sec
    single_val = a1 / s1
    cb = {"value": val, "unit": "events/s", "cores": (K + 1) if kind == "parallel" else 1, "kind": "port",
          "sample": (f"{args.steps} runs of spdmp over [0,{T}] on the same d={d} GMRF, event loop only (setup excluded): "
                     + (f"multithreaded parallel_spdmp restatement (src/parallel.jl), {K} chunk threads + outer task, Delta={delta}"
                        if kind == "parallel" else "single-thread spdmp restatement (src/sfact.jl), faithful draw order")
                     + f"; {ev} switches in {sec:.2f} s; host has {cores} cores"),
          "proposals_per_s": pr / sec, "seconds": sec, "variant": kind,
          "single_thread_events_per_s_probe": single_val, "probe_events_per_s": {k: v for k, _, _, v in cand}, "host_cores": cores}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args, d)},
        "note": "CPU restatement of the reference (Julia is not installed here or on the GPU box); a step is the bounded sample described in cpu_baseline.sample",
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    zzb, G, x0, th0, c = workload(args)
    zzb.init(local)
    from zzb200 import _capi

    if world > 1:
        return run_ours_sharded(args, zzb, G, x0, th0, c, world, rank, local)

    d = G.n
    prob = zzb.Problem(zzb.GaussianPotential(G), zzb.ZigZag(G, np.zeros(d)))
    run = zzb.Run(prob, record_trace=False)
    run.set(target_frac=args.frac)
    run.upload(0.0, x0, th0, c, seed=(1, 2))

    def step():
        run.reset()
        return run.execute(args.T)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    kernel_ms = 0.0
    with ClockSampler(local) as clk:
        _capi.event_record(0)
        for _ in range(args.steps):
            kernel_ms += step()
        _capi.event_record(1)
        total_ms = _capi.event_elapsed_ms()
        torch.cuda.synchronize()
    acc, num = run.counts()
    nacc = int(acc.sum())
    st = run.stats()
    value = nacc * args.steps / (total_ms * 1e-3)
    launches_per_step = 3  # zz_setup_kernel, zz_init_kernel, zz_run_kernel_grid (one cooperative launch)

    # ---- e2e: host buffers in pinned memory, H2D + kernels + D2H inside the timed region
    px0 = torch.from_numpy(x0).pin_memory().numpy()
    pth = torch.from_numpy(th0).pin_memory().numpy()
    pc = torch.from_numpy(c).pin_memory().numpy()
    run2 = zzb.Run(prob, record_trace=False)
    run2.set(target_frac=args.frac)
    e_steps = max(1, min(args.steps, 5))
    pin = lambda dt: torch.empty(d, dtype=dt).pin_memory().numpy()
    out = dict(t=pin(torch.float64), x=pin(torch.float64), theta=pin(torch.float64), c=pin(torch.float64),
               acc=pin(torch.int64), s1=pin(torch.float64), s2=pin(torch.float64))
    for _ in range(2):
        run2.upload(0.0, px0, pth, pc, seed=(1, 2)); run2.execute(args.T); run2.fetch_into(**out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        run2.upload(0.0, px0, pth, pc, seed=(1, 2))       # H2D of x0, theta0, c + initialisation kernels
        run2.execute(args.T)                              # the persistent event-loop kernel
        n2, a2 = run2.fetch_into(**out)                   # D2H of final state, adapted c, counts, moment sums
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    assert a2 == nacc and n2 == num
    e2e_val = a2 * e_steps / e_dt
    h2d = 3 * d * 8
    d2h = 7 * d * 8  # t, x, theta, c, acc, s1, s2

    # ---- e2e WITH the sampler's primary output: the full trace (32 B per switching event) ordered by time in host memory
    e2e_trace = None
    trace_events = None
    if not args.no_trace:
        run4 = zzb.Run(prob, record_trace=True)
        run4.set(target_frac=args.frac)
        t_steps = max(1, min(args.steps, 3))
        pinned_ev = torch.empty(int(1.25 * nacc) * 32 + 4096, dtype=torch.uint8).pin_memory().numpy().view(_capi.EVENT_DTYPE)
        for _ in range(1):
            run4.upload(0.0, px0, pth, pc, seed=(1, 2)); run4.execute(args.T); evbuf = run4.events(out=pinned_ev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(t_steps):
            run4.upload(0.0, px0, pth, pc, seed=(1, 2))
            run4.execute(args.T)
            n4, a4 = run4.fetch_into(**out)
            evbuf = run4.events(out=pinned_ev)             # events ordered by (time, coordinate) on the device, D2H into pinned memory
        torch.cuda.synchronize()
        t_dt = time.perf_counter() - t0
        assert a4 == nacc and len(evbuf) == nacc
        e2e_trace = {"value": a4 * t_steps / t_dt, "unit": "events/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h + 32 * len(evbuf), "steps": t_steps, "ms_per_step": 1e3 * t_dt / t_steps,
                     "what": "as e2e, plus the complete trace (t, i, x, theta per event): ordered by (time, coordinate) on the device "
                             "(zz_tsort_* kernels), copied into pinned host memory"}
        trace_events = evbuf
        run4.close()

    # ---- parity of what was timed: the same workload on the CPU oracle (contract mode), bit for bit; outside the timed regions
    parity = None
    if not args.no_parity:
        import oracle_lib as O
        got = dict(acc=out["acc"].copy(), num=n2, t=out["t"], x=out["x"], theta=out["theta"], c=out["c"], s1=out["s1"], s2=out["s2"],
                   events=trace_events)
        parity = oracle_parity(O, G, x0, th0, c, args.T, got)

    # ---- roofline of the dominant kernel (the persistent event loop)
    peak, peak_src = peaks()
    alg_bytes = B_PROPOSAL * num + B_ACCEPT * nacc        # per launch
    k_ms = kernel_ms / args.steps
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1)
    cpu = None
    if not args.no_cpu:
        import oracle_lib as O
        cpu = cpu_baselines(O, G, x0, th0, c, args.cpu_T if args.cpu_T else args.T)

    # ---- the same step with the tight bound c = sqrt(eps) (scripts/example.jl:39; acceptance ~ 1): reported beside the headline
    tight = None
    if not args.tight and not args.no_tight:
        ct = np.full(d, np.sqrt(np.finfo(np.float64).eps))
        run3 = zzb.Run(prob, record_trace=False)
        run3.set(target_frac=args.frac)
        run3.upload(0.0, x0, th0, ct, seed=(1, 2))
        for _ in range(2):
            run3.reset(); run3.execute(args.T)
        _capi.event_record(0)
        for _ in range(3):
            run3.reset(); run3.execute(args.T)
        _capi.event_record(1)
        t_ms = _capi.event_elapsed_ms() / 3
        acc3, num3 = run3.counts()
        tight = {"c": "sqrt(eps)", "events_per_s": int(acc3.sum()) / (t_ms * 1e-3), "ms_per_step": t_ms,
                 "switches_per_step": int(acc3.sum()), "proposals_per_step": int(num3)}
        run3.close()

    # ---- config 3 of BASELINE.json (sparse logistic regression, n = 8840, p = 442, subsampled ZigZag) as replicas side by side:
    # reported beside the headline, not part of it (sequential chains)
    config3 = None
    if not args.no_config3:
        try:
            config3 = config3_leg(zzb, args.config3_replicas)
        except Exception as e:   # (never let the side measurement take the headline down)
            config3 = {"error": f"{type(e).__name__}: {e}"}

    out = {
        "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "other_configs": {"config3_logistic_replicas": config3},
        "config": {"workload": workload_string(args, d),
                   "l2": "device working set (state + flip lists + work lists ~ 0.4 KB/coordinate = 400 MB) exceeds the 126 MB L2",
                   "switches_per_step": nacc, "proposals_per_step": int(num), "proposals_per_s": num * args.steps / (total_ms * 1e-3),
                   "windows_per_step": st["windows"], "rounds_of_cta0_per_step": st["passes"], "timeline_evaluations_per_step": st["node_evals"],
                   "target_frac": args.frac, "schedule": "asynchronous tile-local relaxation (DESIGN.md section 3b)",
                   "tight_bound_variant": tight},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "zz_run_kernel_grid", "kernel_ms_per_launch": k_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "latency-bound sparse event loop: see DESIGN.md (roofline) for why frac is small"},
        "e2e": {"value": e2e_val, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e_steps,
                "ms_per_step": 1e3 * e_dt / e_steps,
                "what": "upload of (x0, theta0, c) from pinned host memory + initialisation + event loop + D2H of final state, adapted c, counts and moment sums (no trace: see e2e_trace)"},
        "e2e_trace": e2e_trace,
        "parity_check": parity,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk.summary(),
    }
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_ours_sharded(args, zzb, G, x0, th0, c, world, rank, local):
    """N > 1: the d coordinates are sharded over the N GPUs (strong scaling: the problem does not grow).  Every rank
    times its own persistent kernel with CUDA events; the step time is the maximum over ranks."""
    import torch
    import torch.distributed as dist
    from zzb200 import _capi

    d = G.n
    prob = zzb.Problem(zzb.GaussianPotential(G), zzb.ZigZag(G, np.zeros(d)))
    run = zzb.Run(prob, record_trace=False)
    run.set(target_frac=args.frac)
    run.shard(rank, world)
    blobs = zzb.multigpu.exchange_blobs(run.ipc_export())
    for p, b in enumerate(blobs):
        if p != rank:
            run.ipc_import(p, b)
    dist.barrier()
    run.upload(0.0, x0, th0, c, seed=(1, 2))

    def step():
        dist.barrier()
        run.reset()
        dist.barrier()          # mailboxes of every rank are clean before anybody starts
        return run.execute(args.T)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    kernel_ms = 0.0
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            kernel_ms += step()
        torch.cuda.synchronize()
        dist.barrier()
        wall = time.perf_counter() - t0
    acc, num = run.counts()
    lo, hi = run.owned_range()
    tot = torch.tensor([float(acc[lo:hi].sum()), float(num), kernel_ms], dtype=torch.float64, device="cuda")
    mx = tot.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    nacc, nprop = int(tot[0].item()), int(tot[1].item())
    k_ms = mx[2].item() / args.steps          # device time of the slowest rank's event-loop kernel
    st = run.stats()
    # e2e through the host-buffer path: every rank uploads its slab (+ one halo column on either side) from pinned memory, runs,
    # and reads back the results of its OWNED coordinates -- the node moves ~ d values each way whatever N is
    e_steps = max(1, min(args.steps, 5))
    px0 = torch.from_numpy(x0).pin_memory().numpy(); pth = torch.from_numpy(th0).pin_memory().numpy(); pc = torch.from_numpy(c).pin_memory().numpy()
    pin = lambda dt: torch.empty(d, dtype=dt).pin_memory().numpy()
    outb = dict(t=pin(torch.float64), x=pin(torch.float64), theta=pin(torch.float64), c=pin(torch.float64),
                acc=pin(torch.int64), s1=pin(torch.float64), s2=pin(torch.float64))
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        dist.barrier()
        run.upload(0.0, px0, pth, pc, seed=(1, 2))
        dist.barrier()
        run.execute(args.T)
        run.fetch_into(**outb)
    torch.cuda.synchronize()
    dist.barrier()
    e_dt = time.perf_counter() - t0
    m_rows = args.n
    h2d_rank = 3 * 8 * (min(d, hi + m_rows) - max(0, lo - m_rows))
    d2h_rank = 7 * 8 * (hi - lo)
    io = torch.tensor([float(h2d_rank), float(d2h_rank)], dtype=torch.float64, device="cuda")
    dist.all_reduce(io, op=dist.ReduceOp.SUM)
    h2d_total, d2h_total = int(io[0].item()), int(io[1].item())
    # ---- context: the same N GPUs running N INDEPENDENT chains of the full problem (how multi-GPU MCMC is usually run; no
    # exchange at all).  Reported beside the sharded headline, never instead of it.
    runc = zzb.Run(prob, record_trace=False)
    runc.set(target_frac=args.frac)
    runc.upload(0.0, x0, th0, c, seed=(1 + rank, 2))
    for _ in range(2):
        runc.reset(); runc.execute(args.T)
    torch.cuda.synchronize()
    dist.barrier()
    c_ms = 0.0
    for _ in range(3):
        runc.reset(); c_ms += runc.execute(args.T)
    accc, _numc = runc.counts()
    chains = torch.tensor([float(accc.sum()) * 3, c_ms], dtype=torch.float64, device="cuda")
    chains_max = chains.clone()
    dist.all_reduce(chains, op=dist.ReduceOp.SUM)
    dist.all_reduce(chains_max, op=dist.ReduceOp.MAX)
    chains_events_per_s = chains[0].item() / (chains_max[1].item() * 1e-3)
    runc.close()

    # ---- parity of what was timed: gather the owned parts, compare with the CPU oracle (contract mode) bit for bit on rank 0
    parity = None
    if not args.no_parity:
        t_, x_, th_, c_ = run.final_state(); s1_, s2_ = run.sums()
        part = dict(lo=lo, hi=hi, num=num, **{k: np.ascontiguousarray(v[lo:hi]) for k, v in
                                               (("acc", acc), ("t", t_), ("x", x_), ("theta", th_), ("c", c_), ("s1", s1_), ("s2", s2_))})
        parts = [None] * world
        dist.all_gather_object(parts, part)      # owned slices only: ~ 7 * 8 * d bytes over the whole node
        if rank == 0:
            import oracle_lib as O
            got = {k: np.concatenate([p_[k] for p_ in parts]) for k in ("acc", "t", "x", "theta", "c", "s1", "s2")}
            got["num"] = int(sum(p_["num"] for p_ in parts))
            got["events"] = None
            parity = oracle_parity(O, G, x0, th0, c, args.T, got)
    peak, peak_src = peaks()
    alg_bytes = B_PROPOSAL * nprop + B_ACCEPT * nacc
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    if rank == 0:
        value = nacc * args.steps / (k_ms * args.steps * 1e-3)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": k_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_string(args, d),
                       "sharding": f"coordinates sharded by whole lattice columns over {world} GPUs (one process per GPU)",
                       "l2": "per-GPU working set 400 MB (full-length arrays on every rank) exceeds the 126 MB L2",
                       "switches_per_step": nacc, "proposals_per_step": nprop, "windows_per_step": st["windows"],
                       "rounds_of_cta0_per_step": st["passes"], "wall_ms_per_step_incl_host_barriers": 1e3 * wall / args.steps,
                       "exchange": "asynchronous tile-local relaxation on every GPU; marks that cross a GPU boundary are pushed into the "
                                   "owner's inbox and changed halo records into its replica over NVLink (plain stores); two node-wide "
                                   "barriers per WINDOW (mailbox all-reduce), none per pass or per event",
                       "independent_chains_events_per_s": chains_events_per_s,
                       "note": "independent_chains_events_per_s = the same GPUs running one full-size chain each (context, never the headline)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                         "traffic": None, "peak_source": peak_src + f" x {world} GPUs", "kernel": "zz_run_kernel_grid_multi",
                         "kernel_ms_per_launch": k_ms, "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": {"value": nacc * e_steps / e_dt, "unit": "events/s", "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": d2h_total, "steps": e_steps, "ms_per_step": 1e3 * e_dt / e_steps,
                    "what": "per rank: H2D of its slab + halo columns of (x0, theta0, c) from pinned memory, initialisation, event loop, D2H of the owned final state / counts / moment sums; host barriers between the phases included"},
            "parity_check": parity,
            "gpu_launches": 3 * args.steps * world, "clocks": clk.summary(),
        }))
    dist.barrier()
    run.close()
    prob.close()
    dist.destroy_process_group()


def config3_leg(zzb, R, T=10.0):
    """R independent chains of the logistic configuration (scripts/logistic.jl) as one block-diagonal problem, event loop only."""
    cfg = zzb.logistic_config()
    big = zzb.replicate_logistic(cfg, R) if R > 1 else cfg
    grad = zzb.LogisticSubsampled(big["A"], big["At"], big["y"], big["ny"], big["mu"], cfg["gamma0"], 10)
    Z = zzb.ZigZag(big["Gamma_drop"], big["mu"], big["sigma"], rho=0.5)
    prob = zzb.Problem(grad, Z)
    run = zzb.Run(prob, record_trace=False)
    try:
        run.upload(0.0, big["x0"], big["theta0"], big["c"], seed=(5, 6), adapt=True, factor=5.0)
        run.execute(T)          # warm-up (fills the L2 with the design)
        run.reset()
        run.execute(T)
        acc, num = run.counts()
        ms = run.device_ms
        return {"workload": f"sparse logistic regression n=8840 p=442 k=10 subsampled ZigZag (scripts/logistic.jl), {R} replicas side by side, "
                            f"T=[0,{T:g}], c=0.01, adapt, factor 5", "schedule": "sequential chains: one warp per chain (zz_seq_kernel_logit)",
                "replicas": R, "events": int(acc.sum()), "proposals": int(num), "kernel_ms": ms,
                "events_per_s": int(acc.sum()) / (ms * 1e-3), "proposals_per_s": int(num) / (ms * 1e-3)}
    finally:
        run.close()
        prob.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1000, help="grid side (d = n^2)")
    ap.add_argument("--T", type=float, default=2.0, help="simulated time per step")
    ap.add_argument("--cpu-T", type=float, default=0.0, help="simulated time of the CPU sample (default: T)")
    ap.add_argument("--frac", type=float, default=0.25, help="window length controller: proposals per window / d")
    ap.add_argument("--tight", action="store_true", help="c = sqrt(eps) (scripts/example.jl:39) instead of column norms")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tight", action="store_true", help="skip the extra c = sqrt(eps) measurement")
    ap.add_argument("--no-config3", action="store_true", help="skip the side measurement of config 3 (logistic replicas)")
    ap.add_argument("--config3-replicas", type=int, default=296)
    ap.add_argument("--no-trace", action="store_true", help="skip the e2e measurement with the full trace")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the bench workload")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
