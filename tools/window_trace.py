"""Development: per-CTA round log of ONE window of the asynchronous schedule (ZZB200_DBG_WINDOW / ZZB200_DBG_FILE).
Usage on the GPU box: python tools/window_trace.py run 12 gpurun_out/wtrace.bin ; here: python tools/window_trace.py show gpurun_out/wtrace.bin"""
import os, sys
import numpy as np

REC = 96
if sys.argv[1] == "run":
    os.environ["ZZB200_DBG_WINDOW"] = sys.argv[2]
    os.environ["ZZB200_DBG_FILE"] = sys.argv[3]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as g
    z = g.load_package(); z.init(0)
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
    G, x0, th0, c = z.gmrf_config(n)
    prob = z.Problem(z.GaussianPotential(G), z.ZigZag(G, np.zeros(G.n)))
    for rep in range(2):
        run = z.Run(prob, record_trace=False); run.upload(0.0, x0, th0, c, seed=(1, 2)); ms = run.execute(2.0); run.close()
    print("kernel ms", ms)
else:
    a = np.fromfile(sys.argv[2], dtype=np.uint64); a = a[:len(a) - 4096 * 16].reshape(-1, REC, 4)
    ncta = a.shape[0]
    # align the per-CTA time axis at the start record (kind 1) -- globaltimer is global already
    t0 = min(int(a[b, 0, 2]) for b in range(ncta) if a[b, 0, 0] == 1)
    names = {1: "begin", 2: "scan", 3: "round", 4: "poll", 5: "conv", 6: "bar1", 7: "commit", 8: "bar2"}
    summ = []
    for b in range(ncta):
        recs = [(int(k), int(c), (int(t) - t0) / 1e3, int(ck)) for k, c, t, ck in a[b] if k]
        rounds = [r for r in recs if r[0] == 3]
        conv = [r for r in recs if r[0] == 5]
        scan = [r for r in recs if r[0] == 2]
        polls = [r for r in recs if r[0] == 4]
        summ.append((b, scan[0][2] if scan else -1, scan[0][1] if scan else -1, len(rounds), sum(r[1] for r in rounds), conv[0][2] if conv else -1, len(polls), recs))
    conv_t = np.array([s[5] for s in summ]); print("convergence seen at (us): min %.1f median %.1f max %.1f" % (conv_t.min(), np.median(conv_t), conv_t.max()))
    print("scan done at (us): median %.1f max %.1f" % (np.median([s[1] for s in summ]), max(s[1] for s in summ)))
    print("rounds per CTA: median %d max %d; evals per CTA median %d max %d" % (np.median([s[3] for s in summ]), max(s[3] for s in summ), np.median([s[4] for s in summ]), max(s[4] for s in summ)))
    # last busy time per CTA = start of its last poll that ended in DONE
    last_round_start = np.array([max([r[2] for r in s[7] if r[0] == 3], default=0) for s in summ])
    print("start of last round (us): median %.1f p90 %.1f max %.1f" % (np.median(last_round_start), np.percentile(last_round_start, 90), last_round_start.max()))
    order = np.argsort(-last_round_start)
    for b in list(order[:4]) + [0]:
        print("CTA", b, " ".join("%s:%d@%.1f" % (names[r[0]], r[1], r[2]) for r in summ[b][7]))
