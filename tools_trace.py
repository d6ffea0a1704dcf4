"""Phase-level trace of one fused decode step (where the time goes inside the persistent kernel).

Stamps are %globaltimer ns (one clock for all SMs): per phase and CTA start / inputs-arrived / end.
A phase's share of the token's critical path = (latest end over CTAs) - (latest end of the previous phase)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemma_ggml_b200 as G

NAMES = ["QKV", "SCORE", "PV", "O", "GATE_UP", "DOWN", "LOGITS", "SAMPLE"]


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "2b"
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    wtype = G.TYPE_Q8_0 if (len(sys.argv) > 3 and sys.argv[3] == "q8") else G.TYPE_Q4_0
    G.init(0)
    dims = dict(getattr(G, {"2b": "GEMMA_2B", "7b": "GEMMA_7B"}[model]), n_ctx=ctx)
    m = G.GemmaModel(G.make_config(wtype=wtype, **dims))
    m.init_synthetic_weights(); m.init_kv_cache(); m.fill_kv_cache_synthetic(99)
    m.build_compute_graph(False, True)
    m.set_state(4242, ctx - 1, 0); m.run_steps(3); G.sync()
    m.set_state(4242, ctx - 1, 0)
    st, kinds = m.trace_step()
    st = st.astype(np.float64) / 1e3  # us
    t_begin = st[0, :, 0].min()
    prev_end = t_begin
    tot = {}
    rows = []
    for pi, k in enumerate(kinds):
        s = st[pi]
        busy = s[:, 2] > s[:, 0]
        end = s[:, 2].max()
        crit = end - prev_end
        wait = (s[:, 1] - s[:, 0])          # prologue incl. waiting for inputs
        work = (s[:, 2] - s[:, 1])
        d = tot.setdefault(NAMES[k], dict(n=0, crit=0.0, wait_min=0.0, wait_max=0.0, work_max=0.0, work_mean=0.0))
        d["n"] += 1; d["crit"] += crit; d["wait_min"] += wait[busy].min() if busy.any() else 0.0
        d["wait_max"] += wait.max(); d["work_max"] += work.max(); d["work_mean"] += work[busy].mean() if busy.any() else 0.0
        rows.append((pi, NAMES[k], crit, wait[busy].min() if busy.any() else 0.0, wait.max(), work.max(), int(busy.sum())))
        prev_end = end
    print("phase      n   critical-path  prologue(min)  prologue(max)  work(max)  work(mean)   [us, summed over the token]")
    for k, d in tot.items():
        print(f"{k:9s} {d['n']:3d}  {d['crit']:12.1f}  {d['wait_min']:12.1f}  {d['wait_max']:12.1f}  {d['work_max']:9.1f}  {d['work_mean']:9.1f}")
    print("token (first start -> last end): %.1f us" % (prev_end - t_begin))
    print("layer 5 detail: phase, critical us, prologue min/max us, work max us, busy CTAs")
    for r in rows[30:36]:
        print("  %-8s crit %6.2f  prologue %5.2f / %5.2f  work %5.2f  ctas %d" % r[1:])
    if os.environ.get("GGB_PROF_PHASE"):
        pr = m.prof.astype(np.float64) / 1965.0  # us at 1965 MHz
        ph = int(os.environ["GGB_PROF_PHASE"]); kind = NAMES[kinds[ph]]
        np.set_printoptions(linewidth=200, suppress=True)
        if kind in ("SCORE", "PV"):
            ok = pr[:, 0] > 0
            rel = pr[ok, :8] - pr[ok, 0:1]
            print("inside phase %d (%s): stamps in us since item start, median / max over CTAs" % (ph, kind))
            print("   median", np.round(np.median(rel, axis=0), 2)); print("   max   ", np.round(rel.max(axis=0), 2))
        else:
            print("inside phase %d (%s), us per CTA (median / max): consumer warp 0: prologue, wait weights, wait term buffer, compute, total, passes" % (ph, kind))
            print("   median", np.round(np.median(pr[:, :6], axis=0), 2)); print("   max   ", np.round(pr[:, :6].max(axis=0), 2))
            print("   chain warp: wait terms, chain, finish (incl. GeGLU), rounds, total, passes, of chain: adds, row finish")
            ch = pr[:, 16:24].copy(); ch[:, 3] *= 1965.0; ch[:, 5] *= 1965.0  # rounds / passes are counts
            print("   median", np.round(np.median(ch, axis=0), 2)); print("   max   ", np.round(ch.max(axis=0), 2))
            print("   consumer warp 0, slots 0..3: (weights arrived, term buffer free, terms stored) us since phase start:", np.round(np.median(pr[:, 32:44], axis=0), 2))
            print("   chain warp 0, its slots: (terms arrived, slot done) us since phase start:", np.round(np.median(pr[:, 48:64], axis=0), 2))
            print("   row-lane mode: end of chain rounds 1..8, us since phase start (median)", np.round(np.median(pr[:, 24:32], axis=0), 2))
    print("layer 5, per phase: CTA start / inputs-arrived / end relative to the previous phase's last end (min, median, max us)")
    prev = st[29][:, 2].max()
    for pi in range(30, 36):
        s_ = st[pi] - prev
        f = lambda a: "%5.2f %5.2f %5.2f" % (a.min(), np.median(a), a.max())
        print("  %-8s start %s | arrived %s | end %s" % (NAMES[kinds[pi]], f(s_[:, 0]), f(s_[:, 1]), f(s_[:, 2])))
        prev = st[pi][:, 2].max()
    print("last phases:")
    for r in rows[-2:]:
        print("  %-8s crit %6.2f  prologue %5.2f / %5.2f  work %5.2f  ctas %d" % r[1:])


main()
