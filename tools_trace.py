"""Phase-level trace of one fused decode step (where the time goes inside the persistent kernel)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemma_ggml_b200 as G

def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "2b"
    ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    G.init(0)
    dims = dict(getattr(G, {"2b": "GEMMA_2B", "7b": "GEMMA_7B"}[model]), n_ctx=ctx)
    m = G.GemmaModel(G.make_config(wtype=G.TYPE_Q4_0, **dims))
    m.init_synthetic_weights(); m.init_kv_cache(); m.fill_kv_cache_synthetic(99)
    m.build_compute_graph(False, True)
    m.set_state(4242, ctx - 1, 0); m.run_steps(3); G.sync()
    m.set_state(4242, ctx - 1, 0)
    st, kinds = m.trace_step()
    names = ["QKV", "ATTN", "COMBINE", "O", "GATE_UP", "DOWN", "LOGITS", "SAMPLE"]
    mhz = 1965.0
    tot = {}
    for pi, k in enumerate(kinds):
        s = st[pi].astype(np.float64)
        pro = (s[:, 1] - s[:, 0]) / mhz
        work = (s[:, 2] - s[:, 1]) / mhz
        bar = (s[:, 3] - s[:, 2]) / mhz
        whole = (s[:, 3] - s[:, 0]) / mhz
        d = tot.setdefault(names[k], dict(n=0, pro=0.0, work_max=0.0, work_mean=0.0, bar_min=0.0, whole=0.0))
        d["n"] += 1; d["pro"] += pro.max(); d["work_max"] += work.max(); d["work_mean"] += work.mean()
        d["bar_min"] += bar.min(); d["whole"] += np.median(whole)
    print("phase      n   prologue(max)  tiles(max)  tiles(mean)  barrier(min)  phase(median)   [us, summed over the token]")
    total = 0
    for k, d in tot.items():
        print(f"{k:9s} {d['n']:3d}  {d['pro']:10.1f}  {d['work_max']:10.1f}  {d['work_mean']:10.1f}  {d['bar_min']:10.1f}  {d['whole']:10.1f}")
        total += d["whole"]
    print("sum of phase medians: %.1f us" % total)
    ts = m.tile_stamps.astype(np.float64)
    for cta in (0, 73, 147):
        t = ts[cta]
        wait = (t[:, 1] - t[:, 0]) / mhz
        comp = (t[:, 2] - t[:, 1]) / mhz
        gap = (t[1:, 0] - t[:-1, 2]) / mhz
        print(f"LOGITS cta {cta}: wait us", np.round(wait[:12], 2), "compute us", np.round(comp[:12], 2), "gap", np.round(gap[:6], 2))
    ex = ts[:, 31, :]
    okk = ex[:, 0] > 0
    print('prologue experiment (2nd, 3rd run back-to-back) us: 2nd mean %.2f max %.2f | 3rd mean %.2f max %.2f' % (((ex[okk,1]-ex[okk,0])/mhz).mean(), ((ex[okk,1]-ex[okk,0])/mhz).max(), ((ex[okk,2]-ex[okk,1])/mhz).mean(), ((ex[okk,2]-ex[okk,1])/mhz).max()))
    # layer 1 detail
    for pi in range(6, 12):
        s = st[pi].astype(np.float64)
        print(names[kinds[pi]], "per-CTA us: pro max %.2f work max %.2f mean %.2f bar min %.2f max %.2f" % (
            ((s[:, 1]-s[:, 0])/mhz).max(), ((s[:, 2]-s[:, 1])/mhz).max(), ((s[:, 2]-s[:, 1])/mhz).mean(),
            ((s[:, 3]-s[:, 2])/mhz).min(), ((s[:, 3]-s[:, 2])/mhz).max()))

main()
