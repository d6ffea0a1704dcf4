"""Run a few fused decode steps (Gemma-2B Q4_0, ctx 2048, pinned position) — the ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemma_ggml_b200 as G
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
G.init(0)
m = G.GemmaModel(G.make_config(wtype=G.TYPE_Q4_0, **dict(G.GEMMA_2B, n_ctx=2048)))
m.init_synthetic_weights(); m.init_kv_cache(); m.fill_kv_cache_synthetic(99)
m.build_compute_graph(False, True)
m.set_state(4242, 2047, 0)
m.run_steps(steps)
G.sync()
print("ok", m.get_state())
