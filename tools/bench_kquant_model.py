"""Gemma-2B architecture with K-quant weights (q4_k_m-like: Q4_K matrices, Q6_K tied embedding), random blocks,
per-operator lowering captured in a CUDA graph: decode tokens/s at a pinned position."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemma_ggml_b200 as G

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
G.init(0)
dims = dict(G.GEMMA_2B, n_ctx=ctx)
m = G.GemmaModel(G.make_config(wtype=G.TYPE_Q4_K, **dims))
rng = np.random.default_rng(0)
E, F, V, L = dims["n_embd"], dims["n_ff"], dims["n_vocab"], dims["n_layer"]
QD, KD = dims["n_head"] * dims["head_dim"], dims["n_head_kv"] * dims["head_dim"]


def blocks(rows, cols, q6):
    nb, size = cols // 256, 210 if q6 else 144
    b = rng.integers(0, 256, size=(rows, nb, size), dtype=np.uint8)
    off = 208 if q6 else 0
    b[:, :, off] = 0; b[:, :, off + 1] = 0x10       # fp16 super-scale 2^-11
    if not q6:
        b[:, :, 2] = 0x80; b[:, :, 3] = 0x1b        # dmin ~ 7.5 d
    return np.ascontiguousarray(b.reshape(rows, nb * size))


nbytes = 0
t0 = time.time()
w = blocks(V, E, True); m.set_tensor_typed(0, G.TYPE_Q6_K, w); nbytes += w.nbytes
m.set_tensor_typed(1, G.TYPE_F32, np.ones(E, np.float32))
shapes = [None, (QD, E), (KD, E), (KD, E), (E, QD), None, (F, E), (F, E), (E, F)]
for il in range(L):
    for k, sh in enumerate(shapes):
        tid = 2 + 9 * il + k
        if sh is None:
            m.set_tensor_typed(tid, G.TYPE_F32, np.ones(E, np.float32))
        else:
            w = blocks(sh[0], sh[1], False); m.set_tensor_typed(tid, G.TYPE_Q4_K, w); nbytes += w.nbytes
print("weights %.2f GB uploaded in %.1f s" % (nbytes / 1e9, time.time() - t0))
m.init_kv_cache(); m.fill_kv_cache_synthetic(99)
m.build_compute_graph(False, False)
m.set_state(4242, ctx - 1, 0); m.run_steps(8); G.sync()
L_ = G.cuda_lib()
L_.ggb_prof_enable(1); G.prof_reset(); L_.ggb_prof_start(b"t")
m.run_steps(64)
L_.ggb_prof_stop(b"t"); G.sync()
ms, _ = G.prof_get("t"); L_.ggb_prof_enable(0)
us = ms * 1e3 / 64
kv = 2 * L * KD * 2 * ctx
print("K-quant Gemma-2B (Q4_K + Q6_K embd), n_kv %d, per-operator graph (%d kernels/step): %.1f us/token = %.0f tok/s; %.0f GB/s = %.3f of 6548"
      % (ctx, m.kernels_per_step, us, 1e6 / us, (nbytes + kv) / us / 1e3, (nbytes + kv) / us / 1e3 / 6548))
