"""ops lowering vs oracle REF: bitwise"""
import sys, os, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import gemma_ggml_b200 as G
import oracle_lib as O
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
G.init(0)
def bits(a): return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
SMALL = {
    "mqa": dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=96),
    "mha": dict(n_embd=512, n_layer=2, n_head=4, n_head_kv=4, head_dim=128, n_ff=4608, n_vocab=1024, n_ctx=64),
    "k96": dict(n_embd=3072, n_layer=1, n_head=2, n_head_kv=2, head_dim=128, n_ff=1280, n_vocab=512, n_ctx=48),
    "plain": dict(n_embd=2048, n_layer=1, n_head=8, n_head_kv=1, head_dim=256, n_ff=4096, n_vocab=512, n_ctx=48),
    "odd": dict(n_embd=384, n_layer=2, n_head=4, n_head_kv=2, head_dim=96, n_ff=1056, n_vocab=777, n_ctx=48),
}
bad = 0
for name, dims in SMALL.items():
    for wt in (O.TYPE_Q4_0, O.TYPE_Q8_0):
        m = G.GemmaModel(G.make_config(wtype=wt, sigma_embd=0.002, **dims)); m.init_synthetic_weights(); m.init_kv_cache()
        m.build_compute_graph(capture_acts=True, fused=fused)
        om = O.Model(wtype=wt, sigma_embd=0.002, **dims)
        prompt = ((1 + 7919 * np.arange(35)) % dims["n_vocab"]).astype(np.int32)
        toks, _, _ = m.generate(prompt, 8)
        rt, _, _ = om.generate(prompt, 8, O.REF)
        ok = np.array_equal(toks, rt)
        res = [("tok", ok)]
        for cap in ("kqv_out", "sa_out", "ffn_act", "l_out"):
            for il in range(dims["n_layer"]):
                res.append((f"{cap}{il}", np.array_equal(bits(m.capture(cap, il)), bits(om.capture(cap, il)))))
        res.append(("logits", np.array_equal(bits(m.logits()), bits(om.capture("result_output")))))
        fails = [k for k, v in res if not v]
        bad += len(fails)
        print(name, wt, "fused" if m.is_fused else "ops", "OK" if not fails else "FAIL " + ",".join(fails), flush=True)
        m.close(); om.close()
print("small configs: mismatching tensors", bad)
if len(sys.argv) > 2:
    for w, wt in (("q8_0", O.TYPE_Q8_0), ("q4_0", O.TYPE_Q4_0)):
        z = np.load(f"/root/repo/tests/golden/config1_2b_{w}.npz")
        dims = dict(G.GEMMA_2B, n_ctx=int(z["n_ctx"]))
        m = G.GemmaModel(G.make_config(wtype=wt, sigma_embd=float(z["sigma_embd"]), **dims)); m.init_synthetic_weights(); m.init_kv_cache()
        m.build_compute_graph(capture_acts=False, fused=fused)
        prompt = z["prompt"]
        t0 = time.time()
        for t, tok in enumerate(prompt[:-1]): m.decode_step(int(tok), t)
        tok = int(prompt[-1]); pos = len(prompt) - 1
        nmatch = 0; maxdev = 0.0; bitsame = 0
        for s in range(len(z["tokens"])):
            tok = m.decode_step(tok, pos); pos += 1
            lg = m.logits()
            dev = np.abs(lg[z["sample_idx"]] - z["sample_val"][s]).max()
            maxdev = max(maxdev, dev)
            bitsame += int(np.array_equal(bits(lg[z["top_idx"][s]]), bits(z["top_val"][s])) and np.array_equal(bits(lg[z["sample_idx"]]), bits(z["sample_val"][s])))
            if tok == int(z["tokens"][s]): nmatch += 1
            else:
                print("  first mismatch at step", s, tok, int(z["tokens"][s])); break
        print(w, "fused" if m.is_fused else "ops", "tokens matched %d/64, steps with bit-identical sampled logits %d, max dev %.3e, %.1f s" % (nmatch, bitsame, maxdev, time.time() - t0), flush=True)
        m.close()
