"""Generates tests/golden/config1_<model>_<wtype>.npz — BASELINE.json config 1 at FULL size on the CPU oracle.

Gemma-2B architecture, seeded random-init weights, prompt id_i = (1 + 7919 i) mod V (128 tokens), 64 greedy
tokens, through the oracle's REF order (the scalar ggml order of the reference's CPU path; loop of
ref: src/gemma_model.cpp:548-575: prefill, then one token per forward).  Stored per decode step:

  tokens[64]            the greedy ids (the GPU path must reproduce them exactly)
  top_idx/top_val[64,16] the 16 largest REF logits (top-2 margin = top_val[:,0] - top_val[:,1])
  sample_idx[1024], sample_val[64,1024]   a strided sample of the REF logits (max-abs deviation of the GPU logits)
  absmax[64]            max |logit| of the step

The oracle needs ~1-2 s per token on 8-16 cores, so this runs in minutes; the GPU box only reads the fixture.

Usage: python tests/golden/make_config1.py [2b] [q8_0|q4_0] [n_prompt] [n_decode] [sigma_embd]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

GEMMA = {
    "2b": dict(n_embd=2048, n_layer=18, n_head=8, n_head_kv=1, head_dim=256, n_ff=16384, n_vocab=256000),
    "7b": dict(n_embd=3072, n_layer=28, n_head=16, n_head_kv=16, head_dim=256, n_ff=24576, n_vocab=256000),
}
WT = {"q4_0": O.TYPE_Q4_0, "q8_0": O.TYPE_Q8_0}


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "2b"
    wname = sys.argv[2] if len(sys.argv) > 2 else "q8_0"
    n_prompt = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    n_decode = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    sigma_embd = float(sys.argv[5]) if len(sys.argv) > 5 else 0.004
    n_ctx = 256
    dims = dict(GEMMA[model], n_ctx=n_ctx)
    V = dims["n_vocab"]
    prompt = ((1 + 7919 * np.arange(n_prompt)) % V).astype(np.int32)
    t0 = time.time()
    m = O.Model(wtype=WT[wname], sigma_embd=sigma_embd, **dims)
    print("weights %.1f s" % (time.time() - t0), flush=True)
    sample_idx = np.arange(0, V, V // 1024, dtype=np.int64)[:1024]
    tokens = np.zeros(n_decode, np.int32)
    top_idx = np.zeros((n_decode, 16), np.int32)
    top_val = np.zeros((n_decode, 16), np.float32)
    sample_val = np.zeros((n_decode, sample_idx.size), np.float32)
    absmax = np.zeros(n_decode, np.float32)
    t0 = time.time()
    logits = m.forward(prompt, 0, O.REF)
    print("prefill %.1f s" % (time.time() - t0), flush=True)
    for s in range(n_decode):
        tok = O.argmax(logits)
        tokens[s] = tok
        order = np.argsort(-logits, kind="stable")[:16]
        top_idx[s], top_val[s] = order, logits[order]
        assert order[0] == tok or logits[order[0]] == logits[tok]
        sample_val[s] = logits[sample_idx]
        absmax[s] = np.abs(logits).max()
        print(s, tok, "margin %.4f absmax %.3f  %.0f s" % (top_val[s, 0] - top_val[s, 1], absmax[s], time.time() - t0), flush=True)
        if s + 1 < n_decode:
            logits = m.forward(np.array([tok], np.int32), n_prompt + s, O.REF)
    out = os.path.join(HERE, "config1_%s_%s.npz" % (model, wname))
    np.savez_compressed(out, tokens=tokens, top_idx=top_idx, top_val=top_val, sample_idx=sample_idx,
                        sample_val=sample_val.astype(np.float32), absmax=absmax, prompt=prompt,
                        sigma_embd=np.float32(sigma_embd), n_ctx=np.int32(n_ctx))
    print("distinct tokens:", len(set(tokens.tolist())), "->", out)
    m.close()


if __name__ == "__main__":
    main()
