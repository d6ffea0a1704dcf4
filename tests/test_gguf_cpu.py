"""GGUF loader + tokenizer, host side only (no GPU): header KV / tensor table parse, hyper-parameters by the reference's
keys (ref: src/gemma_model.cpp:403-415), vocabulary, print_tokens text (ref: :749-793)."""
import os

import numpy as np
import pytest

import gemma_ggml_b200 as G
import oracle_lib as O
from gguf_fixture import write_gemma_gguf

DIMS = dict(n_embd=256, n_layer=2, n_head=4, n_head_kv=1, head_dim=64, n_ff=512, n_vocab=512, n_ctx=64)


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_probe_reads_header_tensors_and_tokenizer(tmp_path, wtype):
    om = O.Model(wtype=wtype, sigma_embd=0.002, **DIMS)
    path = os.path.join(tmp_path, "tiny.gguf")
    toks = write_gemma_gguf(path, om, DIMS, wtype)
    cfg, n_tensors, n_kv, n_tokens, bos, eos = G.gguf_probe(path)
    for k, v in DIMS.items():
        assert getattr(cfg, k) == v, k
    assert cfg.wtype == wtype and abs(cfg.eps - 1e-6) < 1e-12
    assert n_tensors == 2 + 9 * DIMS["n_layer"] and n_tokens == len(toks) and (bos, eos) == (2, 1) and n_kv >= 10
    # "<bos>" dropped, U+2581 -> space (ref: print_tokens)
    assert G.gguf_detokenize(path, [2, 4, 5, 6]) == " hello world!"
    om.close()


def test_probe_rejects_garbage(tmp_path):
    path = os.path.join(tmp_path, "bad.gguf")
    open(path, "wb").write(b"GGUF" + b"\x03\x00\x00\x00" + b"\xff" * 64)
    with pytest.raises(G.GgbError):
        G.gguf_probe(path)
    open(path, "wb").write(b"nope" * 16)
    with pytest.raises(G.GgbError):
        G.gguf_probe(path)


def test_probe_accepts_k_quant_files(tmp_path):
    """a q4_k_m-like file (Q4_K matrices, Q6_K token_embd): hyper-parameters read, wtype = the embedding's K-quant type"""
    dims = dict(DIMS, n_ff=1024)  # every row length a multiple of 256
    om = O.Model(wtype=O.TYPE_Q4_K, sigma_embd=0.005, **dims)
    path = os.path.join(tmp_path, "tiny_q4_k_m.gguf")
    write_gemma_gguf(path, om, dims, O.TYPE_Q4_K)
    cfg, n_tensors, _, _, _, _ = G.gguf_probe(path)
    assert cfg.wtype == O.TYPE_Q6_K and cfg.n_ff == 1024 and n_tensors == 2 + 9 * dims["n_layer"]
    om.close()


def test_oracle_k_quant_model_forward_is_finite_and_deterministic():
    dims = dict(DIMS, n_ff=1024, n_ctx=32)
    for wt in (O.TYPE_Q4_K, O.TYPE_Q6_K):
        m = O.Model(wtype=wt, sigma_embd=0.005, **dims)
        prompt = ((1 + 7919 * np.arange(5)) % dims["n_vocab"]).astype(np.int32)
        a, _, _ = m.generate(prompt, 4, O.REF)
        la = m.capture("result_output").copy()
        b, _, _ = m.generate(prompt, 4, O.REF)
        assert np.array_equal(a, b) and np.array_equal(la.view(np.uint32), m.capture("result_output").view(np.uint32))
        assert np.isfinite(la).all() and np.abs(la).max() > 0
        # the tied embedding of a Q4_K model is Q6_K (the q4_k_m mix)
        raw, rows, cols = m.tensor(0)
        assert raw.nbytes == rows * (cols // 256) * 210
        raw, rows, cols = m.tensor(3)
        assert raw.nbytes == rows * (cols // 256) * (144 if wt == O.TYPE_Q4_K else 210)
        m.close()
