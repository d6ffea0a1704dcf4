"""GGUF -> device: a model loaded from a file equals the model built from the same synthetic tensors, and the generation
loop stops at EOS (ref: src/gemma_model.cpp:19-51, 159-194, 548-575)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from gguf_fixture import write_gemma_gguf

pytestmark = pytest.mark.gpu
DIMS = dict(n_embd=256, n_layer=2, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=96)


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_loaded_model_equals_synthetic_model(gpu, tmp_path, wtype):
    om = O.Model(wtype=wtype, sigma_embd=0.002, **DIMS)
    path = os.path.join(tmp_path, "tiny.gguf")
    toks = write_gemma_gguf(path, om, DIMS, wtype)
    m = gpu.GemmaModel.load_model_from_file(path)
    assert all(getattr(m.cfg, k) == v for k, v in DIMS.items())
    m.init_kv_cache()
    m.build_compute_graph(capture_acts=True)
    prompt = ((1 + 7919 * np.arange(12)) % DIMS["n_vocab"]).astype(np.int32)
    got, _, _ = m.generate(prompt, 6)
    want, _, _ = om.generate(prompt, 6, O.REF)
    assert np.array_equal(got, want)
    assert np.array_equal(m.logits().view(np.uint32), om.capture("result_output").view(np.uint32))
    # the reference's round: stops after max_tokens, prints prompt + new tokens
    ids, text, _, _ = m.begin_one_round_inference(prompt[:5], 4)
    assert np.array_equal(ids, om.generate(prompt[:5], 4, O.REF)[0])
    # like the reference, the loop prints a token when it feeds it back: the last one is returned but not printed
    assert text == "".join(toks[int(t)] for t in list(prompt[:5]) + list(ids[:3])).replace("▁", " ")
    m.close()
    om.close()


def test_generation_stops_at_eos(gpu, tmp_path):
    om = O.Model(wtype=O.TYPE_Q8_0, sigma_embd=0.002, **DIMS)
    prompt = ((1 + 7919 * np.arange(9)) % DIMS["n_vocab"]).astype(np.int32)
    stream, _, _ = om.generate(prompt, 8, O.REF)
    eos = int(stream[3])  # declare the 4th generated id the EOS token
    path = os.path.join(tmp_path, "tiny_eos.gguf")
    write_gemma_gguf(path, om, DIMS, O.TYPE_Q8_0, eos_id=eos)
    m = gpu.GemmaModel.load_model_from_file(path)
    m.init_kv_cache()
    m.build_compute_graph()
    ids, text, _, _ = m.begin_one_round_inference(prompt, 64)
    first = list(stream).index(eos)
    assert list(ids) == list(stream[:first + 1])  # ends with the EOS id, nothing after it
    m.close()
    om.close()


def test_k_quant_gguf_loads_into_k_quant_mode(gpu, tmp_path):
    """a q4_k_m-like file (Q4_K matrices, Q6_K token_embd): loaded, run operator by operator on the K-quant kernels,
    same tokens and logits as the oracle on the same blocks"""
    dims = dict(n_embd=256, n_layer=2, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=512, n_ctx=48)
    om = O.Model(wtype=O.TYPE_Q4_K, sigma_embd=0.005, **dims)
    path = os.path.join(tmp_path, "tiny_q4_k_m.gguf")
    write_gemma_gguf(path, om, dims, O.TYPE_Q4_K)
    m = gpu.GemmaModel.load_model_from_file(path)
    assert m.kquant_mode and all(getattr(m.cfg, k) == v for k, v in dims.items())
    m.init_kv_cache()
    m.build_compute_graph()
    prompt = ((1 + 7919 * np.arange(7)) % dims["n_vocab"]).astype(np.int32)
    got, _, _ = m.generate(prompt, 8)
    want, _, _ = om.generate(prompt, 8, O.REF)
    assert np.array_equal(got, want)
    assert np.array_equal(m.logits().view(np.uint32), om.capture("result_output").view(np.uint32))
    m.close()
    om.close()
