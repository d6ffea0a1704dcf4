"""GPU parity tests, model level: the host graph builder (host/gemma_model.cpp) driving the C-ABI layer
versus the CPU oracle's restatement of the reference forward pass, on identical synthetic weights."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


SMALL = {
    "mqa": dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=96),
    # Gemma-7B-like: MHA, n_head*head_dim != n_embd, K > 4096 in the down projection (2 K-segments)
    "mha": dict(n_embd=512, n_layer=2, n_head=4, n_head_kv=4, head_dim=128, n_ff=4608, n_vocab=1024, n_ctx=64),
    # K = 3072 as in Gemma-7B: 96 blocks = one full 64-block segment + a short one (quad passes, kf = 1); n_ff = 1280:
    # 40 blocks, a short segment too long for a quad pass (zero-padded plain pass).  The configs above cover kf = 0
    # (K <= 1024) and kf = 2 (n_ff 4608)
    "k96": dict(n_embd=3072, n_layer=1, n_head=2, n_head_kv=2, head_dim=128, n_ff=1280, n_vocab=512, n_ctx=48),
    # every K a multiple of 2048 (64-block segments only), as in Gemma-2B: the plain (non-quad) instantiation of the
    # persistent kernel, which the configs above never reach (their short K rows all take the quad layout)
    "plain": dict(n_embd=2048, n_layer=1, n_head=8, n_head_kv=1, head_dim=256, n_ff=4096, n_vocab=512, n_ctx=48),
    # outside the fused kernel's envelope (n_embd % 256 != 0): exercises the per-operator lowering
    "odd": dict(n_embd=384, n_layer=2, n_head=4, n_head_kv=2, head_dim=96, n_ff=1056, n_vocab=777, n_ctx=48),
}


def _models(gpu, dims, wtype, capture=True, sigma_embd=0.002, fused=True):
    cfg = gpu.make_config(wtype=wtype, sigma_embd=sigma_embd, **dims)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights()
    m.init_kv_cache()
    m.build_compute_graph(capture_acts=capture, fused=fused)
    om = O.Model(wtype=wtype, sigma_embd=sigma_embd, **dims)
    return m, om


@pytest.mark.parametrize("name", ["mha", "k96", "mqa"])
@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_two_attention_groups_match_oracle(gpu, monkeypatch, name, wtype):
    """The long-context MHA form of the persistent kernel (two 8-warp groups, each on its own ATTN item, 4-slot
    weight ring; chosen automatically for MHA models with n_ctx >= 4096) forced onto the small configs: bit-exact."""
    monkeypatch.setenv("GGB_ATTN_GROUPS", "2")
    dims = SMALL[name]
    m, om = _models(gpu, dims, wtype, fused=True)
    assert m.is_fused
    prompt = ((1 + 7919 * np.arange(35)) % dims["n_vocab"]).astype(np.int32)  # crosses a 32-position chunk boundary
    toks, _, _ = m.generate(prompt, 6)
    dtoks, _, _ = om.generate(prompt, 6, O.DEV)
    assert np.array_equal(toks, dtoks)
    assert np.array_equal(bits(m.logits()), bits(om.capture("result_output")))
    for il in range(dims["n_layer"]):
        for cap in ("kqv_out", "l_out"):
            assert np.array_equal(bits(m.capture(cap, il)), bits(om.capture(cap, il))), (cap, il)
    m.close()
    om.close()


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "ops"])
@pytest.mark.parametrize("name", ["mqa", "mha", "k96", "plain", "odd"])
@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_generate_matches_oracle(gpu, name, wtype, fused):
    dims = SMALL[name]
    m, om = _models(gpu, dims, wtype, fused=fused)
    assert m.is_fused == (fused and name != "odd")
    if m.is_fused:  # the computed soft-max exp must have been proved identical to the fp16 table on this GPU
        assert m.exp_computed
    prompt = ((1 + 7919 * np.arange(12)) % dims["n_vocab"]).astype(np.int32)
    n_decode = 10
    toks, _, _ = m.generate(prompt, n_decode)
    dtoks, _, _ = om.generate(prompt, n_decode, O.DEV)
    assert np.array_equal(toks, dtoks)
    # every named intermediate of the last step is BIT-IDENTICAL to the oracle in DEV order
    for il in range(dims["n_layer"]):
        for cap in ("kqv_out", "sa_out", "ffn_act", "l_out"):
            assert np.array_equal(bits(m.capture(cap, il)), bits(om.capture(cap, il))), (cap, il)
    assert np.array_equal(bits(m.capture("result_norm")), bits(om.capture("result_norm")))
    glog = m.logits()
    assert np.array_equal(bits(glog), bits(om.capture("result_output")))
    # KV cache contents (fp16 bits) identical
    for il in range(dims["n_layer"]):
        gk, gv = m.kv(il)
        ok, ov = om.kv(il)
        n = len(prompt) + n_decode - 1
        assert np.array_equal(gk[:n], ok[:n]) and np.array_equal(gv[:n], ov[:n])
    # versus the REF (scalar ggml) order: greedy tokens identical; logits / per-layer outputs within
    # REF_TOL of the tensor's max magnitude.  The tolerance is set by Q8_0 re-quantisation of the
    # activations, not by fp32 rounding: a 1-ulp difference in an activation can move one int8 code by
    # one step (amax/127 = 0.8 % of the block's range), and that step then propagates through the
    # following mat-vecs.  REF_TOL = 2e-2 of max|tensor|.
    REF_TOL = 2e-2
    rtoks, _, _ = om.generate(prompt, n_decode, O.REF)
    assert np.array_equal(toks, rtoks)
    rlog = om.capture("result_output")
    assert np.abs(glog - rlog).max() < REF_TOL * max(1.0, float(np.abs(rlog).max()))
    for il in range(dims["n_layer"]):
        r = om.capture("l_out", il)
        g = m.capture("l_out", il)
        assert np.abs(g - r).max() < REF_TOL * max(1.0, float(np.abs(r).max())), il
    m.close()
    om.close()


def test_uploaded_weights_equal_synthetic_weights(gpu):
    """set_tensor(host AoS blocks from the oracle) gives the same model as on-device generation."""
    dims = SMALL["mqa"]
    m, om = _models(gpu, dims, O.TYPE_Q4_0)
    prompt = np.array([5, 9, 200, 31], dtype=np.int32)
    t1, _, _ = m.generate(prompt, 4)
    l1 = m.logits()
    n_tensors = 2 + 9 * dims["n_layer"]
    for tid in range(n_tensors):
        raw, rows, cols = om.tensor(tid)
        m.set_tensor(tid, raw)
    m.build_compute_graph(capture_acts=True)
    t2, _, _ = m.generate(prompt, 4)
    assert np.array_equal(t1, t2) and np.array_equal(bits(l1), bits(m.logits()))
    m.close()
    om.close()


def test_device_resident_loop_equals_host_stepped_loop(gpu):
    """run_steps (no host round trip, token fed back on the device) == decode_step per token."""
    dims = SMALL["mqa"]
    m, om = _models(gpu, dims, O.TYPE_Q8_0, capture=False)
    prompt = np.array([7, 8, 9], dtype=np.int32)
    toks, _, _ = m.generate(prompt, 9)
    m.init_kv_cache()
    for t, tok in enumerate(prompt[:-1]):
        m.decode_step(tok, t)
    m.set_state(int(prompt[-1]), len(prompt) - 1, 1)
    m.run_steps(9)
    gpu.sync()
    tok, pos = m.get_state()
    assert tok == toks[-1] and pos == len(prompt) - 1 + 9
    m.close()
    om.close()


def test_pinned_position_is_idempotent(gpu):
    """pos_step = 0 (fixed n_kv, used by bench.py): repeating a step reproduces the same logits."""
    dims = SMALL["mqa"]
    m, om = _models(gpu, dims, O.TYPE_Q4_0, capture=False)
    m.fill_kv_cache_synthetic(3)
    m.set_state(11, dims["n_ctx"] - 1, 0)
    m.run_steps(1)
    gpu.sync()
    l1, t1 = m.logits(), m.get_state()
    m.set_state(11, dims["n_ctx"] - 1, 0)
    m.run_steps(1)
    gpu.sync()
    assert np.array_equal(bits(l1), bits(m.logits())) and t1 == m.get_state()
    m.close()
    om.close()


def test_full_size_gemma_2b_q4_0_properties(gpu):
    """BASELINE config 2 at full size (Gemma-2B, Q4_0, ctx 2048).  The oracle needs minutes there, so
    parity is carried by size-independent properties: (1) every resident weight equals the oracle's
    generator on sampled rows; (2) a step is deterministic / idempotent at pinned n_kv; (3) the logits
    of one step equal the oracle's DEV-order output projection of the GPU's own final hidden state;
    (4) the sampled token is the first maximum of those logits."""
    dims = dict(gpu.GEMMA_2B, n_ctx=2048)
    cfg = gpu.make_config(wtype=O.TYPE_Q4_0, **dims)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights()
    m.init_kv_cache()
    m.fill_kv_cache_synthetic(99)
    m.build_compute_graph(capture_acts=False)
    assert m.weight_bytes == 1409679360  # SURVEY.md §8d algorithmic bytes for 2B Q4_0
    m.set_state(4242, 2047, 0)
    m.run_steps(1)
    gpu.sync()
    logits = m.logits()
    tok, pos = m.get_state()
    assert pos == 2047 and tok == O.argmax(logits)
    m.set_state(4242, 2047, 0)
    m.run_steps(3)
    gpu.sync()
    # after the first repeat the fed-back token differs, so compare a fresh identical step instead
    m.set_state(4242, 2047, 0)
    m.run_steps(1)
    gpu.sync()
    assert np.array_equal(bits(logits), bits(m.logits()))
    # (3) output projection rows vs the oracle on the GPU's own normalised hidden state
    h = m.capture("result_norm")
    a = O.quantize(O.TYPE_Q8_0, h[None, :])
    rows = np.r_[0:64, 100000:100032, 255968:256000]
    w = np.stack([O.quantize(O.TYPE_Q4_0, O.synth_row(cfg.seed, 0, int(r), 2048, 0.02)[None, :])[0] for r in rows])
    expect = O.mul_mat(O.TYPE_Q4_0, O.DEV, w, a, 2048)[0]
    assert np.array_equal(bits(logits[rows]), bits(expect))
    m.close()
