"""GPU parity tests for the prefill path (mul_mat with col_num = T > 1): the dequant -> bf16 tcgen05 GEMM
(csrc/prefill.cu) against a numpy model of the same contraction, and the prefill forward pass against the
CPU oracle's REF-order forward pass within a stated tolerance (prefill rounds activations to bf16 instead
of re-quantising them to Q8_0, so it is tolerance-exact, not bit-exact: DESIGN.md section 3)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def to_bf16(x):
    """round-to-nearest-even fp32 -> bf16, returned as fp32 values and as uint16 bit patterns"""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32)
    return (r << 16).astype(np.uint32).view(np.float32), r.astype(np.uint16)


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
@pytest.mark.parametrize("T,N,K", [(512, 512, 2048), (100, 96, 256), (512, 256, 16384), (37, 2048, 2048), (300, 1024, 512)])
def test_gemm_bf16_matches_dequantised_model(gpu, wtype, T, N, K):
    rng = np.random.default_rng(T + N + K)
    L = gpu.cuda_lib()
    wq = O.quantize(wtype, rng.standard_normal((N, K)).astype(np.float32) * 0.02)
    W = gpu.Weight.upload(wtype, wq, K)
    w_f32 = O.dequantize(wtype, wq, K)                      # exact code * scale
    w_bf, _ = to_bf16(w_f32)
    a_f32 = rng.standard_normal((T, K)).astype(np.float32)
    a_bf, a_bits = to_bf16(a_f32)
    resid = rng.standard_normal((T, N)).astype(np.float32)
    A = gpu.DeviceBuffer.from_numpy(a_bits)
    R = gpu.DeviceBuffer.from_numpy(resid)
    Cb = gpu.DeviceBuffer(4 * T * N)
    ws_bytes = L.ggb_gemm_workspace_bytes(T, N, K)
    WS = gpu.DeviceBuffer(ws_bytes)
    gpu.check(L.ggb_op_gemm_bf16(W.h, A.ptr, T, Cb.ptr, N, R.ptr, WS.ptr, ws_bytes))
    gpu.sync()
    got = Cb.download(np.float32).reshape(T, N)
    ref = a_bf.astype(np.float64) @ w_bf.astype(np.float64).T + resid
    # fp32 accumulation of exact bf16 x bf16 products: the error is the accumulation rounding only
    scale = np.abs(a_bf).astype(np.float64) @ np.abs(w_bf).astype(np.float64).T + np.abs(resid)
    err = np.abs(got - ref)
    assert np.all(err <= 1e-5 * scale + 1e-6), (float(err.max()), float((err / (scale + 1e-30)).max()))
    # and without residual / workspace
    gpu.check(L.ggb_op_gemm_bf16(W.h, A.ptr, T, Cb.ptr, N, None, None, 0))
    gpu.sync()
    got2 = Cb.download(np.float32).reshape(T, N)
    assert np.all(np.abs(got2 - (ref - resid)) <= 1e-5 * scale + 1e-6)


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
@pytest.mark.parametrize("T,rows,K", [(512, (2048, 256, 256), 2048), (70, (64, 32, 96), 256), (300, (1024, 1024), 512)])
def test_gemm_multi_equals_separate_launches(gpu, wtype, T, rows, K):
    """Q | K | V (gate | up) in one launch: bit-identical to one launch per matrix (same tiles, same K-split rule
    is NOT guaranteed - more tiles per launch can mean fewer splits - so the comparison is against the same model,
    and bitwise against the single launches whenever neither form splits K)."""
    import ctypes as C
    rng = np.random.default_rng(T + K + sum(rows))
    L = gpu.cuda_lib()
    a_bf, a_bits = to_bf16(rng.standard_normal((T, K)).astype(np.float32))
    A = gpu.DeviceBuffer.from_numpy(a_bits)
    ws, outs, refs, scales = [], [], [], []
    for n in rows:
        wq = O.quantize(wtype, rng.standard_normal((n, K)).astype(np.float32) * 0.02)
        ws.append(gpu.Weight.upload(wtype, wq, K))
        w_bf, _ = to_bf16(O.dequantize(wtype, wq, K))
        refs.append(a_bf.astype(np.float64) @ w_bf.astype(np.float64).T)
        scales.append(np.abs(a_bf).astype(np.float64) @ np.abs(w_bf).astype(np.float64).T)
        outs.append(gpu.DeviceBuffer(4 * T * n))
    n_total = sum((n + 255) // 256 * 256 for n in rows)
    ws_bytes = L.ggb_gemm_workspace_bytes(T, n_total, K)
    WS = gpu.DeviceBuffer(ws_bytes)
    wp = (C.c_void_p * len(rows))(*[w.h for w in ws])
    cp = (C.c_void_p * len(rows))(*[o.ptr for o in outs])
    ld = (C.c_int64 * len(rows))(*rows)
    for use_ws in (True, False):
        gpu.check(L.ggb_op_gemm_bf16_multi(len(rows), wp, A.ptr, T, cp, ld, WS.ptr if use_ws else None, ws_bytes if use_ws else 0))
        gpu.sync()
        for i, n in enumerate(rows):
            got = outs[i].download(np.float32).reshape(T, n)
            assert np.all(np.abs(got - refs[i]) <= 1e-5 * scales[i] + 1e-6), (use_ws, i)
    # without a workspace nothing splits K: bitwise equal to the single launches
    single = gpu.DeviceBuffer(4 * T * max(rows))
    for i, n in enumerate(rows):
        gpu.check(L.ggb_op_gemm_bf16(ws[i].h, A.ptr, T, single.ptr, n, None, None, 0))
        gpu.sync()
        a = single.download(np.float32)[:T * n]
        b = outs[i].download(np.float32)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), i


SMALL = {
    "mqa": dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=160),
    "mha": dict(n_embd=512, n_layer=2, n_head=4, n_head_kv=4, head_dim=128, n_ff=4608, n_vocab=1024, n_ctx=96),
}
# Prefill rounds activations to bf16 (8 significant bits) where the reference re-quantises them to Q8_0 per
# block of 32 (7 bits + sign relative to the block maximum), and weights q*d are rounded once to bf16.  Both
# are ~0.4 % relative perturbations per mat-mul input; PREFILL_TOL bounds the resulting difference of the
# logits / layer outputs relative to the tensor's maximum magnitude.
PREFILL_TOL = 3e-2


@pytest.mark.parametrize("name,T", [("mqa", 128), ("mqa", 37), ("mha", 64)])
@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_prefill_matches_oracle_ref_forward(gpu, name, wtype, T):
    dims = SMALL[name]
    cfg = gpu.make_config(wtype=wtype, sigma_embd=0.002, **dims)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=False, fused=True)
    om = O.Model(wtype=wtype, sigma_embd=0.002, **dims)
    prompt = ((1 + 7919 * np.arange(T)) % dims["n_vocab"]).astype(np.int32)
    nxt = m.prefill(prompt, 0)
    rlog = om.forward(prompt, 0, O.REF)
    glog = m.prefill_logits(T - 1)
    tol = PREFILL_TOL * max(1.0, float(np.abs(rlog).max()))
    assert np.abs(glog - rlog).max() < tol, float(np.abs(glog - rlog).max())
    top = np.sort(rlog)[-2:]
    if top[1] - top[0] > 2 * tol:  # the greedy token is only pinned when the reference's margin exceeds the tolerance
        assert nxt == O.argmax(rlog)
    # KV cache written by the prefill agrees with the reference's (fp16 values)
    for il in range(dims["n_layer"]):
        gk, gv = m.kv(il)
        ok, ov = om.kv(il)
        for g_, o_ in ((gk, ok), (gv, ov)):
            gf, of = g_[:T].view(np.float16).astype(np.float32), o_[:T].view(np.float16).astype(np.float32)
            assert np.abs(gf - of).max() < PREFILL_TOL * max(1.0, float(np.abs(of).max())), il
    # decode continues from the prefilled cache: the next decode step runs and yields a valid token
    tok2 = m.decode_step(nxt, T)
    assert 0 <= tok2 < dims["n_vocab"]
    m.close(); om.close()


def test_prefill_agrees_with_token_by_token_decode_path(gpu):
    """Same prompt through the bit-exact decode kernel (one token at a time) and through the GEMM prefill:
    last-position logits agree within the tolerance, at a size the oracle would need minutes for."""
    dims = dict(n_embd=1024, n_layer=4, n_head=8, n_head_kv=1, head_dim=128, n_ff=4096, n_vocab=8192, n_ctx=320)
    T = 256
    cfg = gpu.make_config(wtype=O.TYPE_Q8_0, sigma_embd=0.002, **dims)
    prompt = ((1 + 7919 * np.arange(T)) % dims["n_vocab"]).astype(np.int32)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=False, fused=True)
    toks, _, _ = m.generate(prompt, 1)
    dlog = m.logits()
    m.init_kv_cache()
    nxt = m.prefill(prompt, 0)
    glog = m.prefill_logits(T - 1)
    tol = PREFILL_TOL * max(1.0, float(np.abs(dlog).max()))
    assert np.abs(glog - dlog).max() < tol
    top = np.sort(dlog)[-2:]
    if top[1] - top[0] > 2 * tol:
        assert nxt == int(toks[0])
    m.close()


# Full size (BASELINE.json config 3: Gemma-2B Q8_0, 512-token prompt).  The reference for the comparison is our own decode
# path run token by token, which is bit-identical to the oracle's scalar ggml order (tests/test_model_gpu.py) - the oracle
# itself would need ~15 minutes for 512 positions.  The two paths round differently BY DESIGN (bf16 activations and
# weights vs Q8_0 re-quantised activations), and 18 layers of random-init weights amplify that: the full-size tolerance
# is therefore looser than PREFILL_TOL and is stated here from the measurement (profiles/r2_config3_prefill_parity.json:
# what was measured, not just the bound).
FULL_PREFILL_TOL = 1.2e-1  # measured: logits 5.3 % of their range, K / V rows 0.4 % (layer 0) .. 7.4 % (layer 17); same greedy token


def test_config3_full_size_prefill_against_the_bit_exact_path(gpu):
    import json
    import os
    dims = dict(gpu.GEMMA_2B, n_ctx=576)
    T = 512
    cfg = gpu.make_config(wtype=O.TYPE_Q8_0, sigma_embd=0.004, **dims)
    prompt = ((1 + 7919 * np.arange(T)) % dims["n_vocab"]).astype(np.int32)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=False, fused=True)
    toks, _, _ = m.generate(prompt, 1)          # token by token through the persistent decode kernel
    dlog = m.logits().copy()
    dkv = [tuple(a[:T].view(np.float16).astype(np.float32) for a in m.kv(il)) for il in (0, 8, 17)]
    m.init_kv_cache()
    nxt = m.prefill(prompt, 0)
    glog = m.prefill_logits(T - 1)
    rng_ = float(np.abs(dlog).max())
    dev = float(np.abs(glog - dlog).max())
    top = np.sort(dlog)[-2:]
    report = {"workload": "gemma-2b q8_0, 512-token prompt, sigma_embd 0.004", "logit_absmax": rng_, "logit_max_abs_dev": dev,
              "logit_dev_rel": dev / rng_, "top2_margin_exact_path": float(top[1] - top[0]), "token_exact_path": int(toks[0]),
              "token_prefill": int(nxt), "rank_of_exact_token_in_prefill": int((glog > glog[int(toks[0])]).sum()), "kv": {}}
    for il, (dk, dv) in zip((0, 8, 17), dkv):
        gk, gv = (a[:T].view(np.float16).astype(np.float32) for a in m.kv(il))
        report["kv"][str(il)] = {"k_dev_rel": float(np.abs(gk - dk).max() / max(1.0, np.abs(dk).max())),
                                 "v_dev_rel": float(np.abs(gv - dv).max() / max(1.0, np.abs(dv).max()))}
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                                        "config3_prefill_parity.json"), "w"), indent=1)
    print(json.dumps(report))
    assert dev < FULL_PREFILL_TOL * max(1.0, rng_), report
    for il in ("0", "8", "17"):
        assert report["kv"][il]["k_dev_rel"] < FULL_PREFILL_TOL and report["kv"][il]["v_dev_rel"] < FULL_PREFILL_TOL, report
    if top[1] - top[0] > 2 * dev:               # the greedy token is pinned when the margin exceeds the measured deviation
        assert nxt == int(toks[0]), report
    assert report["rank_of_exact_token_in_prefill"] < 8, report  # and is always among the prefill path's top candidates
    m.close()
