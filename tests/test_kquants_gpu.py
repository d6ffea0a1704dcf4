"""K-quant kernels against the oracle, bit for bit: quantize_row_q8_K, Q4_K / Q6_K x Q8_K and
F16 x F16 through the device operator and through the drop-in entry ggb_mul_mat with host pointers
(ref: src/kernals.cl:48-164, src/hpc.cpp:132-143, 216-277)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("k", [256, 2048, 16384])
def test_quantize_row_q8_K_bitwise(gpu, k):
    rng = np.random.default_rng(k)
    x = (rng.standard_normal((5, k)) * rng.choice([1e-3, 1.0, 40.0], size=(5, 1))).astype(np.float32)
    x[1, :256] = 0.0                      # an all-zero block
    x[2, 7] = -np.abs(x[2]).max() * 2     # negative extreme -> q = 127 clamp on the opposite sign
    x[3, 10] = x[3, 3] = np.float32(3.5)  # a tie: the first maximum decides the sign
    x[3, 11] = np.float32(-3.5)
    x[3, :256] = np.clip(x[3, :256], -3.5, 3.5)
    assert np.array_equal(gpu.quantize_q8_K(x), O.quantize_q8_K(x))


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_K, O.TYPE_Q6_K])
@pytest.mark.parametrize("n,k,t", [(64, 2048, 1), (40, 16384, 3), (13, 256, 2), (256, 3072, 1)])
def test_kquant_mul_mat_matches_scalar_reference_order(gpu, wtype, n, k, t):
    rng = np.random.default_rng(n + k + wtype)
    w = O.random_kquant_rows(rng, wtype, n, k)
    act = O.quantize_q8_K(rng.standard_normal((t, k)) * 2)
    want = O.mul_mat_blocks(wtype, w, act, k)
    got = gpu.mul_mat_host(wtype, w, act, k)
    assert np.array_equal(bits(got), bits(want))
    # the device quantizer feeds the same kernel: end to end from fp32 activations
    act_dev = gpu.quantize_q8_K(np.ascontiguousarray(rng.standard_normal((1, k)), dtype=np.float32))
    assert np.array_equal(bits(gpu.mul_mat_host(wtype, w, act_dev, k)), bits(O.mul_mat_blocks(wtype, w, act_dev, k)))


@pytest.mark.parametrize("n,k,t", [(32, 256, 1), (17, 2050, 2), (8, 8191, 1)])
def test_f16_mul_mat_double_accumulator(gpu, n, k, t):
    rng = np.random.default_rng(k)
    kp = k + (k & 1)  # row pitch stays 4-byte friendly; the dot runs over k elements
    w = np.zeros((n, kp), np.float16)
    a = np.zeros((t, kp), np.float16)
    w[:, :k] = rng.standard_normal((n, k))
    a[:, :k] = rng.standard_normal((t, k)) * 8
    wb, ab = w.view(np.uint8).reshape(n, kp * 2), a.view(np.uint8).reshape(t, kp * 2)
    want = O.mul_mat_blocks(O.TYPE_F16, wb, ab, k)
    got = gpu.mul_mat_host(O.TYPE_F16, wb, ab, k)
    assert np.array_equal(bits(got), bits(want))


def test_mul_mat_strided_dst_and_batches(gpu):
    """ne12 > 1 and a padded dst pitch: the reference's addressing dst + (col % ne1)*nb1 + (col / ne1)*nb2"""
    import ctypes as C
    rng = np.random.default_rng(1)
    n, k, ne11, ne12 = 24, 512, 3, 2
    w = O.random_kquant_rows(rng, O.TYPE_Q4_K, n, k)
    act = O.quantize_q8_K(rng.standard_normal((ne11 * ne12, k)))
    want = O.mul_mat_blocks(O.TYPE_Q4_K, w, act, k)
    pitch = n + 8
    dst = np.full((ne12, ne11, pitch), np.float32(-7), dtype=np.float32)
    L = gpu.cuda_lib()
    gpu.check(L.ggb_mul_mat(n, ne11, ne12, w.shape[1], ne11, pitch * 4, pitch * 4 * ne11, act.shape[1], k,
                          w.ctypes.data, dst.ctypes.data, O.TYPE_Q4_K, act.ctypes.data), "ggb_mul_mat")
    L.ggb_mul_mat_forget(w.ctypes.data)
    assert np.array_equal(bits(dst[:, :, :n].reshape(-1, n)), bits(want))
    assert np.all(dst[:, :, n:] == -7)


def test_unsupported_type_is_reported_not_computed(gpu):
    L = gpu.cuda_lib()
    buf = np.zeros(1024, np.uint8)
    assert L.ggb_mul_mat(1, 1, 1, 64, 1, 4, 4, 64, 32, buf.ctypes.data, buf.ctypes.data, 10, buf.ctypes.data) == -4
    assert b"no kernel" in L.ggb_last_error()


@pytest.mark.parametrize("n,k,t", [(40, 2048, 2), (9, 256, 1), (300, 4096, 1), (17, 16384, 3), (64, 1280, 1)])
def test_q4_K_device_operators_raw_and_packed_layouts(gpu, n, k, t):
    """the AoS operator (ggb_op_mul_mat_blocks) and the packed one (ggb_op_q4k_repack + ggb_op_mul_mat_q4k_packed) on
    device pointers: both bit-identical to the scalar code; K with a partial group of four super-blocks included"""
    rng = np.random.default_rng(n * k + t)
    L = gpu.cuda_lib()
    w = O.random_kquant_rows(rng, O.TYPE_Q4_K, n, k)
    act = O.quantize_q8_K(rng.standard_normal((t, k)) * 1.5)
    want = O.mul_mat_blocks(O.TYPE_Q4_K, w, act, k)
    wd, ad = gpu.DeviceBuffer.from_numpy(w), gpu.DeviceBuffer.from_numpy(act)
    dst = gpu.DeviceBuffer(4 * n * t)
    gpu.check(L.ggb_op_mul_mat_blocks(O.TYPE_Q4_K, wd.ptr, n, w.shape[1], k, ad.ptr, act.shape[1], t, dst.ptr, n))
    gpu.sync()
    assert np.array_equal(bits(dst.download(np.float32).reshape(t, n)), bits(want))
    packed = gpu.DeviceBuffer(w.nbytes)
    gpu.check(L.ggb_op_q4k_repack(wd.ptr, n, w.shape[1], k, packed.ptr))
    gpu.check(L.ggb_buffer_memset(dst.ptr, 0, 4 * n * t))
    gpu.check(L.ggb_op_mul_mat_q4k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], t, dst.ptr, n))
    gpu.sync()
    assert np.array_equal(bits(dst.download(np.float32).reshape(t, n)), bits(want))


@pytest.mark.parametrize("n,k,t", [(40, 2048, 2), (9, 256, 1), (300, 4096, 1), (17, 16384, 3), (64, 1280, 1)])
def test_q6_K_device_operators_raw_and_packed_layouts(gpu, n, k, t):
    rng = np.random.default_rng(7 * n * k + t)
    L = gpu.cuda_lib()
    w = O.random_kquant_rows(rng, O.TYPE_Q6_K, n, k)
    act = O.quantize_q8_K(rng.standard_normal((t, k)) * 1.5)
    want = O.mul_mat_blocks(O.TYPE_Q6_K, w, act, k)
    wd, ad = gpu.DeviceBuffer.from_numpy(w), gpu.DeviceBuffer.from_numpy(act)
    dst = gpu.DeviceBuffer(4 * n * t)
    gpu.check(L.ggb_op_mul_mat_blocks(O.TYPE_Q6_K, wd.ptr, n, w.shape[1], k, ad.ptr, act.shape[1], t, dst.ptr, n))
    gpu.sync()
    assert np.array_equal(bits(dst.download(np.float32).reshape(t, n)), bits(want))
    packed = gpu.DeviceBuffer(L.ggb_q6k_packed_bytes(n, k))
    gpu.check(L.ggb_op_q6k_repack(wd.ptr, n, w.shape[1], k, packed.ptr))
    gpu.check(L.ggb_buffer_memset(dst.ptr, 0, 4 * n * t))
    gpu.check(L.ggb_op_mul_mat_q6k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], t, dst.ptr, n))
    gpu.sync()
    assert np.array_equal(bits(dst.download(np.float32).reshape(t, n)), bits(want))


# ------------------------------------------------------------------ a K-quant MODEL (the author's q4_k_m mix in miniature)
KQ_MODELS = {
    # Q4_K matrices, Q6_K tied embedding / output matrix (what a q4_k_m file holds)
    "mqa-q4_K": (O.TYPE_Q4_K, dict(n_embd=256, n_layer=2, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=512, n_ctx=48)),
    # all Q6_K, MHA, n_head * head_dim != n_embd, K = 2048 in the down projection (eight super-blocks)
    "mha-q6_K": (O.TYPE_Q6_K, dict(n_embd=512, n_layer=2, n_head=4, n_head_kv=4, head_dim=256, n_ff=2048, n_vocab=768, n_ctx=40)),
}


@pytest.mark.parametrize("name", list(KQ_MODELS))
def test_kquant_model_generate_matches_oracle_bitwise(gpu, name):
    """K-quant mode of the host model (set_tensor_typed -> packed K-quant mat-vecs, quantize_row_q8_K, get_rows on the
    blocks, the fp32 operators of the Q4_0 / Q8_0 path) against the oracle running the same node list on the same
    blocks: tokens, every captured intermediate of every layer and the logits, bit for bit."""
    wtype, dims = KQ_MODELS[name]
    om = O.Model(wtype=wtype, sigma_embd=0.005, **dims)
    cfg = gpu.make_config(wtype=wtype, sigma_embd=0.005, **dims)
    m = gpu.GemmaModel(cfg)
    for tid in range(2 + 9 * dims["n_layer"]):
        raw, rows, cols = om.tensor(tid)
        is_norm = tid == 1 or (tid >= 2 and (tid - 2) % 9 in (0, 5))
        if is_norm:
            m.set_tensor_typed(tid, gpu.TYPE_F32, raw)
        else:
            ttype = gpu.TYPE_Q6_K if raw.nbytes == rows * (cols // 256) * 210 else gpu.TYPE_Q4_K
            m.set_tensor_typed(tid, ttype, raw)
    assert m.kquant_mode
    m.init_kv_cache()
    m.build_compute_graph(capture_acts=True)
    assert not m.is_fused
    prompt = ((1 + 7919 * np.arange(7)) % dims["n_vocab"]).astype(np.int32)
    toks, _, _ = m.generate(prompt, 10)
    otoks, _, _ = om.generate(prompt, 10, O.REF)
    assert np.array_equal(toks, otoks), (toks, otoks)
    assert np.array_equal(bits(m.logits()), bits(om.capture("result_output")))
    for il in range(dims["n_layer"]):
        for cap in ("kqv_out", "sa_out", "ffn_act", "l_out"):
            assert np.array_equal(bits(m.capture(cap, il)), bits(om.capture(cap, il))), (cap, il)
    m.close(); om.close()


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_K, O.TYPE_Q6_K])
@pytest.mark.parametrize("n,k,t", [(40, 2048, 2), (9, 256, 1), (33, 4096, 3)])
def test_packed_mat_vec_on_fp32_activations_quantises_like_quantize_row_q8_K(gpu, wtype, n, k, t):
    """INIT stage fused into the mat-vec: fp32 rows in, same bits as quantize_row_q8_K + vec_dot of the scalar code"""
    rng = np.random.default_rng(n + k + t + wtype)
    L = gpu.cuda_lib()
    w = O.random_kquant_rows(rng, wtype, n, k)
    x = (rng.standard_normal((t, k)) * 1.5).astype(np.float32)
    x[0, :256] = 0.0  # an all-zero super-block
    want = O.mul_mat_blocks(wtype, w, O.quantize_q8_K(x), k)
    wd, xd = gpu.DeviceBuffer.from_numpy(w), gpu.DeviceBuffer.from_numpy(x)
    if wtype == O.TYPE_Q4_K:
        packed = gpu.DeviceBuffer(w.nbytes)
        gpu.check(L.ggb_op_q4k_repack(wd.ptr, n, w.shape[1], k, packed.ptr))
    else:
        packed = gpu.DeviceBuffer(L.ggb_q6k_packed_bytes(n, k))
        gpu.check(L.ggb_op_q6k_repack(wd.ptr, n, w.shape[1], k, packed.ptr))
    dst = gpu.DeviceBuffer(4 * n * t)
    gpu.check(L.ggb_op_mul_mat_kq_packed_f32(wtype, packed.ptr, n, k, xd.ptr, t, dst.ptr, n))
    gpu.sync()
    assert np.array_equal(bits(dst.download(np.float32).reshape(t, n)), bits(want))
