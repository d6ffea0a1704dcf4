"""GPU parity tests, operator level: every kernel is called through the C-ABI (ctypes) and compared
with the CPU oracle on the same seeded inputs.  Bar: BIT-EXACT everywhere - integer / byte / index results and
fp32 results alike: the kernels add in the scalar ggml order (one accumulator, blocks in order; double accumulators
where ggml's scalar code has them), which is the oracle's only order."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------ K1 quantize_row_q8_0
@pytest.mark.parametrize("cols,T", [(32, 1), (2048, 1), (3072, 3), (16384, 2), (24576, 1)])
def test_quantize_row_q8_0_bit_exact(gpu, cols, T):
    rng = np.random.default_rng(cols + T)
    x = (rng.standard_normal((T, cols)) * rng.choice([1e-3, 1.0, 40.0], size=(T, 1))).astype(np.float32)
    x[0, :32] = 0.0                                  # all-zero block: d = 0, id = 0
    if cols >= 64:
        x[0, 32:64] = np.arange(32, dtype=np.float32) - 15.5  # exact .5 ties -> round half away
    _, blocks = gpu.quantize_row_q8_0(x)
    assert np.array_equal(blocks, O.quantize(O.TYPE_Q8_0, x))


def test_q8_planar_roundtrip(gpu):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 2048)).astype(np.float32)
    ref = O.quantize(O.TYPE_Q8_0, x)
    L = gpu.cuda_lib()
    raw = gpu.DeviceBuffer.from_numpy(ref)
    q8 = gpu.DeviceBuffer(gpu.q8_row_bytes(2048) * 4)
    back = gpu.DeviceBuffer(ref.nbytes)
    gpu.check(L.ggb_op_q8_from_blocks(raw.ptr, 2048, 4, ref.shape[1], q8.ptr))
    gpu.check(L.ggb_op_q8_to_blocks(q8.ptr, 2048, 4, back.ptr))
    gpu.sync()
    assert np.array_equal(back.download(np.uint8).reshape(ref.shape), ref)


# ------------------------------------------------------------------ weights
@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_weight_upload_download_roundtrip(gpu, wtype):
    rng = np.random.default_rng(1)
    w = O.quantize(wtype, rng.standard_normal((37, 2048)).astype(np.float32) * 0.02)
    W = gpu.Weight.upload(wtype, w, 2048)
    assert np.array_equal(W.download(), w)
    S = W.slice(5, 20)
    assert np.array_equal(S.download(), w[5:25])


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_weight_synth_matches_oracle_generator(gpu, wtype):
    rows, cols, seed, tid = 48, 2048, 1234, 7
    W = gpu.Weight.synth(wtype, rows, cols, seed, tid, 0.02)
    x = np.stack([O.synth_row(seed, tid, r, cols, 0.02) for r in range(rows)])
    assert np.array_equal(W.download(), O.quantize(wtype, x))
    v = gpu.DeviceBuffer(4 * 2048)
    gpu.check(gpu.cuda_lib().ggb_vector_synth(v.ptr, 2048, seed, 3, 0.02, 1.0))
    gpu.sync()
    assert np.array_equal(bits(v.download(np.float32)), bits(O.synth_row(seed, 3, 0, 2048, 0.02, 1.0)))


# ------------------------------------------------------------------ K2/K3 mat-vec
def _matvec(gpu, W, act_f32, T, n, resid=None):
    L = gpu.cuda_lib()
    q8, _ = gpu.quantize_row_q8_0(act_f32)
    dst = gpu.DeviceBuffer(4 * n * T)
    r = gpu.DeviceBuffer.from_numpy(resid) if resid is not None else None
    gpu.check(L.ggb_op_mul_mat_q8(W.h, q8.ptr, T, dst.ptr, n, r.ptr if r else None))
    gpu.sync()
    return dst.download(np.float32).reshape(T, n)


# the Gemma-2B / 7B (K, N) shapes of SURVEY.md §8a (N cut to keep the oracle in seconds) + ragged ones
SHAPES = [(2048, 2048), (2048, 256), (2048, 512), (16384, 256), (3072, 512), (4096, 384), (24576, 128),
          (2048, 1001), (64, 5), (4128, 33)]


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
@pytest.mark.parametrize("k,n", SHAPES)
def test_matvec_matches_oracle(gpu, wtype, k, n):
    rng = np.random.default_rng(k * 7 + n + wtype)
    w = O.quantize(wtype, rng.standard_normal((n, k)).astype(np.float32) * 0.02)
    x = rng.standard_normal((1, k)).astype(np.float32)
    a = O.quantize(O.TYPE_Q8_0, x)
    W = gpu.Weight.upload(wtype, w, k)
    got = _matvec(gpu, W, x, 1, n)
    ref = O.mul_mat(wtype, O.REF, w, a, k)
    assert np.array_equal(bits(got), bits(ref)), f"max abs vs the scalar ggml order {np.abs(got - ref).max()}"


def test_matvec_batch_and_residual(gpu):
    rng = np.random.default_rng(11)
    k, n, T = 2048, 300, 5
    w = O.quantize(O.TYPE_Q4_0, rng.standard_normal((n, k)).astype(np.float32) * 0.02)
    x = rng.standard_normal((T, k)).astype(np.float32)
    resid = rng.standard_normal((T, n)).astype(np.float32)
    W = gpu.Weight.upload(O.TYPE_Q4_0, w, k)
    got = _matvec(gpu, W, x, T, n, resid)
    dev = O.mul_mat(O.TYPE_Q4_0, O.REF, w, O.quantize(O.TYPE_Q8_0, x), k)
    assert np.array_equal(bits(got), bits(dev + resid))


def test_matvec_multi_qkv(gpu):
    rng = np.random.default_rng(12)
    k = 2048
    L = gpu.cuda_lib()
    ws = [O.quantize(O.TYPE_Q4_0, rng.standard_normal((n, k)).astype(np.float32) * 0.02) for n in (512, 96, 96)]
    Ws = [gpu.Weight.upload(O.TYPE_Q4_0, w, k) for w in ws]
    x = rng.standard_normal((1, k)).astype(np.float32)
    q8, _ = gpu.quantize_row_q8_0(x)
    outs = [gpu.DeviceBuffer(4 * w.shape[0]) for w in ws]
    wp = (C.c_void_p * 3)(*[W.h for W in Ws])
    dp = (C.c_void_p * 3)(*[o.ptr for o in outs])
    gpu.check(L.ggb_op_mul_mat_q8_multi(3, wp, q8.ptr, dp))
    gpu.sync()
    a = O.quantize(O.TYPE_Q8_0, x)
    for w, o in zip(ws, outs):
        assert np.array_equal(bits(o.download(np.float32)), bits(O.mul_mat(O.TYPE_Q4_0, O.REF, w, a, k)[0]))


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_ffn_gate_up_geglu_quantize(gpu, wtype):
    rng = np.random.default_rng(13)
    k, f = 2048, 256
    L = gpu.cuda_lib()
    g = O.quantize(wtype, rng.standard_normal((f, k)).astype(np.float32) * 0.05)
    u = O.quantize(wtype, rng.standard_normal((f, k)).astype(np.float32) * 0.05)
    x = rng.standard_normal((1, k)).astype(np.float32)
    a = O.quantize(O.TYPE_Q8_0, x)
    Gw, Uw = gpu.Weight.upload(wtype, g, k), gpu.Weight.upload(wtype, u, k)
    q8, _ = gpu.quantize_row_q8_0(x)
    out_q8 = gpu.DeviceBuffer(gpu.q8_row_bytes(f))
    act = gpu.DeviceBuffer(4 * f)
    blocks = gpu.DeviceBuffer(f // 32 * 34)
    gpu.check(L.ggb_op_ffn_gate_up(Gw.h, Uw.h, q8.ptr, out_q8.ptr, act.ptr))
    gpu.check(L.ggb_op_q8_to_blocks(out_q8.ptr, f, 1, blocks.ptr))
    gpu.sync()
    gate = O.mul_mat(wtype, O.REF, g, a, k)[0]
    up = O.mul_mat(wtype, O.REF, u, a, k)[0]
    expect = O.gelu(gate) * up
    assert np.array_equal(bits(act.download(np.float32)), bits(expect))
    assert np.array_equal(blocks.download(np.uint8), O.quantize(O.TYPE_Q8_0, expect[None, :])[0])


# ------------------------------------------------------------------ drop-in mul_mat (host pointers)
@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
@pytest.mark.parametrize("k,n,T", [(2048, 256, 1), (2048, 100, 7), (16384, 64, 2)])
def test_mul_mat_dropin_host_pointers(gpu, wtype, k, n, T):
    rng = np.random.default_rng(k + n + T)
    w = O.quantize(wtype, rng.standard_normal((n, k)).astype(np.float32) * 0.02)
    a = O.quantize(O.TYPE_Q8_0, rng.standard_normal((T, k)).astype(np.float32))
    got = gpu.mul_mat_host(wtype, w, a, k)
    assert np.array_equal(bits(got), bits(O.mul_mat(wtype, O.REF, w, a, k)))
    assert np.abs(got - O.mul_mat(wtype, O.REF, w, a, k)).max() < 1e-4


def test_mul_mat_dropin_strided_dst_and_errors(gpu):
    rng = np.random.default_rng(3)
    k, n, T = 64, 10, 3
    L = gpu.cuda_lib()
    w = O.quantize(O.TYPE_Q8_0, rng.standard_normal((n, k)).astype(np.float32))
    a = O.quantize(O.TYPE_Q8_0, rng.standard_normal((T, k)).astype(np.float32))
    pitch = n + 6
    dst = np.full((T, pitch), np.float32(-3))
    gpu.check(L.ggb_mul_mat(n, T, 1, w.shape[1], T, pitch * 4, pitch * 4 * T, a.shape[1], k, w.ctypes.data,
                            dst.ctypes.data, O.TYPE_Q8_0, a.ctypes.data))
    L.ggb_mul_mat_forget(w.ctypes.data)
    assert np.array_equal(bits(dst[:, :n]), bits(O.mul_mat(O.TYPE_Q8_0, O.REF, w, a, k)))
    assert (dst[:, n:] == -3).all()
    # a type without a kernel is an error, as in the reference ("kernel is null", src/hpc.cpp:162-166)
    assert L.ggb_mul_mat(n, T, 1, w.shape[1], T, n * 4, n * 4 * T, a.shape[1], k, w.ctypes.data, dst.ctypes.data,
                         10, a.ctypes.data) == -4  # Q2_K (Q4_K / Q6_K / F16 have kernels: tests/test_kquants_gpu.py)
    assert b"no kernel" in L.ggb_last_error()


def test_host_hpc_mul_mat_entry_point(gpu):
    """the reference-shaped entry (host/hpc.cpp): mul_mat(...ggml_tensor*...) with C linkage."""
    rng = np.random.default_rng(4)
    k, n, T = 2048, 64, 2
    w = O.quantize(O.TYPE_Q4_0, rng.standard_normal((n, k)).astype(np.float32) * 0.02)
    a = O.quantize(O.TYPE_Q8_0, rng.standard_normal((T, k)).astype(np.float32))
    dst = np.zeros((T, n), dtype=np.float32)

    class Tensor(C.Structure):
        _fields_ = [("data", C.c_void_p)]

    s0, s1, d = Tensor(w.ctypes.data), Tensor(None), Tensor(dst.ctypes.data)
    gpu.host_lib().mul_mat(n, T, 1, w.shape[1], T, n * 4, n * 4 * T, a.shape[1], k, C.addressof(s0), C.addressof(s1),
                           C.addressof(d), None, O.TYPE_Q4_0, a.ctypes.data)
    gpu.cuda_lib().ggb_mul_mat_forget(w.ctypes.data)
    assert np.array_equal(bits(dst), bits(O.mul_mat(O.TYPE_Q4_0, O.REF, w, a, k)))


# ------------------------------------------------------------------ K5 rms_norm
@pytest.mark.parametrize("E,T", [(2048, 1), (3072, 4), (256, 2)])
def test_rms_norm_mul_quantize(gpu, E, T):
    rng = np.random.default_rng(E)
    L = gpu.cuda_lib()
    x = (rng.standard_normal((T, E)) * 3).astype(np.float32)
    w = (1 + 0.02 * rng.standard_normal(E)).astype(np.float32)
    xd, wd = gpu.DeviceBuffer.from_numpy(x), gpu.DeviceBuffer.from_numpy(w)
    y, q8, blocks = gpu.DeviceBuffer(x.nbytes), gpu.DeviceBuffer(gpu.q8_row_bytes(E) * T), gpu.DeviceBuffer(E // 32 * 34 * T)
    gpu.check(L.ggb_op_rms_norm(xd.ptr, wd.ptr, 1e-6, E, T, y.ptr, q8.ptr))
    gpu.check(L.ggb_op_q8_to_blocks(q8.ptr, E, T, blocks.ptr))
    gpu.sync()
    dev = np.stack([O.rms_norm(x[t], 1e-6, O.REF) * w for t in range(T)])
    got = y.download(np.float32).reshape(T, E)
    assert np.array_equal(bits(got), bits(dev))
    assert np.array_equal(blocks.download(np.uint8).reshape(T, -1), O.quantize(O.TYPE_Q8_0, dev))


# ------------------------------------------------------------------ element-wise nodes
def test_elementwise_nodes_bit_exact(gpu):
    rng = np.random.default_rng(21)
    L = gpu.cuda_lib()
    n = 5000
    a = rng.standard_normal(n).astype(np.float32) * 4
    b = rng.standard_normal(n).astype(np.float32)
    ad, bd, y = gpu.DeviceBuffer.from_numpy(a), gpu.DeviceBuffer.from_numpy(b), gpu.DeviceBuffer(4 * n)
    for fn, expect in ((L.ggb_op_add, a + b), (L.ggb_op_mul, a * b)):
        gpu.check(fn(ad.ptr, bd.ptr, n, y.ptr))
        gpu.sync()
        assert np.array_equal(bits(y.download(np.float32)), bits(expect))
    gpu.check(L.ggb_op_scale(ad.ptr, np.float32(45.254833995939045), n, y.ptr))
    gpu.sync()
    assert np.array_equal(bits(y.download(np.float32)), bits(a * np.float32(45.254833995939045)))
    gpu.check(L.ggb_op_gelu(ad.ptr, n, y.ptr))
    gpu.sync()
    assert np.array_equal(bits(y.download(np.float32)), bits(O.gelu(a)))
    h = gpu.DeviceBuffer(2 * n)
    gpu.check(L.ggb_op_cpy_f32_f16(ad.ptr, n, h.ptr))
    gpu.sync()
    assert np.array_equal(h.download(np.uint16), a.astype(np.float16).view(np.uint16))


@pytest.mark.parametrize("nc,n_valid", [(32, 1), (96, 70), (2048, 2048), (2048, 1500)])
def test_soft_max_ext(gpu, nc, n_valid):
    rng = np.random.default_rng(nc + n_valid)
    L = gpu.cuda_lib()
    rows = 3
    x = rng.standard_normal((rows, nc)).astype(np.float32) * 2
    mask = np.zeros((rows, nc), dtype=np.float32)
    mask[:, n_valid:] = -np.inf
    xd, md, y = gpu.DeviceBuffer.from_numpy(x), gpu.DeviceBuffer.from_numpy(mask), gpu.DeviceBuffer(x.nbytes)
    gpu.check(L.ggb_op_soft_max_ext(xd.ptr, md.ptr, 1.0, nc, rows, y.ptr))
    gpu.sync()
    got = y.download(np.float32).reshape(rows, nc)
    ref = np.stack([O.soft_max_ext(x[r], mask[r], 1.0) for r in range(rows)])
    # identical table lookups; only the order of the double sum differs -> at most 1 ulp of the scale
    assert (got[:, n_valid:] == 0).all()
    assert np.abs(got - ref).max() <= 2e-7 * ref.max()


def test_rope_neox_bit_exact(gpu):
    rng = np.random.default_rng(31)
    L = gpu.cuda_lib()
    n_ctx, d, H, T = 300, 256, 8, 4
    gpu.check(L.ggb_rope_table_init(n_ctx, d, 10000.0))
    x = rng.standard_normal((T, H, d)).astype(np.float32)
    pos = np.array([0, 1, 77, 299], dtype=np.int32)
    xd, pd, y = gpu.DeviceBuffer.from_numpy(x), gpu.DeviceBuffer.from_numpy(pos), gpu.DeviceBuffer(x.nbytes)
    gpu.check(L.ggb_op_rope_neox(xd.ptr, pd.ptr, d, H, T, y.ptr))
    gpu.sync()
    got = y.download(np.float32).reshape(T, H, d)
    for t in range(T):
        assert np.array_equal(bits(got[t]), bits(O.rope_neox(x[t], int(pos[t]))))


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_get_rows_scale(gpu, wtype):
    rng = np.random.default_rng(41)
    L = gpu.cuda_lib()
    V, E = 500, 2048
    w = O.quantize(wtype, rng.standard_normal((V, E)).astype(np.float32) * 0.02)
    W = gpu.Weight.upload(wtype, w, E)
    ids = np.array([0, 499, 17, 17, 250], dtype=np.int32)
    idd, y = gpu.DeviceBuffer.from_numpy(ids), gpu.DeviceBuffer(4 * E * len(ids))
    scale = np.float32(np.sqrt(np.float32(E)))
    gpu.check(L.ggb_op_get_rows(W.h, idd.ptr, len(ids), scale, y.ptr))
    gpu.sync()
    expect = O.dequantize(wtype, w[ids], E) * scale
    assert np.array_equal(bits(y.download(np.float32).reshape(len(ids), E)), bits(expect))


def test_argmax_first_maximum_and_ties(gpu):
    rng = np.random.default_rng(51)
    L = gpu.cuda_lib()
    for n in (10, 4097, 256000):
        x = rng.standard_normal(n).astype(np.float32)
        x[[n // 3, n // 2, n - 1]] = 9.5  # three-way tie: lowest index wins (strict '>')
        xd, out = gpu.DeviceBuffer.from_numpy(x), gpu.DeviceBuffer(4)
        gpu.check(L.ggb_op_argmax(xd.ptr, n, out.ptr))
        gpu.sync()
        assert int(out.download(np.int32)[0]) == O.argmax(x) == n // 3
    x = np.full(1000, -np.inf, dtype=np.float32)
    xd, out = gpu.DeviceBuffer.from_numpy(x), gpu.DeviceBuffer(4)
    gpu.check(L.ggb_op_argmax(xd.ptr, 1000, out.ptr))
    gpu.sync()
    assert int(out.download(np.int32)[0]) == -1 == O.argmax(x)


# ------------------------------------------------------------------ K6 fused + K4 attention
@pytest.mark.parametrize("H,Hkv,d,n_ctx,positions", [
    (8, 1, 256, 128, [0, 1, 30, 31, 32, 33, 100, 127]),   # Gemma-2B MQA, ragged chunk edges
    (16, 16, 256, 96, [0, 47, 95]),                       # Gemma-7B MHA
    (4, 2, 64, 64, [5, 63]),                              # grouped, small head
])
def test_rope_store_and_attention_decode(gpu, H, Hkv, d, n_ctx, positions):
    rng = np.random.default_rng(H + n_ctx)
    L = gpu.cuda_lib()
    gpu.check(L.ggb_rope_table_init(n_ctx, d, 10000.0))
    kv = Hkv * d
    kc_h = np.zeros((n_ctx, kv), dtype=np.float16)
    vc_h = np.zeros((n_ctx, kv), dtype=np.float16)
    last = max(positions)
    kc_h[:last] = rng.standard_normal((last, kv)).astype(np.float16)
    vc_h[:last] = rng.standard_normal((last, kv)).astype(np.float16)
    kc, vc = gpu.DeviceBuffer.from_numpy(kc_h), gpu.DeviceBuffer.from_numpy(vc_h)
    scratch = gpu.DeviceBuffer(int(L.ggb_attn_scratch_bytes(H, d, n_ctx)))
    q_scale = np.float32(1.0) / np.sqrt(np.float32(d))
    for pos in positions:
        q = rng.standard_normal((H, d)).astype(np.float32)
        k = rng.standard_normal((Hkv, d)).astype(np.float32)
        v = rng.standard_normal((Hkv, d)).astype(np.float32)
        qd, kd, vd = gpu.DeviceBuffer.from_numpy(q), gpu.DeviceBuffer.from_numpy(k), gpu.DeviceBuffer.from_numpy(v)
        pd = gpu.DeviceBuffer.from_numpy(np.array([pos], dtype=np.int32))
        out, q8, blocks = gpu.DeviceBuffer(4 * H * d), gpu.DeviceBuffer(gpu.q8_row_bytes(H * d)), gpu.DeviceBuffer(H * d // 32 * 34)
        gpu.check(L.ggb_op_rope_kv_store(qd.ptr, kd.ptr, vd.ptr, pd.ptr, H, Hkv, d, q_scale, kc.ptr, vc.ptr))
        gpu.check(L.ggb_op_attn_decode(qd.ptr, kc.ptr, vc.ptr, pd.ptr, H, Hkv, d, n_ctx, scratch.ptr, out.ptr, q8.ptr))
        gpu.check(L.ggb_op_q8_to_blocks(q8.ptr, H * d, 1, blocks.ptr))
        gpu.sync()
        # oracle: rope + scale + store, then attention
        q_r = O.rope_neox(q, pos) * q_scale
        kc_h[pos] = O.rope_neox(k, pos).astype(np.float16).reshape(-1)
        vc_h[pos] = v.astype(np.float16).reshape(-1)
        assert np.array_equal(bits(qd.download(np.float32)).reshape(H, d), bits(q_r))
        assert np.array_equal(kc.download(np.uint16).reshape(n_ctx, kv)[pos], kc_h[pos].view(np.uint16))
        assert np.array_equal(vc.download(np.uint16).reshape(n_ctx, kv)[pos], vc_h[pos].view(np.uint16))
        # the reference's two-pass form (kq -> soft_max_ext -> fp16 P . V), bit for bit
        dev = O.attn_decode(O.REF, q_r, kc_h.view(np.uint16), vc_h.view(np.uint16), pos + 1, n_ctx)
        got = out.download(np.float32).reshape(H, d)
        assert np.array_equal(bits(got), bits(dev)), f"pos {pos}: max abs {np.abs(got - dev).max()}"
        assert np.array_equal(blocks.download(np.uint8), O.quantize(O.TYPE_Q8_0, dev.reshape(1, -1))[0])
