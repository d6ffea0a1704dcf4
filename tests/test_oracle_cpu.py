"""CPU tests: pin the oracle against the committed golden vectors (gguf-py) and check its internal
consistency (the scalar ggml order against float64 models).  No GPU needed."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "codecs_gguf.npz"))


def test_fp16_matches_numpy_golden():
    L = O.lib()
    got = np.array([L.orc_fp32_to_fp16(float(v)) for v in GOLD["f32"]], dtype=np.uint16)
    assert np.array_equal(got, GOLD["f16_bits"])
    allh = np.arange(65536, dtype=np.uint16)
    exp = allh.view(np.float16).astype(np.float32)
    step = 7  # every 7th code + the specials: keeps the ctypes loop short
    for h in list(range(0, 65536, step)) + [0x0001, 0x03ff, 0x0400, 0x7bff, 0x7c00, 0xfc00, 0x8000]:
        v = np.float32(L.orc_fp16_to_fp32(h))
        assert v.tobytes() == exp[h].tobytes() or (np.isnan(v) and np.isnan(exp[h]))


@pytest.mark.parametrize("wtype,name", [(O.TYPE_Q4_0, "q4_0"), (O.TYPE_Q8_0, "q8_0")])
def test_codecs_match_gguf_golden(wtype, name):
    x = GOLD["x"]
    q = O.quantize(wtype, x)
    assert np.array_equal(q, GOLD[name]), f"{name} quantize differs from gguf-py"
    deq = O.dequantize(wtype, GOLD[name], x.shape[1])
    assert np.array_equal(deq.view(np.uint32), GOLD[name + "_deq"].view(np.uint32))


def test_quantize_q8_edge_cases():
    # all-zero block -> d = 0, id = 0, q = 0 ; ties round away from zero
    x = np.zeros((1, 64), dtype=np.float32)
    x[0, 32:] = np.array([127.0] + [0.5, -0.5, 1.5, -1.5, 2.5, -2.5] + [0.0] * 25, dtype=np.float32)
    q = O.quantize(O.TYPE_Q8_0, x)[0]
    assert q[0] == 0 and q[1] == 0 and not q[2:34].any()
    blk = q[34:]
    assert blk[:2].view(np.float16)[0] == np.float16(1.0)
    assert list(blk[2:9].view(np.int8)) == [127, 1, -1, 2, -2, 3, -3]


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
@pytest.mark.parametrize("k", [32, 2048, 3072, 16384])
def test_vec_dot_agrees_with_float64_model(wtype, k):
    rng = np.random.default_rng(k + wtype)
    w = O.quantize(wtype, rng.standard_normal((16, k)).astype(np.float32) * 0.02)
    a = O.quantize(O.TYPE_Q8_0, rng.standard_normal((2, k)).astype(np.float32))
    wd = O.dequantize(wtype, w, k).astype(np.float64)
    ad = O.dequantize(O.TYPE_Q8_0, a, k).astype(np.float64)
    exact = ad @ wd.T
    ref = O.mul_mat(wtype, O.REF, w, a, k)
    # the reference's own element tolerance for dot products: 1e-4 abs (src/kernals.cl:208)
    assert np.abs(ref - exact).max() < 1e-4
    # one accumulator, blocks in order: recomputing the sum term by term in numpy fp32 gives the same bits
    wq = w.reshape(16, k // 32, -1)
    aq = a.reshape(2, k // 32, 34)
    for r in (0, 7):
        acc = np.float32(0)
        for b in range(k // 32):
            dy = aq[0, b, :2].copy().view(np.float16)[0].astype(np.float32)
            ya = aq[0, b, 2:].copy().view(np.int8).astype(np.int32)
            if wtype == O.TYPE_Q4_0:
                dx = wq[r, b, :2].copy().view(np.float16)[0].astype(np.float32)
                qs = wq[r, b, 2:18].astype(np.int32)
                xa = np.concatenate([(qs & 15) - 8, (qs >> 4) - 8])
                term = np.float32(np.float32(np.float32(int((xa * ya).sum())) * dx) * dy)
            else:
                dx = wq[r, b, :2].copy().view(np.float16)[0].astype(np.float32)
                xa = wq[r, b, 2:34].copy().view(np.int8).astype(np.int32)
                term = np.float32(np.float32(int((xa * ya).sum())) * np.float32(dx * dy))
            acc = np.float32(acc + term)
        assert acc.view(np.uint32) == ref[0, r].view(np.uint32), (r, acc, ref[0, r])


def test_mul_mat_sub_indexing_matches_hpc_cpp():
    # dst[(col % ne1)*nb1 + (col / ne1)*nb2 + row*4] (src/hpc.cpp:26-31) with a padded dst pitch
    rng = np.random.default_rng(5)
    k, n, t = 64, 12, 3
    w = O.quantize(O.TYPE_Q8_0, rng.standard_normal((n, k)).astype(np.float32))
    a = O.quantize(O.TYPE_Q8_0, rng.standard_normal((t, k)).astype(np.float32))
    pitch = n + 5
    dst = np.full((t, pitch), np.float32(-7), dtype=np.float32)
    O.lib().orc_mul_mat_sub(2, n - 1, k, t, n, t, w.shape[1], pitch * 4, pitch * 4 * t, a.shape[1],
                            O._p(w), O._p(a), O._p(dst), O.TYPE_Q8_0, O.REF)
    full = O.mul_mat(O.TYPE_Q8_0, O.REF, w, a, k)
    assert np.array_equal(dst[:, 2:n - 1], full[:, 2:n - 1])
    assert (dst[:, :2] == -7).all() and (dst[:, n - 1:] == -7).all()


def test_rms_norm_against_float64():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(2048).astype(np.float32) * 3
    ref = O.rms_norm(x, 1e-6, O.REF)
    exact = x / np.sqrt(np.mean(x.astype(np.float64) ** 2) + 1e-6)
    assert np.abs(ref - exact).max() < 1e-5


def test_rope_neox_is_a_rotation_and_matches_closed_form():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 256)).astype(np.float32)
    y = O.rope_neox(x, 17)
    assert np.allclose((x ** 2).sum(1), (y ** 2).sum(1), rtol=1e-5)
    assert np.array_equal(O.rope_neox(x, 0), x)  # pos 0: cos 1, sin 0
    i = np.arange(128)
    theta = 17.0 * 10000.0 ** (-2.0 * i / 256)
    y0 = x[:, :128] * np.cos(theta) - x[:, 128:] * np.sin(theta)
    assert np.abs(y[:, :128] - y0).max() < 1e-4


def test_tables_and_softmax_gelu():
    et, gt = O.exp_table(), O.gelu_table()
    assert et[0] == np.float16(1.0).view(np.uint16)  # exp(+0) = 1
    assert gt[np.float16(0.0).view(np.uint16)] == 0
    x = np.array([0.0, 1.0, -1.0, 3.0, -20.0], dtype=np.float32)
    g = O.gelu(x)
    exact = 0.5 * x * (1 + np.tanh(0.7978845608 * x * (1 + 0.044715 * x * x)))
    assert np.abs(g - exact).max() < 2e-3
    rng = np.random.default_rng(3)
    s = rng.standard_normal(96).astype(np.float32)
    mask = np.zeros(96, dtype=np.float32)
    mask[70:] = -np.inf
    p = O.soft_max_ext(s, mask, 1.0)
    e = np.exp(s[:70] - s[:70].max())
    assert (p[70:] == 0).all() and abs(p.sum() - 1) < 1e-3
    assert np.abs(p[:70] - e / e.sum()).max() < 1e-3


def test_argmax_first_maximum():
    x = np.array([1, 5, 5, 2, 5], dtype=np.float32)
    assert O.argmax(x) == 1
    assert O.argmax(np.full(4, -np.inf, dtype=np.float32)) == -1  # strict '>' never fires


@pytest.mark.parametrize("H,Hkv,d,n_pos", [(8, 1, 256, 1), (8, 1, 256, 33), (4, 4, 64, 70), (4, 2, 128, 32)])
def test_attention_against_float64_softmax(H, Hkv, d, n_pos):
    rng = np.random.default_rng(H * 1000 + n_pos)
    n_ctx = 96
    q = rng.standard_normal((H, d)).astype(np.float32) * 0.3
    k = np.zeros((n_ctx, Hkv * d), dtype=np.float16)
    v = np.zeros((n_ctx, Hkv * d), dtype=np.float16)
    k[:n_pos] = rng.standard_normal((n_pos, Hkv * d)).astype(np.float16)
    v[:n_pos] = rng.standard_normal((n_pos, Hkv * d)).astype(np.float16)
    ref = O.attn_decode(O.REF, q, k.view(np.uint16), v.view(np.uint16), n_pos, n_ctx)
    # float64 softmax attention on the fp16-rounded operands
    q16 = q.astype(np.float16).astype(np.float64)
    G = H // Hkv
    for h in range(H):
        g = h // G
        kk = k[:n_pos, g * d:(g + 1) * d].astype(np.float64)
        vv = v[:n_pos, g * d:(g + 1) * d].astype(np.float64)
        s = kk @ q16[h]
        p = np.exp(s - s.max())
        exact = (p / p.sum()) @ vv
        # tolerance: P is rounded to fp16 in both forms (rel 2^-11 per term)
        assert np.abs(ref[h] - exact).max() < 4e-3


def test_synth_is_counter_based_and_scaled():
    a = O.synth_row(1234, 3, 5, 4096)
    b = O.synth_row(1234, 3, 5, 4096)
    c = O.synth_row(1234, 4, 5, 4096)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    big = np.concatenate([O.synth_row(1234, 3, r, 4096) for r in range(32)])
    assert abs(big.std() - 0.02) < 5e-4 and abs(big.mean()) < 5e-4


def test_model_generate_is_deterministic_and_prefill_equals_decode_small():
    dims = dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=64)
    m = O.Model(wtype=O.TYPE_Q4_0, sigma_embd=0.002, **dims)
    prompt = ((1 + 7919 * np.arange(10)) % 2048).astype(np.int32)
    tr, _, _ = m.generate(prompt, 8, O.REF)
    lr = m.capture("result_output")
    td, _, _ = m.generate(prompt, 8, O.REF)
    ld = m.capture("result_output")
    assert np.array_equal(tr, td) and np.array_equal(lr.view(np.uint32), ld.view(np.uint32))  # same bits every time
    # batched prefill == token-by-token decode (every mul_mat column / soft-max row is independent)
    m.reset()
    lb = m.forward(prompt, 0, O.REF)
    m.reset()
    for t, tok in enumerate(prompt):
        ls = m.forward(np.array([tok], dtype=np.int32), t, O.REF)
    assert np.array_equal(lb.view(np.uint32), ls.view(np.uint32))
    m.close()


def test_hpc_mul_mat_row_partition_covers_every_row():
    # ne01 not divisible by the thread count: ceil() partition + the row0 < ne01 guard (src/hpc.cpp:34,245-265)
    dims = dict(n_embd=64, n_layer=1, n_head=2, n_head_kv=1, head_dim=32, n_ff=96, n_vocab=101, n_ctx=8)
    m = O.Model(wtype=O.TYPE_Q8_0, n_threads=3, **dims)
    raw, rows, cols = m.tensor(0)
    assert rows == 101
    w = raw.reshape(rows, -1)
    a = O.quantize(O.TYPE_Q8_0, np.random.default_rng(0).standard_normal((2, cols)).astype(np.float32))
    dst = np.full((2, rows), np.float32(np.nan))
    O.lib().orc_hpc_mul_mat(m.h, rows, 2, 1, w.shape[1], 2, rows * 4, rows * 8, a.shape[1], cols, O._p(w), O._p(dst),
                            O.TYPE_Q8_0, O._p(a), O.REF)
    assert np.array_equal(dst, O.mul_mat(O.TYPE_Q8_0, O.REF, w, a, cols))
    m.close()
