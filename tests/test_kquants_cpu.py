"""K-quant dots: the oracle's restatements pinned (a) to the REFERENCE's own scalar code compiled from
/root/reference/src/kernals.cl into oracle/_ref (bit-exact: Q4_K x Q8_K and F16 x F16), (b) to gguf-py's block layouts
(dequantize_blocks) and a float64 model (Q4_K and Q6_K; the reference's Q6_K copy is broken, see oracle/ggml_oracle.h)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libkernals_ref.so")


def _oracle():
    L = O.lib()
    L.orc_vec_dot_f16.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_vec_dot_f16.restype = None
    return L


def random_blocks(rng, wtype, rows, k):
    return O.random_kquant_rows(rng, wtype, rows, k)


def q8k(L, x):
    return O.quantize_q8_K(x)


def dot(fn, k, w, a):
    s = C.c_float(0)
    fn(k, C.byref(s), w.ctypes.data, a.ctypes.data)
    return np.float32(s.value)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (make -C oracle _ref)")
def test_oracle_matches_the_reference_code_bit_for_bit():
    R = C.CDLL(REF_SO)
    for f in (R.ref_vec_dot_q4_K_q8_K, R.ref_vec_dot_f16):
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        f.restype = None
    assert (R.ref_sizeof_block_q4_K(), R.ref_sizeof_block_q6_K(), R.ref_sizeof_block_q8_K()) == (144, 210, 292)
    L = _oracle()
    rng = np.random.default_rng(3)
    for k in (256, 2048, 16384):
        w = random_blocks(rng, 12, 16, k)
        a = q8k(L, rng.standard_normal((2, k)) * 3)
        for r in range(16):
            for c in range(2):
                got, ref = dot(L.orc_vec_dot_q4_K_q8_K, k, w[r], a[c]), dot(R.ref_vec_dot_q4_K_q8_K, k, w[r], a[c])
                assert got.view(np.uint32) == ref.view(np.uint32), (k, r, c, got, ref)
    # fp16 x fp16 with a double accumulator (src/kernals.cl:158-164) == orc_vec_dot_f16
    L.orc_vec_dot_f16.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in (32, 256, 2048):
        x = (rng.standard_normal(n)).astype(np.float16)
        y = (rng.standard_normal(n)).astype(np.float16)
        assert dot(L.orc_vec_dot_f16, n, x, y).view(np.uint32) == dot(R.ref_vec_dot_f16, n, x, y).view(np.uint32)


@pytest.mark.parametrize("wtype", [12, 14])
def test_block_layouts_and_dots_against_gguf_py_and_float64(wtype):
    import gguf
    from gguf import quants
    L = _oracle()
    rng = np.random.default_rng(wtype)
    k = 1024
    w = random_blocks(rng, wtype, 8, k)
    qt = gguf.GGMLQuantizationType.Q4_K if wtype == 12 else gguf.GGMLQuantizationType.Q6_K
    want = quants.dequantize(w, qt)
    got = np.zeros((8, k), dtype=np.float32)
    fn = L.orc_dequantize_row_q4_K if wtype == 12 else L.orc_dequantize_row_q6_K
    for r in range(8):
        fn(w[r].ctypes.data, got[r].ctypes.data, k)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7)  # same layout (gguf-py multiplies in another order)
    x = rng.standard_normal((1, k)).astype(np.float32) * 2
    a = q8k(L, x)
    blk = a.reshape(-1, 292)
    ad = blk[:, 0:4].copy().view(np.float32)[:, 0]
    aq = blk[:, 4:260].copy().view(np.int8).astype(np.float64) * ad[:, None].astype(np.float64)
    vd = L.orc_vec_dot_q4_K_q8_K if wtype == 12 else L.orc_vec_dot_q6_K_q8_K
    for r in range(8):
        model = float(np.dot(want[r].astype(np.float64), aq.reshape(-1)))
        assert abs(float(dot(vd, k, w[r], a[0])) - model) <= 1e-4 * max(1.0, abs(model))  # the reference's own tolerance (src/kernals.cl:208)


def test_quantize_row_q8_K_properties():
    L = _oracle()
    rng = np.random.default_rng(9)
    x = rng.standard_normal((3, 512)).astype(np.float32)
    x[1, :256] = 0.0
    a = q8k(L, x).reshape(3, 2, 292)
    for r in range(3):
        for b in range(2):
            d = a[r, b, 0:4].copy().view(np.float32)[0]
            qs = a[r, b, 4:260].copy().view(np.int8).astype(np.int32)
            bs = a[r, b, 260:292].copy().view(np.int16)
            assert np.array_equal(bs, qs.reshape(16, 16).sum(axis=1))
            xb = x[r, 256 * b:256 * b + 256]
            if not np.any(xb):
                assert d == 0 and not np.any(qs)
            else:
                assert np.abs(qs * d - xb).max() <= abs(d) * 1.0 + 1e-7 and qs.min() >= -128 and qs.max() <= 127
