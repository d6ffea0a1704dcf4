"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

TYPE_F32, TYPE_F16, TYPE_Q4_0, TYPE_Q8_0 = 0, 1, 2, 8
REF = 0
QK = 32
BLOCK_BYTES = {TYPE_Q4_0: 18, TYPE_Q8_0: 34}


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in
            ("ggml_oracle.c", "gemma_oracle.c", "ggml_oracle.h", "gemma_oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


class Config(C.Structure):
    _fields_ = [("n_embd", C.c_int32), ("n_layer", C.c_int32), ("n_head", C.c_int32),
                ("n_head_kv", C.c_int32), ("head_dim", C.c_int32), ("n_ff", C.c_int32),
                ("n_vocab", C.c_int32), ("n_ctx", C.c_int32), ("eps", C.c_float),
                ("rope_base", C.c_float), ("wtype", C.c_int32), ("pad_", C.c_int32),
                ("seed", C.c_uint64), ("sigma_embd", C.c_float), ("sigma_mat", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    # the GPU box has no compiler guarantee: use the prebuilt .so when it is there
    if not os.path.exists(LIB_PATH):
        build()
    else:
        try:
            build()
        except Exception:
            pass
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
    sig = {
        "orc_fp16_to_fp32": (f32, [C.c_uint16]),
        "orc_fp32_to_fp16": (C.c_uint16, [f32]),
        "orc_row_size": (C.c_size_t, [i32, i64]),
        "orc_quantize_row_q4_0": (None, [vp, vp, i64]),
        "orc_quantize_row_q8_0": (None, [vp, vp, i64]),
        "orc_dequantize_row": (None, [i32, vp, vp, i64]),
        "orc_vec_dot": (None, [i32, i32, i32, vp, vp, vp]),
        "orc_mul_mat_sub": (None, [i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, vp, vp, vp, i32, i32]),
        "orc_rms_norm": (None, [i64, vp, f32, vp]),
        "orc_mul": (None, [i64, vp, vp, vp]),
        "orc_rope_cache": (None, [i32, i32, f32, vp, vp]),
        "orc_rope_neox": (None, [i32, i32, i32, f32, vp, vp]),
        "orc_exp_table": (vp, []),
        "orc_gelu_table": (vp, []),
        "orc_gelu": (None, [i64, vp, vp]),
        "orc_soft_max_ext": (None, [i64, vp, vp, f32, vp]),
        "orc_argmax": (i32, [i64, vp]),
        "orc_attn_decode_ref": (None, [i32, i32, i32, i64, i64, vp, vp, vp, i64, vp]),
        "orc_synth_value": (f32, [u64, u64, u64, f32]),
        "orc_synth_row": (None, [u64, u64, i64, i64, f32, f32, vp]),
        "orc_model_create": (vp, [C.POINTER(Config), i32]),
        "orc_model_free": (None, [vp]),
        "orc_model_tensor": (vp, [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "orc_model_reset": (None, [vp]),
        "orc_model_fill_kv_synth": (None, [vp, u64]),
        "orc_model_forward": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "orc_model_generate": (i32, [vp, vp, i32, i32, i32, vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "orc_model_capture": (vp, [vp, C.c_char_p, i32, C.POINTER(i64)]),
        "orc_model_kcache": (vp, [vp, i32]),
        "orc_model_vcache": (vp, [vp, i32]),
        "orc_hpc_mul_mat": (None, [vp, i64, i64, i64, i64, i64, i64, i64, C.c_size_t, i64, vp, vp, i32, vp, i32]),
        "orc_model_threads": (i32, [vp]),
        "orc_vec_dot_q4_K_q8_K": (None, [i32, vp, vp, vp]), "orc_vec_dot_q6_K_q8_K": (None, [i32, vp, vp, vp]),
        "orc_quantize_row_q8_K": (None, [vp, vp, i64]),
        "orc_dequantize_row_q4_K": (None, [vp, vp, i64]), "orc_dequantize_row_q6_K": (None, [vp, vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fp32_to_fp16(x):
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    flat_in, flat_out = x.ravel(), out.ravel()
    for i in range(flat_in.size):
        flat_out[i] = L.orc_fp32_to_fp16(float(flat_in[i]))
    return out


def row_size(wtype, k):
    return int(lib().orc_row_size(wtype, k))


def quantize(wtype, x):
    """x: [rows, k] float32 -> uint8 [rows, row_size] of ggml AoS blocks."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    rows, k = x.shape
    rs = row_size(wtype, k)
    out = np.zeros((rows, rs), dtype=np.uint8)
    fn = lib().orc_quantize_row_q4_0 if wtype == TYPE_Q4_0 else lib().orc_quantize_row_q8_0
    for r in range(rows):
        fn(_p(x[r]), C.c_void_p(out[r].ctypes.data), k)
    return out


def dequantize(wtype, blocks, k):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    if blocks.ndim == 1:
        blocks = blocks[None, :]
    out = np.empty((blocks.shape[0], k), dtype=np.float32)
    for r in range(blocks.shape[0]):
        lib().orc_dequantize_row(wtype, C.c_void_p(blocks[r].ctypes.data), C.c_void_p(out[r].ctypes.data), k)
    return out


def vec_dot(wtype, order, k, xrow, yrow):
    s = C.c_float(0)
    lib().orc_vec_dot(wtype, order, k, C.byref(s), _p(xrow), _p(yrow))
    return np.float32(s.value)


def mul_mat(wtype, order, w_blocks, act_q8, k):
    """w_blocks: uint8 [N, rs_w]; act_q8: uint8 [T, rs_a] -> float32 [T, N] (hpc.cpp indexing)."""
    w_blocks = np.ascontiguousarray(w_blocks)
    act_q8 = np.ascontiguousarray(act_q8)
    n, t = w_blocks.shape[0], act_q8.shape[0]
    out = np.zeros((t, n), dtype=np.float32)
    lib().orc_mul_mat_sub(0, n, k, t, n, t, w_blocks.shape[1], n * 4, n * 4 * t, act_q8.shape[1],
                          _p(w_blocks), _p(act_q8), _p(out), wtype, order)
    return out


def rms_norm(x, eps, order=REF):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_rms_norm(x.size, _p(x), eps, _p(y))
    return y


def rope_cache(pos, n_dims, base=10000.0):
    c = np.empty(n_dims // 2, dtype=np.float32)
    s = np.empty(n_dims // 2, dtype=np.float32)
    lib().orc_rope_cache(pos, n_dims, base, _p(c), _p(s))
    return c, s


def rope_neox(x, pos, base=10000.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    n_heads, n_dims = x.shape
    y = np.empty_like(x)
    lib().orc_rope_neox(n_dims, n_heads, pos, base, _p(x), _p(y))
    return y


def exp_table():
    return np.ctypeslib.as_array(C.cast(lib().orc_exp_table(), C.POINTER(C.c_uint16)), shape=(65536,)).copy()


def gelu_table():
    return np.ctypeslib.as_array(C.cast(lib().orc_gelu_table(), C.POINTER(C.c_uint16)), shape=(65536,)).copy()


def gelu(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_gelu(x.size, _p(x), _p(y))
    return y


def soft_max_ext(x, mask=None, scale=1.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.float32)
    lib().orc_soft_max_ext(x.size, _p(x), None if m is None else _p(m), scale, _p(y))
    return y


def argmax(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return int(lib().orc_argmax(x.size, _p(x)))


def attn_decode(order, q, kcache, vcache, n_pos, n_ctx=None):
    """q: [H, d] f32; kcache/vcache: uint16 [n_ctx, Hkv*d]."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    H, d = q.shape
    kcache = np.ascontiguousarray(kcache, dtype=np.uint16)
    vcache = np.ascontiguousarray(vcache, dtype=np.uint16)
    stride = kcache.shape[1]
    Hkv = stride // d
    out = np.empty_like(q)
    lib().orc_attn_decode_ref(H, Hkv, d, n_pos, n_ctx or kcache.shape[0], _p(q), _p(kcache), _p(vcache),
                              stride, _p(out))
    return out


def synth_row(seed, tensor_id, row, k, sigma=0.02, bias=0.0):
    out = np.empty(k, dtype=np.float32)
    lib().orc_synth_row(seed, tensor_id, row, k, sigma, bias, _p(out))
    return out


class Model:
    """Oracle Gemma forward on synthetic weights."""

    def __init__(self, n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx,
                 wtype=TYPE_Q4_0, seed=1234, eps=1e-6, rope_base=10000.0, n_threads=None,
                 sigma_embd=0.02, sigma_mat=0.02):
        self.cfg = Config(n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx,
                          eps, rope_base, wtype, 0, seed, sigma_embd, sigma_mat)
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self.h = lib().orc_model_create(C.byref(self.cfg), self.n_threads)

    def close(self):
        if self.h:
            lib().orc_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tensor(self, tid):
        rows, cols, nbytes = C.c_int64(), C.c_int64(), C.c_int64()
        p = lib().orc_model_tensor(self.h, tid, C.byref(rows), C.byref(cols), C.byref(nbytes))
        raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes.value,))
        return raw, rows.value, cols.value

    def reset(self):
        lib().orc_model_reset(self.h)

    def fill_kv_synth(self, seed=99):
        lib().orc_model_fill_kv_synth(self.h, seed)

    def forward(self, tokens, pos0, order=REF, logits_all=False):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        logits = np.empty(self.cfg.n_vocab, dtype=np.float32)
        rc = lib().orc_model_forward(self.h, _p(tokens), tokens.size, pos0, order, int(logits_all), _p(logits))
        if rc:
            raise RuntimeError(f"orc_model_forward rc={rc}")
        return logits

    def generate(self, prompt, n_decode, order=REF):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_decode, dtype=np.int32)
        ps, ds = C.c_double(), C.c_double()
        rc = lib().orc_model_generate(self.h, _p(prompt), prompt.size, n_decode, order, _p(out),
                                      C.byref(ps), C.byref(ds))
        if rc:
            raise RuntimeError(f"orc_model_generate rc={rc}")
        return out, ps.value, ds.value

    def capture(self, name, layer=0):
        n = C.c_int64()
        p = lib().orc_model_capture(self.h, name.encode(), layer, C.byref(n))
        if not p:
            raise KeyError(name)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n.value,)).copy()

    def kv(self, layer):
        n = self.cfg.n_ctx * self.cfg.n_head_kv * self.cfg.head_dim
        shape = (self.cfg.n_ctx, self.cfg.n_head_kv * self.cfg.head_dim)
        k = np.ctypeslib.as_array(C.cast(lib().orc_model_kcache(self.h, layer), C.POINTER(C.c_uint16)), shape=(n,))
        v = np.ctypeslib.as_array(C.cast(lib().orc_model_vcache(self.h, layer), C.POINTER(C.c_uint16)), shape=(n,))
        return k.reshape(shape), v.reshape(shape)


def tensor_id(layer, which):
    return 2 + 9 * layer + which


# ---- K-quants -------------------------------------------------------------
TYPE_F16, TYPE_Q4_K, TYPE_Q6_K, TYPE_Q8_K = 1, 12, 14, 15
KQ_BLOCK = {TYPE_Q4_K: 144, TYPE_Q6_K: 210, TYPE_Q8_K: 292}


def random_kquant_rows(rng, wtype, rows, k):
    """valid random Q4_K / Q6_K rows: every bit pattern is a block; the fp16 super-scales are kept moderate"""
    nb, size = k // 256, KQ_BLOCK[wtype]
    b = rng.integers(0, 256, size=(rows, nb, size), dtype=np.uint8)
    d = (rng.random((rows, nb)) * 0.01 + 0.001).astype(np.float16).view(np.uint16)
    if wtype == TYPE_Q4_K:
        b[:, :, 0:2] = d.view(np.uint8).reshape(rows, nb, 2)
        b[:, :, 2:4] = (d // 2 * 2).view(np.uint8).reshape(rows, nb, 2)
    else:
        b[:, :, 208:210] = d.view(np.uint8).reshape(rows, nb, 2)
    return np.ascontiguousarray(b.reshape(rows, nb * size))


def quantize_q8_K(x):
    """float32 [T, k] -> uint8 [T, k/256*292] block_q8_K rows ([UPSTREAM] quantize_row_q8_K_reference)"""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    out = np.zeros((x.shape[0], x.shape[1] // 256 * 292), dtype=np.uint8)
    for r in range(x.shape[0]):
        L.orc_quantize_row_q8_K(_p(x[r]), _p(out[r]), x.shape[1])
    return out


def mul_mat_blocks(wtype, w, act, k):
    """w uint8 [N, row bytes] (Q4_K / Q6_K / F16), act uint8 [T, row bytes] (Q8_K / F16) -> float32 [T, N],
    each element one scalar vec_dot (the loop of ref: src/hpc.cpp:15-41)"""
    L = lib()
    fn = {TYPE_Q4_K: L.orc_vec_dot_q4_K_q8_K, TYPE_Q6_K: L.orc_vec_dot_q6_K_q8_K, TYPE_F16: L.orc_vec_dot_f16}[wtype]
    fn.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    fn.restype = None
    out = np.zeros((act.shape[0], w.shape[0]), dtype=np.float32)
    s = C.c_float(0)
    for t in range(act.shape[0]):
        for r in range(w.shape[0]):
            fn(k, C.byref(s), _p(w[r]), _p(act[t]))
            out[t, r] = s.value
    return out
