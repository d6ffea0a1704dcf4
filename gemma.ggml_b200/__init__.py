"""gemma.ggml_b200 — B200-native (sm_100a) replacement of gemma.ggml's quantized decode hot path.

Python is only the test / bench driver here: the product is two native libraries built in-tree,

  libggb_cuda.so  the thin C-ABI CUDA layer declared in include/ggml_b200.h (replaces the
                  reference's src/opencl.cpp + src/kernals.cl)
  libggb_host.so  the host C++ side (rewired hpc.cpp dispatch + gemma_model graph builder)

This module binds them with ctypes.  It never falls back to a CPU implementation: if the libraries
are missing it raises, and ggb_init fails loudly when there is no sm_100 device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CUDA_LIB = os.path.join(HERE, "libggb_cuda.so")
HOST_LIB = os.path.join(HERE, "libggb_host.so")
HEADER = os.path.join(ROOT, "include", "ggml_b200.h")

TYPE_F32, TYPE_F16, TYPE_Q4_0, TYPE_Q8_0 = 0, 1, 2, 8
TYPE_Q4_K, TYPE_Q6_K, TYPE_Q8_K = 12, 14, 15
BLOCK_BYTES = {TYPE_Q4_0: 18, TYPE_Q8_0: 34}


class GgbError(RuntimeError):
    pass


class Config(C.Structure):
    """gemma_config (host/gemma_model.h); same layout as the oracle's orc_config."""
    _fields_ = [("n_embd", C.c_int32), ("n_layer", C.c_int32), ("n_head", C.c_int32),
                ("n_head_kv", C.c_int32), ("head_dim", C.c_int32), ("n_ff", C.c_int32),
                ("n_vocab", C.c_int32), ("n_ctx", C.c_int32), ("eps", C.c_float),
                ("rope_base", C.c_float), ("wtype", C.c_int32), ("pad_", C.c_int32),
                ("seed", C.c_uint64), ("sigma_embd", C.c_float), ("sigma_mat", C.c_float)]


GEMMA_2B = dict(n_embd=2048, n_layer=18, n_head=8, n_head_kv=1, head_dim=256, n_ff=16384, n_vocab=256000)
GEMMA_7B = dict(n_embd=3072, n_layer=28, n_head=16, n_head_kv=16, head_dim=256, n_ff=24576, n_vocab=256000)


def make_config(n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx, wtype=TYPE_Q4_0,
                seed=1234, eps=1e-6, rope_base=10000.0, sigma_embd=0.02, sigma_mat=0.02):
    return Config(n_embd, n_layer, n_head, n_head_kv, head_dim, n_ff, n_vocab, n_ctx, eps, rope_base,
                  wtype, 0, seed, sigma_embd, sigma_mat)


def build(force=False):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ggb_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)


_cuda = None
_host = None

vp, i32, i64, f32, u64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_size_t
P = C.POINTER

_CUDA_SIG = {
    "ggb_init": (i32, [i32]), "ggb_shutdown": (i32, []), "ggb_last_error": (C.c_char_p, []),
    "ggb_device_info": (i32, [C.c_char_p, sz]), "ggb_sm_count": (i32, []), "ggb_sync": (i32, []),
    "ggb_stream": (vp, []), "ggb_launch_count": (i64, []),
    "ggb_buffer_alloc": (i32, [sz, P(vp)]), "ggb_buffer_free": (i32, [vp]),
    "ggb_buffer_memset": (i32, [vp, i32, sz]), "ggb_upload": (i32, [vp, vp, sz]),
    "ggb_download": (i32, [vp, vp, sz]), "ggb_host_alloc": (i32, [sz, P(vp)]), "ggb_host_free": (i32, [vp]),
    "ggb_prof_enable": (i32, [i32]), "ggb_prof_start": (i32, [C.c_char_p]), "ggb_prof_stop": (i32, [C.c_char_p]),
    "ggb_prof_reset": (i32, []), "ggb_prof_get": (i32, [C.c_char_p, P(C.c_double), P(i64)]),
    "ggb_prof_report": (i32, [C.c_char_p, sz]),
    "ggb_capture_begin": (i32, []), "ggb_capture_end": (i32, [P(vp)]), "ggb_graph_launch": (i32, [vp]),
    "ggb_graph_free": (i32, [vp]), "ggb_graph_kernel_count": (i32, [vp]),
    "ggb_weight_upload": (i32, [i32, i64, i64, vp, i64, P(vp)]),
    "ggb_weight_synth": (i32, [i32, i64, i64, u64, u64, f32, P(vp)]),
    "ggb_weight_slice": (i32, [vp, i64, i64, P(vp)]), "ggb_weight_download": (i32, [vp, vp]),
    "ggb_weight_free": (i32, [vp]), "ggb_weight_rows": (i64, [vp]), "ggb_weight_cols": (i64, [vp]),
    "ggb_weight_type": (i32, [vp]), "ggb_weight_algorithmic_bytes": (sz, [vp]),
    "ggb_vector_synth": (i32, [vp, i64, u64, u64, f32, f32]),
    "ggb_mul_mat": (i32, [i64, i64, i64, i64, i64, i64, i64, sz, i64, vp, vp, i32, vp]),
    "ggb_mul_mat_forget": (i32, [vp]),
    "ggb_q8_row_bytes": (sz, [i64]),
    "ggb_op_quantize_q8_0": (i32, [vp, i64, i64, vp]), "ggb_op_q8_to_blocks": (i32, [vp, i64, i64, vp]),
    "ggb_op_q8_from_blocks": (i32, [vp, i64, i64, i64, vp]),
    "ggb_type_row_bytes": (sz, [i32, i64]), "ggb_op_quantize_q8_K": (i32, [vp, i64, i64, vp]),
    "ggb_op_mul_mat_blocks": (i32, [i32, vp, i64, i64, i64, vp, i64, i64, vp, i64]),
    "ggb_op_q4k_repack": (i32, [vp, i64, i64, i64, vp]),
    "ggb_op_mul_mat_q4k_packed": (i32, [vp, i64, i64, vp, i64, i64, vp, i64]),
    "ggb_op_mul_mat_kq_packed_f32": (i32, [i32, vp, i64, i64, vp, i64, vp, i64]),
    "ggb_op_get_rows_blocks": (i32, [i32, vp, i64, i64, i64, vp, i64, f32, vp]),
    "ggb_q6k_packed_bytes": (sz, [i64, i64]), "ggb_op_q6k_repack": (i32, [vp, i64, i64, i64, vp]),
    "ggb_op_mul_mat_q6k_packed": (i32, [vp, i64, i64, vp, i64, i64, vp, i64]),
    "ggb_op_mul_mat_q8": (i32, [vp, vp, i64, vp, i64, vp]),
    "ggb_op_mul_mat_q8_multi": (i32, [i32, P(vp), vp, P(vp)]),
    "ggb_op_ffn_gate_up": (i32, [vp, vp, vp, vp, vp]),
    "ggb_op_ffn_gate_up_shard": (i32, [vp, vp, vp, vp, vp, vp]),
    "ggb_op_argmax_shard": (i32, [vp, i64, i32, vp, vp]), "ggb_op_sample_from_candidates": (i32, [vp, i32, vp, vp]),
    "ggb_op_rms_norm": (i32, [vp, vp, f32, i64, i64, vp, vp]),
    "ggb_op_add": (i32, [vp, vp, i64, vp]), "ggb_op_mul": (i32, [vp, vp, i64, vp]),
    "ggb_op_scale": (i32, [vp, f32, i64, vp]), "ggb_op_gelu": (i32, [vp, i64, vp]),
    "ggb_op_soft_max_ext": (i32, [vp, vp, f32, i64, i64, vp]),
    "ggb_op_rope_neox": (i32, [vp, vp, i32, i32, i64, vp]),
    "ggb_op_get_rows": (i32, [vp, vp, i64, f32, vp]), "ggb_op_argmax": (i32, [vp, i64, vp]),
    "ggb_op_cpy_f32_f16": (i32, [vp, i64, vp]),
    "ggb_rope_table_init": (i32, [i32, i32, f32]),
    "ggb_op_rope_kv_store": (i32, [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp]),
    "ggb_op_rope_kv_store_batch": (i32, [vp, vp, vp, vp, i64, i32, i32, i32, f32, vp, vp]),
    "ggb_attn_scratch_bytes": (sz, [i32, i32, i32]),
    "ggb_op_attn_decode": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "ggb_op_embed_token": (i32, [vp, vp, f32, vp]),
    "ggb_op_sample_greedy": (i32, [vp, i64, vp, vp, vp]), "ggb_sample_scratch_bytes": (sz, [i64]),
    "ggb_gemm_workspace_bytes": (sz, [i64, i64, i64]),
    "ggb_op_gemm_bf16": (i32, [vp, vp, i64, vp, i64, vp, vp, sz]),
    "ggb_op_gemm_bf16_multi": (i32, [i32, vp, vp, i64, vp, vp, vp, sz]),
    "ggb_op_cvt_f32_bf16": (i32, [vp, i64, vp]),
    "ggb_op_rms_norm_bf16": (i32, [vp, vp, f32, i64, i64, vp]), "ggb_op_geglu_bf16": (i32, [vp, vp, i64, vp]),
    "ggb_op_attn_prefill": (i32, [vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]),
    "ggb_decoder_create": (i32, [vp, vp, vp, vp, vp, P(vp)]),
    "ggb_decoder_create_tp": (i32, [vp, vp, vp, vp, vp, i32, i32, P(vp)]), "ggb_decoder_step": (i32, [vp]),
    "ggb_decoder_kv_heads_local": (i32, [i32, i32, i32]),
    "ggb_decoder_stream_bytes": (C.c_longlong, [vp]), "ggb_decoder_kv_heads": (i32, [vp]),
    "ggb_decoder_free": (i32, [vp]), "ggb_decoder_trace": (i32, [vp, vp, vp, i32, P(i32), P(i32)]),
    "ggb_decoder_exp_computed": (i32, [vp]),
    "ggb_comm_unique_id": (i32, [vp]), "ggb_comm_init": (i32, [vp, i32, i32]), "ggb_comm_destroy": (i32, []),
    "ggb_comm_rank": (i32, []), "ggb_comm_size": (i32, []), "ggb_op_allgather": (i32, [vp, vp, sz]),
}

_HOST_SIG = {
    "gm_create": (vp, [P(Config)]), "gm_free": (None, [vp]),
    "gm_init_synthetic_weights": (i32, [vp]), "gm_init_tensor_parallel": (i32, [vp, i32, i32]),
    "gm_tp_plan": (i32, [P(Config), i32, i32, P(i64)]), "gm_set_tensor": (i32, [vp, i32, vp, i64]),
    "gm_set_tensor_typed": (i32, [vp, i32, i32, vp, i64]), "gm_kquant_mode": (i32, [vp]),
    "gm_init_kv_cache": (i32, [vp]), "gm_fill_kv_cache_synthetic": (i32, [vp, u64]),
    "gm_build_compute_graph": (i32, [vp, i32, i32]), "gm_decoder": (vp, [vp]), "gm_is_fused": (i32, [vp]), "gm_decode_step": (i32, [vp, i32, i32, P(i32)]),
    "gm_set_state": (i32, [vp, i32, i32, i32]), "gm_run_steps": (i32, [vp, i32]), "gm_run_steps_profiled": (i32, [vp, i32]),
    "gm_get_state": (i32, [vp, P(i32), P(i32)]),
    "gm_generate": (i32, [vp, vp, i32, i32, vp, P(i32), P(C.c_double), P(C.c_double)]),
    "gm_last_error": (C.c_char_p, []),
    "gm_gguf_probe": (i32, [C.c_char_p, vp, P(i32), P(i32), P(i32), P(i32), P(i32)]),
    "gm_gguf_detokenize": (i32, [C.c_char_p, vp, i32, C.c_char_p, i32]),
    "gm_load_model_from_file": (vp, [C.c_char_p, i32]),
    "gm_get_config": (i32, [vp, vp]),
    "gm_begin_one_round_inference": (i32, [vp, vp, i32, i32, vp, P(i32), C.c_char_p, i32, P(C.c_double), P(C.c_double)]),
    "gm_prefill": (i32, [vp, vp, i32, i32, P(i32)]), "gm_get_prefill_logits": (i32, [vp, i32, vp]),
    "gm_last_prefill_gemm_flops": (C.c_double, [vp]),
    "gm_get_logits": (i32, [vp, vp]), "gm_get_capture": (i32, [vp, C.c_char_p, i32, vp, i64]),
    "td_set_dir": (None, [C.c_char_p]), "td_dump_ptr_data": (i32, [C.c_char_p, vp, C.c_uint64, i32]),
    "td_dump_capture": (i32, [vp, C.c_char_p, i32, i64, i32]), "td_compare_tensors": (i32, [C.c_char_p]),
    "td_compare_tensors_report": (i32, [C.c_char_p, P(C.c_double)]),
    "gm_get_kv": (i32, [vp, i32, vp, vp]), "gm_kernels_per_step": (i32, [vp]), "gm_weight_bytes": (u64, [vp]),
    "mul_mat": (None, [i64, i64, i64, i64, i64, i64, i64, sz, i64, vp, vp, vp, vp, i32, vp]),
    "init_cuda_offload": (i32, [i32]),
}


def _bind(lib, sig):
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


def cuda_lib():
    """libggb_cuda.so (loads without a GPU; ggb_init is what needs one)."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB):
            raise GgbError(f"{CUDA_LIB} is not built (run __graft_entry__.build()); there is no fallback path")
        _cuda = C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL)
        _bind(_cuda, _CUDA_SIG)
    return _cuda


def host_lib():
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(HOST_LIB):
            raise GgbError(f"{HOST_LIB} is not built (run __graft_entry__.build()); there is no fallback path")
        _host = C.CDLL(HOST_LIB, mode=C.RTLD_GLOBAL)
        _bind(_host, _HOST_SIG)
    return _host


def check(rc, what=""):
    if rc != 0:
        raise GgbError(f"{what} failed (rc={rc}): {cuda_lib().ggb_last_error().decode(errors='replace')}")


_inited = False


def init(device=0):
    global _inited
    if not _inited:
        check(cuda_lib().ggb_init(device), "ggb_init")
        _inited = True


def sync():
    check(cuda_lib().ggb_sync(), "ggb_sync")


def _ptr(a):
    return a.ctypes.data_as(vp)


class DeviceBuffer:
    """A ggb_buffer_alloc'd region with numpy up/download helpers."""

    def __init__(self, nbytes):
        init()
        self.nbytes = int(nbytes)
        self.ptr = vp()
        check(cuda_lib().ggb_buffer_alloc(self.nbytes, C.byref(self.ptr)), "ggb_buffer_alloc")

    @classmethod
    def from_numpy(cls, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(arr.nbytes)
        b.upload(arr)
        return b

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(cuda_lib().ggb_upload(self.ptr, _ptr(arr), arr.nbytes), "ggb_upload")
        sync()

    def download(self, dtype, count=None):
        dtype = np.dtype(dtype)
        count = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(count, dtype=dtype)
        check(cuda_lib().ggb_download(_ptr(out), self.ptr, out.nbytes), "ggb_download")
        return out

    def zero(self):
        check(cuda_lib().ggb_buffer_memset(self.ptr, 0, self.nbytes), "memset")

    def free(self):
        if self.ptr:
            cuda_lib().ggb_buffer_free(self.ptr)
            self.ptr = vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Weight:
    """A resident quantised matrix (ggb_weight)."""

    def __init__(self, handle, owner=True):
        self.h = handle
        self.owner = owner

    @classmethod
    def upload(cls, wtype, blocks, cols):
        """blocks: uint8 [rows, row_bytes] of ggml AoS blocks."""
        init()
        blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
        h = vp()
        check(cuda_lib().ggb_weight_upload(wtype, blocks.shape[0], cols, _ptr(blocks), blocks.shape[1], C.byref(h)),
              "ggb_weight_upload")
        return cls(h)

    @classmethod
    def synth(cls, wtype, rows, cols, seed, tensor_id, sigma):
        init()
        h = vp()
        check(cuda_lib().ggb_weight_synth(wtype, rows, cols, seed, tensor_id, sigma, C.byref(h)), "ggb_weight_synth")
        return cls(h)

    def slice(self, row0, nrows):
        h = vp()
        check(cuda_lib().ggb_weight_slice(self.h, row0, nrows, C.byref(h)), "ggb_weight_slice")
        w = Weight(h)
        w._parent = self
        return w

    @property
    def rows(self):
        return int(cuda_lib().ggb_weight_rows(self.h))

    @property
    def cols(self):
        return int(cuda_lib().ggb_weight_cols(self.h))

    @property
    def wtype(self):
        return int(cuda_lib().ggb_weight_type(self.h))

    def download(self):
        rb = self.cols // 32 * BLOCK_BYTES[self.wtype]
        out = np.empty((self.rows, rb), dtype=np.uint8)
        check(cuda_lib().ggb_weight_download(self.h, _ptr(out)), "ggb_weight_download")
        return out

    def free(self):
        if self.h:
            cuda_lib().ggb_weight_free(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def prof_get(name):
    ms, calls = C.c_double(), i64()
    check(cuda_lib().ggb_prof_get(name.encode(), C.byref(ms), C.byref(calls)), "ggb_prof_get")
    return ms.value, calls.value


def prof_reset():
    check(cuda_lib().ggb_prof_reset(), "ggb_prof_reset")


TP_PLAN_FIELDS = ("valid", "head0", "heads", "kv_head0", "kv_heads", "kv_replicated", "q_row0", "q_rows", "k_row0",
                  "k_rows", "o_row0", "o_rows", "gate_up_row0", "gate_up_rows", "down_row0", "down_rows", "logits_row0",
                  "logits_rows")


def tp_plan(cfg, rank, size):
    """Host-only: the row ranges rank `rank` of `size` owns (host/gemma_model.cpp make_tp_plan); None if the
    shape does not shard evenly."""
    out = (i64 * 18)()
    rc = host_lib().gm_tp_plan(C.byref(cfg), rank, size, out)
    d = dict(zip(TP_PLAN_FIELDS, [int(v) for v in out]))
    return d if rc == 0 else None


def q8_row_bytes(cols):
    return int(cuda_lib().ggb_q8_row_bytes(cols))


def quantize_row_q8_0(x):
    """x: float32 [T, cols] -> (planar DeviceBuffer, uint8 [T, cols/32*34] ggml AoS blocks on host)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    T, cols = x.shape
    L = cuda_lib()
    xd = DeviceBuffer.from_numpy(x)
    q8 = DeviceBuffer(q8_row_bytes(cols) * T)
    blocks = DeviceBuffer(cols // 32 * 34 * T)
    check(L.ggb_op_quantize_q8_0(xd.ptr, cols, T, q8.ptr), "quantize")
    check(L.ggb_op_q8_to_blocks(q8.ptr, cols, T, blocks.ptr), "q8_to_blocks")
    sync()
    return q8, blocks.download(np.uint8).reshape(T, cols // 32 * 34)


def quantize_q8_K(x):
    """[UPSTREAM] quantize_row_q8_K on the device: float32 [T, cols] -> uint8 [T, cols/256*292] block_q8_K rows."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    T, cols = x.shape
    L = cuda_lib()
    xd = DeviceBuffer.from_numpy(x)
    out = DeviceBuffer(cols // 256 * 292 * T)
    check(L.ggb_op_quantize_q8_K(xd.ptr, cols, T, out.ptr), "quantize_q8_K")
    sync()
    return out.download(np.uint8).reshape(T, cols // 256 * 292)


def mul_mat_host(wtype, w_blocks, act_blocks, k):
    """The drop-in entry with HOST pointers (ggb_mul_mat): returns float32 [T, N]."""
    init()
    w_blocks = np.ascontiguousarray(w_blocks, dtype=np.uint8)
    act_blocks = np.ascontiguousarray(act_blocks, dtype=np.uint8)
    n, t = w_blocks.shape[0], act_blocks.shape[0]
    dst = np.zeros((t, n), dtype=np.float32)
    check(cuda_lib().ggb_mul_mat(n, t, 1, w_blocks.shape[1], t, n * 4, n * 4 * t, act_blocks.shape[1], k,
                                 _ptr(w_blocks), _ptr(dst), wtype, _ptr(act_blocks)), "ggb_mul_mat")
    cuda_lib().ggb_mul_mat_forget(_ptr(w_blocks))
    return dst


def gguf_probe(path):
    """host-only parse of a Gemma GGUF: (Config, n_tensors, n_kv, n_tokens, bos, eos)"""
    H = host_lib()
    cfg = Config()
    v = [i32() for _ in range(5)]
    rc = H.gm_gguf_probe(os.fsencode(path), C.byref(cfg), *[C.byref(x) for x in v])
    if rc:
        raise GgbError("gguf_probe(%s): %s" % (path, H.gm_last_error().decode(errors="replace")))
    return (cfg,) + tuple(x.value for x in v)


def gguf_detokenize(path, ids):
    """host-only: the text of `ids` with the file's vocabulary (ref: print_tokens, src/gemma_model.cpp:762-772)"""
    H = host_lib()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    buf = C.create_string_buffer(64 * ids.size + 64)
    n = H.gm_gguf_detokenize(os.fsencode(path), _ptr(ids), ids.size, buf, len(buf))
    if n < 0:
        raise GgbError("gguf_detokenize: %s" % H.gm_last_error().decode(errors="replace"))
    return buf.value.decode("utf-8", errors="replace")


_comm_world = 0


def init_comm(dist, rank, world):
    """bring up the library's NCCL communicator (idempotent)"""
    global _comm_world
    if _comm_world == world:
        return
    import torch
    L = cuda_lib()
    ident = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        check(L.ggb_comm_unique_id(_ptr(ident)), "ggb_comm_unique_id")
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(ident).to(dev)
    dist.broadcast(t, src=0)
    ident = t.cpu().numpy().copy()
    check(L.ggb_comm_init(_ptr(ident), rank, world), "ggb_comm_init")
    _comm_world = world


class GemmaModel:
    """ctypes face of host/gemma_model.{h,cpp}."""

    def __init__(self, cfg, _handle=None):
        init()
        self.cfg = cfg
        self.H = host_lib()
        self.h = _handle if _handle is not None else self.H.gm_create(C.byref(cfg))

    @classmethod
    def load_model_from_file(cls, path, n_ctx=0):
        """ref: gemma_model::load_model_from_file (src/gemma_model.cpp:19-51): a Q4_0 / Q8_0 Gemma GGUF -> device."""
        init()
        H = host_lib()
        h = H.gm_load_model_from_file(os.fsencode(path), int(n_ctx))
        if not h:
            raise GgbError("load_model_from_file(%s): %s" % (path, H.gm_last_error().decode(errors="replace")))
        cfg = Config()
        H.gm_get_config(h, C.byref(cfg))
        return cls(cfg, _handle=h)

    def begin_one_round_inference(self, prompt, max_tokens):
        """ref: begin_one_round_inference (src/gemma_model.cpp:548-575): prefill, decode until EOS / max_tokens.
        Returns (new token ids, the text the reference would print, prefill ms, decode ms)."""
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(max_tokens + 2, dtype=np.int32)
        text = C.create_string_buffer(64 * (prompt.size + max_tokens + 2) + 64)
        n_out, pm, dm = i32(), C.c_double(), C.c_double()
        check(self.H.gm_begin_one_round_inference(self.h, _ptr(prompt), prompt.size, int(max_tokens), _ptr(out), C.byref(n_out),
                                                  text, len(text), C.byref(pm), C.byref(dm)), "gm_begin_one_round_inference")
        return out[:n_out.value].copy(), text.value.decode("utf-8", errors="replace"), pm.value, dm.value

    def close(self):
        if self.h:
            self.H.gm_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_tensor_parallel(self, dist, rank, world):
        """Tensor parallelism over `world` ranks (one process per GPU).  `dist` is an initialised torch.distributed
        module/backend used only to ship the 128-byte NCCL unique id from rank 0 (once per process)."""
        init_comm(dist, rank, world)
        check(self.H.gm_init_tensor_parallel(self.h, rank, world), "gm_init_tensor_parallel")

    def init_synthetic_weights(self):
        check(self.H.gm_init_synthetic_weights(self.h), "gm_init_synthetic_weights")

    def set_tensor(self, tensor_id, raw):
        raw = np.ascontiguousarray(raw)
        check(self.H.gm_set_tensor(self.h, tensor_id, _ptr(raw), raw.nbytes), f"gm_set_tensor({tensor_id})")

    def set_tensor_typed(self, tensor_id, wtype, raw):
        """a matrix in ggml type `wtype` (TYPE_Q4_K / TYPE_Q6_K put the model into K-quant mode), or an F32 norm"""
        raw = np.ascontiguousarray(raw)
        check(self.H.gm_set_tensor_typed(self.h, tensor_id, wtype, _ptr(raw), raw.nbytes), f"gm_set_tensor_typed({tensor_id})")

    @property
    def kquant_mode(self):
        return bool(self.H.gm_kquant_mode(self.h))

    def init_kv_cache(self):
        check(self.H.gm_init_kv_cache(self.h), "gm_init_kv_cache")

    def fill_kv_cache_synthetic(self, seed=99):
        check(self.H.gm_fill_kv_cache_synthetic(self.h, seed), "gm_fill_kv_cache_synthetic")

    def build_compute_graph(self, capture_acts=False, fused=True):
        """fused=True: the persistent decode kernel (falls back to the per-operator sequence when the
        shape is outside its envelope; see .is_fused); fused=False: one kernel per (group of) node(s)."""
        check(self.H.gm_build_compute_graph(self.h, int(capture_acts), int(fused)), "gm_build_compute_graph")

    def trace_step(self):
        """One traced fused step: (stamps[n_phases, grid, 4] in %globaltimer ns, kinds[n_phases])."""
        dec = self.H.gm_decoder(self.h)
        if not dec:
            raise GgbError("trace_step needs the fused decoder")
        max_ph = 4096
        grid = cuda_lib().ggb_sm_count()
        stamps = np.zeros(max_ph * grid * 4, dtype=np.int64)
        kinds = np.zeros(max_ph, dtype=np.int32)
        n, g = i32(), i32()
        check(cuda_lib().ggb_decoder_trace(dec, _ptr(stamps), _ptr(kinds), max_ph, C.byref(n), C.byref(g)),
              "ggb_decoder_trace")
        n1 = n.value * g.value * 4
        self.prof = stamps[n1:n1 + g.value * 64].reshape(g.value, 64).copy()  # GGB_PROF_PHASE sums (SM clocks)
        return stamps[:n1].reshape(n.value, g.value, 4), kinds[:n.value].copy()

    @property
    def rank_stream_bytes(self):
        """bytes of weight pass stream this rank reads per token (fused lowering)"""
        dec = self.H.gm_decoder(self.h)
        return int(cuda_lib().ggb_decoder_stream_bytes(dec)) if dec else 0

    @property
    def rank_kv_heads(self):
        dec = self.H.gm_decoder(self.h)
        return int(cuda_lib().ggb_decoder_kv_heads(dec)) if dec else 0

    @property
    def exp_computed(self):
        """True if the fused kernel computes exp (proved identical to the fp16 table at create time)."""
        dec = self.H.gm_decoder(self.h)
        return bool(dec and cuda_lib().ggb_decoder_exp_computed(dec))

    @property
    def is_fused(self):
        return bool(self.H.gm_is_fused(self.h))

    def prefill(self, tokens, pos0=0):
        """Batched prefill through the tcgen05 GEMM path; returns the greedy token of the last position."""
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        nxt = i32()
        check(self.H.gm_prefill(self.h, _ptr(tokens), tokens.size, int(pos0), C.byref(nxt)), "gm_prefill")
        return nxt.value

    def prefill_logits(self, row):
        out = np.empty(self.cfg.n_vocab, dtype=np.float32)
        check(self.H.gm_get_prefill_logits(self.h, int(row), _ptr(out)), "gm_get_prefill_logits")
        return out

    @property
    def prefill_gemm_flops(self):
        return float(self.H.gm_last_prefill_gemm_flops(self.h))

    def decode_step(self, token, pos):
        nxt = i32()
        check(self.H.gm_decode_step(self.h, int(token), int(pos), C.byref(nxt)), "gm_decode_step")
        return nxt.value

    def set_state(self, token, pos, pos_step=1):
        check(self.H.gm_set_state(self.h, int(token), int(pos), int(pos_step)), "gm_set_state")

    def run_steps(self, steps):
        check(self.H.gm_run_steps(self.h, int(steps)), "gm_run_steps")

    def run_steps_profiled(self, steps):
        check(self.H.gm_run_steps_profiled(self.h, int(steps)), "gm_run_steps_profiled")

    def get_state(self):
        t, p = i32(), i32()
        check(self.H.gm_get_state(self.h, C.byref(t), C.byref(p)), "gm_get_state")
        return t.value, p.value

    def generate(self, prompt, n_decode):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_decode + 1, dtype=np.int32)
        n_out, pm, dm = i32(), C.c_double(), C.c_double()
        check(self.H.gm_generate(self.h, _ptr(prompt), prompt.size, n_decode, _ptr(out), C.byref(n_out),
                                 C.byref(pm), C.byref(dm)), "gm_generate")
        return out[:n_out.value].copy(), pm.value, dm.value

    def logits(self):
        out = np.empty(self.cfg.n_vocab, dtype=np.float32)
        check(self.H.gm_get_logits(self.h, _ptr(out)), "gm_get_logits")
        return out

    def capture(self, name, layer=0):
        c = self.cfg
        n = {"l_out": c.n_embd, "sa_out": c.n_embd, "kqv_out": c.n_head * c.head_dim, "ffn_act": c.n_ff,
             "result_norm": c.n_embd, "result_output": c.n_vocab}[name]
        out = np.empty(n, dtype=np.float32)
        check(self.H.gm_get_capture(self.h, name.encode(), layer, _ptr(out), n), f"gm_get_capture({name})")
        return out

    def kv(self, layer):
        c = self.cfg
        shape = (c.n_ctx, c.n_head_kv * c.head_dim)
        k = np.empty(shape, dtype=np.uint16)
        v = np.empty(shape, dtype=np.uint16)
        check(self.H.gm_get_kv(self.h, layer, _ptr(k), _ptr(v)), "gm_get_kv")
        return k, v

    @property
    def kernels_per_step(self):
        return int(self.H.gm_kernels_per_step(self.h))

    @property
    def weight_bytes(self):
        return int(self.H.gm_weight_bytes(self.h))
