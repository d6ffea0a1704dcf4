"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/ggml_b200.h declares, fails loudly (no fallback) when there is no device, and the host
library exports the reference's `mul_mat` entry point."""
import ctypes as C
import os
import re
import subprocess

import pytest

import gemma_ggml_b200 as G


def _declared_symbols():
    src = open(G.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ggb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = G.cuda_lib()
    names = _declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ggml_b200.h but not exported: {missing}"


def test_ctypes_signatures_cover_the_header():
    assert sorted(G._CUDA_SIG) == _declared_symbols()


def test_host_library_exports_reference_entry_points():
    H = G.host_lib()
    for name in ("mul_mat", "init_cuda_offload", "gm_create", "gm_build_compute_graph", "gm_decode_step"):
        assert hasattr(H, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", G.HOST_LIB], capture_output=True, text=True).stdout
    assert re.search(r"\bT mul_mat\b", out), "mul_mat must have C linkage (ref: src/hpc.h:16-18)"


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    lib = G.cuda_lib()
    assert lib.ggb_init(0) != 0
    assert b"no CPU fallback" in lib.ggb_last_error()
    # every operator refuses to run uninitialised
    assert lib.ggb_op_quantize_q8_0(None, 32, 1, None) == -3
    assert lib.ggb_mul_mat(1, 1, 1, 34, 1, 4, 4, 34, 32, None, None, G.TYPE_Q8_0, None) == -3


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under the package or include/ may mention it
    pkg = os.path.dirname(G.CUDA_LIB)
    bad = []
    for root, _, files in os.walk(pkg):
        if os.path.basename(root) == "build":
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                txt = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"liboracle|oracle_lib|orc_[a-z]+\s*\(|#include\s+\"[^\"]*oracle", txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
    ldd = subprocess.run(["ldd", G.CUDA_LIB], capture_output=True, text=True).stdout + \
        subprocess.run(["ldd", G.HOST_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_config_struct_layout_matches_oracle():
    import oracle_lib as O
    assert C.sizeof(G.Config) == C.sizeof(O.Config)
    assert [f[0] for f in G.Config._fields_] == [f[0] for f in O.Config._fields_]


def test_mul_mat_without_a_device_kernel_runs_the_callers_vec_dot():
    """ref: src/hpc.cpp:15-41 — the reference computes rows with the vec_dot it is handed; the drop-in keeps that for
    src0 types that have no device kernel (here F32 x F32, no GPU needed: the type check precedes the device check)."""
    import ctypes as C
    import numpy as np
    host = G.host_lib()

    class Tensor(C.Structure):
        _fields_ = [("data", C.c_void_p)]

    VD = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int)
    calls = []

    def vec_dot(n, s, bs, x, bx, y, by, nrc):
        a = np.ctypeslib.as_array(C.cast(x, C.POINTER(C.c_float)), (n,))
        b = np.ctypeslib.as_array(C.cast(y, C.POINTER(C.c_float)), (n,))
        s[0] = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
        calls.append(n)

    rng = np.random.default_rng(0)
    n, k, t = 6, 40, 3
    w = rng.standard_normal((n, k)).astype(np.float32)
    a = rng.standard_normal((t, k)).astype(np.float32)
    dst = np.zeros((t, n), np.float32)
    src0, src1, d = Tensor(w.ctypes.data), Tensor(a.ctypes.data), Tensor(dst.ctypes.data)
    host.mul_mat.restype = None
    host.mul_mat.argtypes = [C.c_int64] * 7 + [C.c_size_t, C.c_int64, C.POINTER(Tensor), C.POINTER(Tensor),
                                               C.POINTER(Tensor), VD, C.c_int, C.c_char_p]
    cb = VD(vec_dot)
    host.mul_mat(n, t, 1, k * 4, t, n * 4, n * 4 * t, k * 4, k, C.byref(src0), C.byref(src1), C.byref(d), cb, 0,
                 C.cast(a.ctypes.data, C.c_char_p))
    assert len(calls) == n * t
    assert np.allclose(dst, a @ w.T, rtol=1e-6, atol=1e-6)
