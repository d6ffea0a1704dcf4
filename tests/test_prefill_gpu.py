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
