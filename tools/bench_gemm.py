"""Per-shape timing of the prefill GEMM (ggb_op_gemm_bf16) at T = 512: TFLOP/s per Gemma-2B matrix shape."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemma_ggml_b200 as G

G.init(0)
L = G.cuda_lib()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 512
wtype = G.TYPE_Q8_0 if (len(sys.argv) < 3 or sys.argv[2] == "q8") else G.TYPE_Q4_0
shapes = [("q/o 2048->2048", 2048, 2048), ("k/v 2048->256", 256, 2048), ("gate/up 2048->16384", 16384, 2048),
          ("down 16384->2048", 2048, 16384), ("logits 2048->256000", 256000, 2048)]
for name, N, K in shapes:
    W = G.Weight.synth(wtype, N, K, 1234, 7, 0.02)
    A = G.DeviceBuffer(T * K * 2); A.zero()
    Cb = G.DeviceBuffer(T * N * 4)
    wsb = L.ggb_gemm_workspace_bytes(T, N, K) if N <= 2048 else 0
    WS = G.DeviceBuffer(max(wsb, 16))
    for _ in range(3):
        G.check(L.ggb_op_gemm_bf16(W.h, A.ptr, T, Cb.ptr, N, None, WS.ptr if wsb else None, wsb))
    G.sync()
    L.ggb_prof_enable(1); G.prof_reset()
    reps = 10
    L.ggb_prof_start(b"g")
    for _ in range(reps):
        G.check(L.ggb_op_gemm_bf16(W.h, A.ptr, T, Cb.ptr, N, None, WS.ptr if wsb else None, wsb))
    L.ggb_prof_stop(b"g")
    ms, _ = G.prof_get("g")
    L.ggb_prof_enable(0)
    fl = 2.0 * T * N * K
    print(f"{name:24s} T={T}: {ms / reps * 1e3:9.1f} us  {fl / (ms / reps / 1e3) / 1e12:7.1f} TFLOP/s")
    del W, A, Cb, WS
