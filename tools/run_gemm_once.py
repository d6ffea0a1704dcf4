"""One prefill GEMM shape a few times (the ncu target for the tcgen05 kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gemma_ggml_b200 as G
G.init(0)
L = G.cuda_lib()
T, N, K = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
W = G.Weight.synth(G.TYPE_Q8_0, N, K, 1234, 7, 0.02)
A = G.DeviceBuffer(T * K * 2); A.zero()
Cb = G.DeviceBuffer(T * N * 4)
for _ in range(3):
    G.check(L.ggb_op_gemm_bf16(W.h, A.ptr, T, Cb.ptr, N, None, None, 0))
G.sync()
print("ok")
