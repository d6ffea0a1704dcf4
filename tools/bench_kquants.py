"""Device-side throughput of the K-quant mat-vecs (csrc/kquants.cu): weights resident, one activation row."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gemma_ggml_b200 as G
import oracle_lib as O

G.init(0)
L = G.cuda_lib()
rng = np.random.default_rng(0)
for wtype, name in ((O.TYPE_Q4_K, "q4_K"), (O.TYPE_Q6_K, "q6_K"), (O.TYPE_F16, "f16")):
    for n, k in ((16384, 2048), (2048, 16384), (256000, 2048)):
        if wtype == O.TYPE_F16:
            w = rng.standard_normal((n, k)).astype(np.float16).view(np.uint8).reshape(n, 2 * k)
            act = rng.standard_normal((1, k)).astype(np.float16).view(np.uint8).reshape(1, 2 * k)
        else:
            w = O.random_kquant_rows(rng, wtype, n, k)
            act = O.quantize_q8_K(rng.standard_normal((1, k)))
        wd, ad = G.DeviceBuffer.from_numpy(w), G.DeviceBuffer.from_numpy(act)
        dst = G.DeviceBuffer(4 * n)
        for rep in range(2):
            G.sync(); t0 = time.perf_counter()
            for _ in range(50):
                G.check(L.ggb_op_mul_mat_blocks(wtype, wd.ptr, n, w.shape[1], k, ad.ptr, act.shape[1], 1, dst.ptr, n), "mm")
            G.sync(); dt = (time.perf_counter() - t0) / 50
        print(f"{name} {n}x{k}: {dt * 1e6:8.1f} us  {w.nbytes / dt / 1e9:7.0f} GB/s")
        if wtype == O.TYPE_Q4_K:
            packed = G.DeviceBuffer(w.nbytes)
            G.check(L.ggb_op_q4k_repack(wd.ptr, n, w.shape[1], k, packed.ptr), "repack")
            for rep in range(2):
                G.sync(); t0 = time.perf_counter()
                for _ in range(50):
                    G.check(L.ggb_op_mul_mat_q4k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], 1, dst.ptr, n), "mm")
                G.sync(); dt = (time.perf_counter() - t0) / 50
            print(f"{name} packed {n}x{k}: {dt * 1e6:8.1f} us  {w.nbytes / dt / 1e9:7.0f} GB/s")
        if wtype == O.TYPE_Q6_K:
            packed = G.DeviceBuffer(L.ggb_q6k_packed_bytes(n, k))
            G.check(L.ggb_op_q6k_repack(wd.ptr, n, w.shape[1], k, packed.ptr), "repack")
            for rep in range(2):
                G.sync(); t0 = time.perf_counter()
                for _ in range(50):
                    G.check(L.ggb_op_mul_mat_q6k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], 1, dst.ptr, n), "mm")
                G.sync(); dt = (time.perf_counter() - t0) / 50
            print(f"{name} packed {n}x{k}: {dt * 1e6:8.1f} us  {w.nbytes / dt / 1e9:7.0f} GB/s (algorithmic 210 B / super-block; the packed copy streams 274)")
