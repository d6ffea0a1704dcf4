"""Worker of the multi-GPU parity test (one process per GPU): Gemma decode, tensor-parallel over `world` ranks, must
reproduce the oracle's results bit for bit (each output row is computed by exactly one rank in the same K-order as on
one GPU).  check() is also what bench.py runs before timing a multi-GPU line (its "tp_parity" key)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # MHA: attention sharded by KV head when the head counts divide the world size, replicated otherwise
    "mha": dict(n_embd=512, n_layer=2, n_head=4, n_head_kv=4, head_dim=128, n_ff=4608, n_vocab=1024, n_ctx=64),
    # MQA: attention block replicated, FFN / logits row-sharded
    "mqa": dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=96),
    # multi-segment DOWN with fewer row pairs per rank than CTAs: CTAs without rows in the row-lane chain
    "thin": dict(n_embd=256, n_layer=2, n_head=4, n_head_kv=1, head_dim=64, n_ff=4096, n_vocab=1024, n_ctx=48),
}


def check(G, O, dist, rank, world, lowerings=(True, False)):
    """fused = the persistent kernel with flagged peer stores over NVLink (no NCCL on the token path);
    not fused = per-operator kernels + NCCL allgathers at the layer boundaries.  Returns the number of cases checked."""
    n = 0
    G.init_comm(dist, rank, world)
    for fused in lowerings:
        os.environ["GGB_TP_FUSED"] = "1" if fused else "0"
        for name, dims in CONFIGS.items():
            for wtype in (G.TYPE_Q4_0, G.TYPE_Q8_0):
                cfg = G.make_config(wtype=wtype, sigma_embd=0.002, **dims)
                m = G.GemmaModel(cfg)
                if m.H.gm_init_tensor_parallel(m.h, rank, world) != 0:
                    # the per-operator lowering shards attention by heads: world must divide the head count
                    assert not fused, (name, world)
                    m.close()
                    continue
                m.init_synthetic_weights()
                m.init_kv_cache()
                m.build_compute_graph(capture_acts=True)
                assert m.is_fused == fused
                prompt = ((1 + 7919 * np.arange(35)) % dims["n_vocab"]).astype(np.int32)  # crosses a 32-position chunk
                toks, _, _ = m.generate(prompt, 7)
                logits = m.logits()  # collective (gathers the shards)
                if rank == 0:
                    om = O.Model(wtype=wtype, sigma_embd=0.002, **dims)
                    otoks, _, _ = om.generate(prompt, 7, O.REF)
                    assert np.array_equal(toks, otoks), (fused, name, wtype, toks, otoks)
                    ol = om.capture("result_output")
                    assert np.array_equal(logits.view(np.uint32), ol.view(np.uint32)), (fused, name, wtype, np.abs(logits - ol).max())
                    if not fused:  # the fused kernel dumps only the rows a rank computes itself
                        for il in range(dims["n_layer"]):
                            for cap in ("kqv_out", "sa_out", "ffn_act", "l_out"):
                                a, b = m.capture(cap, il), om.capture(cap, il)
                                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, wtype, cap, il)
                    om.close()
                dist.barrier()
                m.close()
                n += 1
    os.environ["GGB_TP_FUSED"] = "1"
    return n


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch
    import torch.distributed as dist
    import gemma_ggml_b200 as G
    import oracle_lib as O

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    G.init(rank)
    check(G, O, dist, rank, world)
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
