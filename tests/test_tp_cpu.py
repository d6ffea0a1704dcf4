"""CPU tests of the multi-GPU (row-sharded tensor-parallel) host logic: the shard plan covers every row
exactly once and respects head / Q8_0-block boundaries, and a world_size-2 gloo run ships the communicator
id and derives consistent plans on both ranks.  No GPU."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import gemma_ggml_b200 as G

MODELS = {"2b": G.GEMMA_2B, "7b": G.GEMMA_7B}


@pytest.mark.parametrize("model", ["2b", "7b"])
@pytest.mark.parametrize("size", [1, 2, 4, 8])
def test_plan_partitions_every_matrix(model, size):
    dims = MODELS[model]
    cfg = G.make_config(n_ctx=2048, **dims)
    plans = [G.tp_plan(cfg, r, size) for r in range(size)]
    assert all(p is not None for p in plans)
    d, H, Hkv = dims["head_dim"], dims["n_head"], dims["n_head_kv"]
    totals = {"q": H * d, "o": dims["n_embd"], "gate_up": dims["n_ff"], "down": dims["n_embd"], "logits": dims["n_vocab"]}
    for name, total in totals.items():
        covered = np.zeros(total, dtype=np.int32)
        for p in plans:
            covered[p[f"{name}_row0"]:p[f"{name}_row0"] + p[f"{name}_rows"]] += 1
        assert (covered == 1).all(), name  # every output row is computed by exactly one rank
    for r, p in enumerate(plans):
        assert p["q_row0"] == p["head0"] * d and p["q_rows"] == p["heads"] * d  # whole heads
        assert p["gate_up_row0"] % 32 == 0 and p["gate_up_rows"] % 32 == 0      # whole Q8_0 blocks
        if Hkv % size == 0:
            assert not p["kv_replicated"] and p["k_rows"] == Hkv // size * d and p["kv_head0"] == r * (Hkv // size)
            # a rank's query heads use only KV heads it holds
            group = H // Hkv
            assert p["head0"] // group >= p["kv_head0"]
            assert (p["head0"] + p["heads"] - 1) // group < p["kv_head0"] + p["kv_heads"]
        else:
            assert p["kv_replicated"] and p["k_row0"] == 0 and p["k_rows"] == Hkv * d


def test_plan_rejects_uneven_shapes():
    cfg = G.make_config(n_ctx=64, **dict(G.GEMMA_2B, n_head=6))
    assert G.tp_plan(cfg, 0, 4) is None
    cfg = G.make_config(n_ctx=64, **dict(G.GEMMA_7B, n_ff=24576 + 32))
    assert G.tp_plan(cfg, 0, 8) is None
    cfg = G.make_config(n_ctx=64, **G.GEMMA_2B)
    assert G.tp_plan(cfg, 2, 2) is None  # rank out of range


def test_gathered_shards_reassemble_the_single_rank_vector():
    """numpy model of the data path: each rank computes its rows, allgather concatenates by rank."""
    rng = np.random.default_rng(0)
    cfg = G.make_config(n_ctx=64, **G.GEMMA_7B)
    full = rng.standard_normal(G.GEMMA_7B["n_ff"]).astype(np.float32)
    for size in (2, 4, 8):
        parts = []
        for r in range(size):
            p = G.tp_plan(cfg, r, size)
            parts.append(full[p["gate_up_row0"]:p["gate_up_row0"] + p["gate_up_rows"]])
        assert np.array_equal(np.concatenate(parts), full)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import gemma_ggml_b200 as G
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    # ship a 128-byte communicator id from rank 0 exactly as GemmaModel.init_tensor_parallel does
    ident = np.arange(128, dtype=np.uint8) if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(ident)
    dist.broadcast(t, src=0)
    assert np.array_equal(t.numpy(), np.arange(128, dtype=np.uint8))
    cfg = G.make_config(n_ctx=2048, **G.GEMMA_7B)
    p = G.tp_plan(cfg, rank, 2)
    rows = torch.tensor([p["q_rows"], p["o_rows"], p["gate_up_rows"], p["logits_rows"], p["logits_row0"]])
    both = [torch.zeros_like(rows) for _ in range(2)]
    dist.all_gather(both, rows)
    assert int(both[0][0] + both[1][0]) == 16 * 256 and int(both[0][3] + both[1][3]) == 256000
    assert int(both[1][4]) == int(both[0][3])  # rank 1's logits start where rank 0's end
    # the 8-byte arg-max candidates: first maximum over ranks, lowest index on ties
    cand = torch.tensor([3.5, float(100 + rank)]) if rank == 0 else torch.tensor([3.5, 128000.0 + 7])
    cs = [torch.zeros(2) for _ in range(2)]
    dist.all_gather(cs, cand)
    best = max(cs, key=lambda c: (float(c[0]), -float(c[1])))
    assert int(best[1]) == 100
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_world_size_2_gloo_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    port = 29000 + (os.getpid() % 2000)
    script.write_text(WORKER.format(root=G.ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs), outs
