"""Writes a small synthetic Gemma GGUF (Q4_0 / Q8_0 / K-quant matrices from the oracle's generator, f32 norms, a toy vocabulary)
with gguf-py's GGUFWriter - the file format the reference loads (ref: src/gemma_model.cpp:19-51)."""
import numpy as np

import oracle_lib as O

TOKENS_HEAD = ["<pad>", "<eos>", "<bos>", "<unk>", "▁hello", "▁world", "!", "▁the"]


def write_gemma_gguf(path, om, dims, wtype, eos_id=1, bos_id=2):
    import gguf
    w = gguf.GGUFWriter(path, "gemma")
    w.add_uint32("gemma.block_count", dims["n_layer"])
    w.add_uint32("gemma.embedding_length", dims["n_embd"])
    w.add_uint32("gemma.attention.head_count", dims["n_head"])
    w.add_uint32("gemma.attention.head_count_kv", dims["n_head_kv"])
    w.add_uint32("gemma.attention.key_length", dims["head_dim"])
    w.add_uint32("gemma.feed_forward_length", dims["n_ff"])
    w.add_uint32("gemma.context_length", dims["n_ctx"])
    w.add_float32("gemma.attention.layer_norm_rms_epsilon", 1e-6)
    toks = TOKENS_HEAD + ["▁t%d" % i for i in range(len(TOKENS_HEAD), dims["n_vocab"])]
    w.add_token_list(toks)
    w.add_token_scores([0.0] * len(toks))
    w.add_token_types([1] * len(toks))
    w.add_bos_token_id(bos_id)
    w.add_eos_token_id(eos_id)
    QT = gguf.GGMLQuantizationType
    # (Q4_0 and Q4_K are both 4.5 bits per weight: the family comes from the model's wtype, the member from the block size)
    block = {144: (QT.Q4_K, 256), 210: (QT.Q6_K, 256)} if wtype in (O.TYPE_Q4_K, O.TYPE_Q6_K) else {18: (QT.Q4_0, 32), 34: (QT.Q8_0, 32)}
    names = ["token_embd.weight", "output_norm.weight"]
    per = ["attn_norm", "attn_q", "attn_k", "attn_v", "attn_output", "ffn_norm", "ffn_gate", "ffn_up", "ffn_down"]
    for il in range(dims["n_layer"]):
        names += ["blk.%d.%s.weight" % (il, p) for p in per]
    for tid, name in enumerate(names):
        raw, rows, cols = om.tensor(tid)
        if "norm" in name:
            w.add_tensor(name, raw.view(np.float32).copy())
        else:
            # the ggml type of THIS tensor from its block size (a K-quant model mixes Q4_K and Q6_K tensors)
            qt = next(t for bs, (t, per) in block.items() if raw.nbytes == rows * (cols // per) * bs)
            w.add_tensor(name, raw.reshape(rows, -1).copy(), raw_dtype=qt)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return toks
