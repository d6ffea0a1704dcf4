"""Per-layer parity report at FULL model size through the tensor_dump harness (script, not a test: it uses the CPU
oracle, so it lives under tests/).  Gemma-2B, a short prompt decoded token by token on the GPU (persistent kernel) and by
the oracle (scalar ggml order); every named intermediate of the LAST step is dumped as <name>-<layer>_source /
_target and compared.  Writes gpurun_out/tensor_dump_report_<model>_<wtype>.json; copy it to profiles/.

usage: python tests/tensor_dump_report.py [2b|7b] [q4_0|q8_0] [n_prompt] [n_decode]"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gemma_ggml_b200 as G  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "2b"
    wname = sys.argv[2] if len(sys.argv) > 2 else "q4_0"
    n_prompt = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    n_decode = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    wtype = G.TYPE_Q4_0 if wname == "q4_0" else G.TYPE_Q8_0
    dims = dict(getattr(G, {"2b": "GEMMA_2B", "7b": "GEMMA_7B"}[model]), n_ctx=64)
    G.init(0)
    H = G.host_lib()
    tmp = tempfile.mkdtemp(prefix="tensor_dump_")
    H.td_set_dir(tmp.encode())
    m = G.GemmaModel(G.make_config(wtype=wtype, **dims))
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(True, True)
    prompt = ((1 + 7919 * np.arange(n_prompt)) % dims["n_vocab"]).astype(np.int32)
    toks, _, _ = m.generate(prompt, n_decode)
    sizes = {"kqv_out": dims["n_head"] * dims["head_dim"], "sa_out": dims["n_embd"], "ffn_act": dims["n_ff"], "l_out": dims["n_embd"]}
    for il in range(dims["n_layer"]):
        for cap, n in sizes.items():
            G.check(H.td_dump_capture(m.h, cap.encode(), il, n, 0), "td_dump_capture")
    glog = m.logits()
    H.td_dump_ptr_data(b"result_output", glog.ctypes.data, glog.nbytes, 0)
    om = O.Model(wtype=wtype, **dims)
    rep = (C.c_double * 5)()
    rows, summary = [], {}
    for order, oname in ((O.REF, "REF"),):  # one order since round 2: the scalar ggml order
        otoks, _, _ = om.generate(prompt, n_decode, order)
        names = [(f"{cap}-{il}", om.capture(cap, il)) for il in range(dims["n_layer"]) for cap in sizes]
        names.append(("result_output", om.capture("result_output")))
        worst_abs = worst_rel_scale = 0.0
        identical = True
        for key, t in names:
            t = np.ascontiguousarray(t, dtype=np.float32)
            H.td_dump_ptr_data(key.encode(), t.ctypes.data, t.nbytes, 1)
            assert H.td_compare_tensors_report(key.encode(), rep) == 0, key
            same = bool(H.td_compare_tensors(key.encode()))
            identical &= same
            rows.append(dict(order=oname, tensor=key, n=int(rep[0]), n_diff=int(rep[1]), byte_identical=same,
                             max_abs=rep[2], max_rel=rep[3], max_target=rep[4]))
            worst_abs = max(worst_abs, rep[2])
            worst_rel_scale = max(worst_rel_scale, rep[2] / max(1.0, rep[4]))
        summary[oname] = dict(tokens_identical=bool(np.array_equal(toks, otoks)), all_tensors_byte_identical=identical,
                              worst_max_abs=worst_abs, worst_max_abs_over_tensor_max=worst_rel_scale)
    out = dict(model=model, wtype=wname, n_prompt=n_prompt, n_decode=n_decode, tokens=[int(x) for x in toks],
               exp_computed=bool(m.exp_computed), summary=summary, tensors=rows)
    path = os.path.join(ROOT, "gpurun_out", f"tensor_dump_report_{model}_{wname}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=0)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
