"""Per-layer parity through the tensor_dump harness (ref: src/tensor_dump.cpp:11-89; BASELINE.json: "max-abs/rel error
reported per layer via tensor_dump"): the GPU's named intermediates are dumped as <name>-<layer>_source, the oracle's as
_target and must compare byte-identical (the reference's compare_tensors): the GPU path reproduces the scalar ggml order."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

DIMS = dict(n_embd=256, n_layer=3, n_head=4, n_head_kv=1, head_dim=64, n_ff=1024, n_vocab=2048, n_ctx=96)
CAPS = ("kqv_out", "sa_out", "ffn_act", "l_out")
REF_TOL = 2e-2  # of the tensor's max magnitude (set by Q8_0 re-quantisation of activations, see test_model_gpu.py)


@pytest.mark.parametrize("wtype", [O.TYPE_Q4_0, O.TYPE_Q8_0])
def test_per_layer_parity_via_tensor_dump(gpu, tmp_path, wtype):
    H = gpu.host_lib()
    H.td_set_dir(str(tmp_path).encode())
    cfg = gpu.make_config(wtype=wtype, sigma_embd=0.002, **DIMS)
    m = gpu.GemmaModel(cfg)
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=True, fused=True)
    om = O.Model(wtype=wtype, sigma_embd=0.002, **DIMS)
    prompt = ((1 + 7919 * np.arange(10)) % DIMS["n_vocab"]).astype(np.int32)
    toks, _, _ = m.generate(prompt, 5)
    sizes = {"kqv_out": DIMS["n_head"] * DIMS["head_dim"], "sa_out": DIMS["n_embd"], "ffn_act": DIMS["n_ff"], "l_out": DIMS["n_embd"]}
    for il in range(DIMS["n_layer"]):
        for cap in CAPS:
            gpu.check(H.td_dump_capture(m.h, cap.encode(), il, sizes[cap], 0), "td_dump_capture")
    rep = (C.c_double * 5)()
    report = []
    for order, name in ((O.REF, "REF"),):
        otoks, _, _ = om.generate(prompt, 5, order)
        assert np.array_equal(toks, otoks), name
        for il in range(DIMS["n_layer"]):
            for cap in CAPS:
                t = np.ascontiguousarray(om.capture(cap, il), dtype=np.float32)
                key = f"{cap}-{il}".encode()
                assert H.td_dump_ptr_data(key, t.ctypes.data, t.nbytes, 1) == 0
                assert H.td_compare_tensors_report(key, rep) == 0
                report.append(dict(order=name, tensor=key.decode(), n=int(rep[0]), n_diff=int(rep[1]), max_abs=rep[2],
                                   max_rel=rep[3], max_target=rep[4]))
                if order == O.REF:
                    assert H.td_compare_tensors(key) == 1, key   # the reference's byte compare
                else:
                    assert rep[2] <= REF_TOL * max(1.0, rep[4]), (key, list(rep))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(report, open(os.path.join(out_dir, f"tensor_dump_report_small_{'q4_0' if wtype == O.TYPE_Q4_0 else 'q8_0'}.json"), "w"), indent=0)
    m.close(); om.close()
