"""tensor_dump mirror (host/tensor_dump.{h,cpp}; ref: src/tensor_dump.{h,cpp}): file naming, byte compare, fp32 report."""
import ctypes as C
import os

import numpy as np

import gemma_ggml_b200 as G


def test_dump_compare_and_report(tmp_path):
    H = G.host_lib()
    H.td_set_dir(str(tmp_path).encode())
    rng = np.random.default_rng(5)
    a = rng.standard_normal(1000).astype(np.float32)
    b = a.copy()
    assert H.td_dump_ptr_data(b"l_out-3", a.ctypes.data, a.nbytes, 0) == 0
    assert H.td_dump_ptr_data(b"l_out-3", b.ctypes.data, b.nbytes, 1) == 0
    # the reference's naming: <dir>/<name>_source and <name>_target, raw bytes
    assert (tmp_path / "l_out-3_source").read_bytes() == a.tobytes()
    assert (tmp_path / "l_out-3_target").read_bytes() == b.tobytes()
    assert H.td_compare_tensors(b"l_out-3") == 1
    rep = (C.c_double * 5)()
    assert H.td_compare_tensors_report(b"l_out-3", rep) == 0
    assert list(rep)[:4] == [1000.0, 0.0, 0.0, 0.0]
    b[17] += 0.25
    b[400] = np.nextafter(b[400], np.float32(10))
    assert H.td_dump_ptr_data(b"l_out-3", b.ctypes.data, b.nbytes, 1) == 0
    assert H.td_compare_tensors(b"l_out-3") == 0
    assert H.td_compare_tensors_report(b"l_out-3", rep) == 0
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    assert rep[0] == 1000 and rep[1] == 2
    assert rep[2] == d.max()
    assert np.isclose(rep[3], (d / np.maximum(np.abs(b.astype(np.float64)), 1e-30)).max())
    assert rep[4] == np.abs(b).max()
    # missing file / size mismatch
    assert H.td_compare_tensors(b"nope") == 0 and H.td_compare_tensors_report(b"nope", rep) != 0
    c = a[:10].copy()
    assert H.td_dump_ptr_data(b"l_out-3", c.ctypes.data, c.nbytes, 0) == 0
    assert H.td_compare_tensors(b"l_out-3") == 0 and H.td_compare_tensors_report(b"l_out-3", rep) != 0
