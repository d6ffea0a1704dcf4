"""Generate tests/golden/codecs_gguf.npz — known-answer vectors for the Q4_0 / Q8_0 codecs from
gguf-py 0.19.0 (gguf/quants.py:220-251, 378-402; its Q8_0 is documented bit-exact to ggml-quants.c).
The reference repo commits no golden tensors, so these pin the oracle's codecs.
Run in the build container:  python tests/golden/make_golden.py
"""
import os

import numpy as np
from gguf import quants
from gguf.constants import GGMLQuantizationType as T

rng = np.random.default_rng(20240320)
rows = []
rows.append(rng.standard_normal((8, 256)).astype(np.float32) * 0.02)          # weight-like
rows.append(rng.standard_normal((4, 256)).astype(np.float32))                  # activation-like
rows.append((rng.standard_normal((2, 256)) * 50).astype(np.float32))           # large
edge = np.zeros((6, 256), dtype=np.float32)
edge[1, :] = 1.0
edge[2, ::2] = -3.5
edge[2, 1::2] = 3.5
edge[3, :32] = np.linspace(-127, 127, 32, dtype=np.float32)                    # exact grid
edge[3, 32:64] = np.linspace(-127, 127, 32, dtype=np.float32) + 0.5            # ties for roundf
edge[4, :] = np.float32(1e-30)
edge[5, 5] = -7.25
edge[5, 40] = 7.25                                                              # +/- max tie
rows.append(edge)
x = np.concatenate(rows, axis=0)

out = {"x": x}
for name, t in (("q4_0", T.Q4_0), ("q8_0", T.Q8_0)):
    q = quants.quantize(x, t)
    out[name] = q.reshape(x.shape[0], -1)
    out[name + "_deq"] = quants.dequantize(q, t).reshape(x.shape).astype(np.float32)
# fp16 conversion KAT: numpy's IEEE binary16 round-to-nearest-even
f = np.concatenate([rng.standard_normal(512).astype(np.float32) * s for s in (1e-7, 1e-4, 1, 300, 7e4)])
f = np.concatenate([f, np.float32([0, -0.0, 65504, 65519.996, 65520, 2.0 ** -25, 2.0 ** -24, 6.1e-5, 5.96e-8, 1e-10])])
out["f32"] = f
with np.errstate(over="ignore"):
    out["f16_bits"] = f.astype(np.float16).view(np.uint16)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "codecs_gguf.npz"), **out)
print({k: v.shape for k, v in out.items()})
