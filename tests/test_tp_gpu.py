"""Multi-GPU parity: spawns one process per GPU (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _gpu_count():
    try:
        import pynvml
        pynvml.nvmlInit()
        return pynvml.nvmlDeviceGetCount()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_decode_matches_oracle(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29500 + (os.getpid() % 400) + world
    procs = [subprocess.Popen([sys.executable, os.path.join(here, "tp_worker.py"), str(r), str(world), str(port)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=600)[0].decode())
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
