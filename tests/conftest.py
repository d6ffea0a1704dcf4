import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def G():
    import gemma_ggml_b200 as g
    return g


@pytest.fixture(scope="session")
def gpu(G):
    """Initialised CUDA layer; a GPU test without a device must fail loudly, never fall back."""
    G.init(0)
    return G
