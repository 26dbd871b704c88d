import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def renderer():
    """One rt_ctx on cuda:0 for the whole GPU session.  Fails loudly without the CUDA library."""
    import raytracingincuda_b200 as rt
    r = rt.Renderer(0)
    yield r
    r.close()
