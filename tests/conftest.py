import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library; GPU tests fail loudly (no fallback) if it is missing or no device exists."""
    import pomcp_jl_b200  # noqa: F401
    from pomcp_jl_b200._lib import lib
    return lib()
