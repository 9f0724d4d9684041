import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def built():
    """CPU-side artefacts the tests need: the oracle port, the generator and the product library
    (compiled here without a GPU; prebuilt copies travel to the GPU box)."""
    import __graft_entry__ as ge
    ge.build()
    return True
