import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

GOLDEN = os.path.join(REPO, "tests", "golden")
TBL = os.path.join(GOLDEN, "tbl")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def device():
    """hyb_context on cuda:0 — GPU tests call the product exclusively through the C-ABI."""
    from hyrise_b200.device import DeviceContext

    context = DeviceContext(0)
    yield context
    context.close()
