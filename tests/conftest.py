import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def qp():
    import qprop_b200

    return qprop_b200


@pytest.fixture(scope="session")
def ctx(qp):
    """One device context for the whole GPU session (fails loudly without a GPU)."""
    return qp.default_context(0)
