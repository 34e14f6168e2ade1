import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One C-ABI handle on cuda:0 shared by the GPU tests (fails loudly when the library or GPU is missing)."""
    import itcpd

    eng = itcpd.Engine(0)
    yield eng
    eng.close()
