import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: differential test against the live reference (build container only)")


def pytest_sessionstart(session):
    """Make sure the native artefacts exist (no-op when they are up to date): libf110_b200.so (nvcc cross-compiles without a
    GPU) and the CPU oracle.  The GPU box receives the prebuilt files with the snapshot."""
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    from tests import helpers
    return helpers
