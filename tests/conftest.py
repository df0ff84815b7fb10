"""pytest configuration.  `-m "not gpu"` runs on a CPU-only box (oracle vs golden
vectors, host logic, C-ABI export check); `-m gpu` holds the parity tests proper
and needs a B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle and the product library exist (build() is idempotent)."""
    import __graft_entry__ as g
    g.build()
    return True
