import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the native artefacts once per session (CUDA library cross-compiles without a GPU)."""
    import __graft_entry__ as g
    so = os.path.join(ROOT, "mcmc-date_b200", "libmcmcdate_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(so) and os.path.exists(orc)):
        g.build()
    yield
