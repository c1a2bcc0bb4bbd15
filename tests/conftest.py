import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liborc.so), built on demand. Test infrastructure only."""
    import orc as _orc

    _orc.lib()
    return _orc


@pytest.fixture(scope="session")
def crt():
    """The product binding (libcrt_b200.so must already be built in-tree)."""
    import crust_render_b200 as _crt

    _crt.lib()
    return _crt


@pytest.fixture(scope="session")
def gpu(crt):
    """Product binding bound to cuda:0; fails loudly (no skip) when there is no device."""
    crt.init(0)
    return crt
