import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle as _o
    _o.build()
    return _o


@pytest.fixture(scope="session", autouse=True)
def _product_library():
    """The C-ABI library is built in tree (nvcc cross-compiles without a GPU); tests never fall back to
    anything else when it is missing."""
    import os
    from dugksfoam_b200 import capi
    if not os.path.exists(capi.LIB_PATH):   # only when absent: never a surprise rebuild on the GPU box
        capi.build_library()
    yield
