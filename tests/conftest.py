import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    return O.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own headers compiled verbatim; only available where /root/reference or a prebuilt .so exists."""
    import oracle as O
    try:
        return O.Ref()
    except (FileNotFoundError, OSError) as e:
        pytest.skip("oracle/_ref not available: %s" % e)


@pytest.fixture(scope="session")
def gcr_ctx():
    from pygcransac import _cabi
    ctx = _cabi.Context(0)
    yield ctx
    ctx.close()
