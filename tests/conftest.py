import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("SOURCE_DATE_EPOCH", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the in-tree build products exist (cheap when already built)."""
    import __graft_entry__ as g
    import jfutil
    if not (os.path.exists(jfutil.LIB) and os.path.exists(jfutil.OUR_JF) and os.path.exists(jfutil.ORACLE_C)):
        g.build()
    return True


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("jf"))


@pytest.fixture(scope="session")
def inputs(workdir):
    """The deterministic input files every parity test uses (tests/gen.py)."""
    import gen
    return gen.make_all(workdir)
