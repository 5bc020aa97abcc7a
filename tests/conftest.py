import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libonephase_b200.so")):
        g.build()
    return g.package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as g
    return g.oracle()


@pytest.fixture(scope="session")
def have_gpu():
    import torch
    return torch.cuda.is_available()
