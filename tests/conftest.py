import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0.  No skip / fallback: a gpu-marked test that cannot get
    a device must fail loudly."""
    import multitypetree_b200 as mtt
    c = mtt.Context(0)
    yield c
    c.close()
