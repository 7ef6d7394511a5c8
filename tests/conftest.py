import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/liboracle.so).  Test infrastructure only."""
    import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def ctx():
    """A pb200 context on cuda:0.  Fails loudly (no fallback) when there is no GPU."""
    import plonk_prototype_b200 as pb
    c = pb.Context(0)
    yield c
    c.close()
