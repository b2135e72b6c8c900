import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def csg():
    import certificate_stark_b200 as m
    if not m.LIB_PATH.exists():
        m.build()
    m.lib()
    return m


@pytest.fixture(scope="session")
def ctx(csg):
    c = csg.Context(0)
    yield c
    c.close()
