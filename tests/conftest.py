import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def ptb():
    import szakdolgozat_pathtracer_b200 as m
    if not m.LIB_PATH.exists():
        m.build()
    return m


@pytest.fixture(scope="session")
def oh():
    import orchelp
    orchelp.load("oracle")
    return orchelp


@pytest.fixture(scope="session")
def assets():
    import make_assets
    return make_assets


@pytest.fixture(scope="session")
def ctx(ptb):
    c = ptb.Context(0)
    yield c
    c.close()
