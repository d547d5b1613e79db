import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import _pkg
    return _pkg.load_pkg()


@pytest.fixture(scope="session")
def synth():
    import _pkg
    return _pkg.load_synth()


@pytest.fixture(scope="session")
def simdir():
    import build_sim
    return build_sim.build()
