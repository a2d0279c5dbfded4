import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def b200():
    """The product package (directory `montecarlo.jl_b200`, imported as montecarlo_jl_b200)."""
    import _b200_loader
    return _b200_loader.load()
