import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nh3_golden():
    return dict(np.load(GOLDEN / "nh3_golden.npz"))


@pytest.fixture(scope="session")
def gauss_golden():
    return dict(np.load(GOLDEN / "gauss_golden.npz"))


@pytest.fixture(scope="session")
def prior_golden():
    return dict(np.load(GOLDEN / "prior_golden.npz"))


@pytest.fixture(scope="session")
def nb():
    """The product package with its CUDA library built (no GPU needed to build)."""
    from nestfit_b200 import build
    build.build()
    import nestfit_b200
    return nestfit_b200


@pytest.fixture(scope="session")
def n2hp_golden():
    return dict(np.load(GOLDEN / "n2hp_golden.npz"))
