import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test (excluded from the default CPU run with -m 'not slow')")


@pytest.fixture(scope="session")
def built():
    """Everything compiled (idempotent: make only rebuilds what changed)."""
    entry.build()
    return True


@pytest.fixture(scope="session")
def pkg(built):
    return entry.load_package()


@pytest.fixture(scope="session")
def orc(built):
    return entry.load_oracle()


def has_gpu() -> bool:
    try:
        return entry.load_package().library().lbm_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu(pkg):
    n = pkg.library().lbm_device_count()
    if n <= 0:
        # a gpu-marked test on a box without a device must fail loudly, not skip: there is no fallback
        pytest.fail("no CUDA device visible: -m gpu tests need a B200")
    return n
