import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from nutpie_b200 import _lib

        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently; but when
    # the whole suite is collected on a CPU box the gpu tests are deselected by the
    # driver's `-m "not gpu"`.  Nothing to do here.
    return


@pytest.fixture(scope="session")
def radon_data():
    from nutpie_b200.datasets import make_radon_data

    return make_radon_data()
