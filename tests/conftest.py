import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from spin_ed_b200 import ffi

        return ffi.deviceCount() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when selected on a machine without a device they must
    # FAIL loudly (no silent skip, no CPU fallback), so nothing is skipped here.
    return


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O
