import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def native():
    """The pybind module; building it if this checkout has not been built yet."""
    try:
        from sapien_b200 import simsense
    except ImportError:
        from sapien_b200 import _build

        _build.build_all()
        from sapien_b200 import simsense
    return simsense
