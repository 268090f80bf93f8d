import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` under gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree by nvcc (works on the CPU-only container: cross-compilation)."""
    from cudalibrarysamples_b200 import build

    return build.build_native()
