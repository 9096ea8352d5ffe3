"""pytest configuration: registers the `gpu` marker and makes the repo root importable."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native_lib():
    """Built libb200sr.so (compiles for sm_100a if stale; works without a GPU)."""
    import framewright_b200  # noqa: F401
    from framewright_b200 import _native

    _native.build()
    return _native.load()
