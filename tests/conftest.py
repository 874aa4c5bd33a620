import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "tests")):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def library():
    """The product library.  Building is __graft_entry__.build()'s / make's job; a missing library is an error."""
    import libgiddy_b200
    return libgiddy_b200.default_library()


@pytest.fixture(scope="session")
def reference_library():
    """The reference's own kernels rebuilt for sm_100a (oracle/_ref), when present."""
    import libgiddy_b200
    path = os.path.join(ROOT, "oracle", "_ref", "libgiddy_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libgiddy_ref.so has not been built (make ref)")
    return libgiddy_b200.Library(path)
