import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the library and the oracle if they are not there yet (CPU only: nvcc cross-compiles)."""
    lib = os.path.join(ROOT, "tsxcount_b200", "lib", "libtsxcuda.so")
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        subprocess.run(["make", "-C", ROOT, "lib", "oracle"], check=True, capture_output=True)
    yield
