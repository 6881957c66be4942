import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """The product library and the oracle are built in-tree; rebuild only if missing (make is incremental)."""
    lib = os.path.join(ROOT, "craytracer_b200", "libcray_b200.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "craytracer_b200", "csrc")], check=True, capture_output=True)
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    yield


@pytest.fixture(scope="session")
def orc():
    import oracle_lib
    return oracle_lib.lib()
