import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests bind the in-tree shared library.  It normally travels with the tree (built by
    __graft_entry__.build()); if it is missing and nvcc is here, build it once -- never fall back to anything."""
    lib = os.path.join(ROOT, "longbow_b200", "liblongbow_b200.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "longbow_b200", "csrc"), "-j8", "-s"])
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o
