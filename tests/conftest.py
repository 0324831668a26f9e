import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "flagstat_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly if it is missing or sees no device."""
    import libflagstats_b200 as fs

    n = fs.available()
    assert n > 0, "libflagstats_cuda.so loaded but reports no CUDA device"
    return fs
