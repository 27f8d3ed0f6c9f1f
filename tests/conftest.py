import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference lib.py (tests/golden/make_golden.py)."""
    with open(os.path.join(REPO, "tests", "golden", "reference_lib_golden.json")) as f:
        return json.load(f)
