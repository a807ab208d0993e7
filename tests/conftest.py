import os
import sys

import pytest

# Before the first CUDA call of the test process: one hardware queue per stream (the default is 8 queues), so that the
# streams of several proofs in flight do not share a queue and no kernel is queued behind another proof's exchange wait.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT
