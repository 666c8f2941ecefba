import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tests that run several ranks inside one process (one stream each, spinning on each other's
# flags) need every stream on its own hardware queue: the default of 8 connections aliases
# streams once a dozen exist.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and every kernel loaded up front: with lazy loading the FIRST launch of a kernel synchronises
# with the device, i.e. with a rank's spinning wait kernel whose peer has not been enqueued yet by
# this same host thread.  (One process per GPU never meets this: a rank's own puts precede its waits.)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
