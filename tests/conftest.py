import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_present():
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    # GPU tests never fall back to anything: without a device they are skipped, with a device but
    # without the built library they fail.
    if _gpu_present():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
