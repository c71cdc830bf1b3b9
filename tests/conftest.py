import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference build (only in the build container)")


def _have_gpu() -> bool:
    try:
        from hydpy_b200 import engine

        return engine.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                have = _have_gpu()
            if not have:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
