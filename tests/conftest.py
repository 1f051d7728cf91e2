import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: longer CPU test")


def _gpu_available():
    try:
        from debwt_b200 import binding
        return binding.lib().debwt_device_count() > 0
    except Exception:  # noqa: BLE001  (library not built, no driver)
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device or without the built library"""
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and debwt_b200/libdebwt_b200.so (no CPU fallback)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
