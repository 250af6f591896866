import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _gpu_count():
    try:
        from babelbrain_b200 import _capi
        return _capi.lib().bb_device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a machine without a GPU skips the `gpu` tests instead of failing them (the product path has no
    CPU fallback and raises there).  On a GPU box nothing is skipped: a missing library fails loudly."""
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason='no CUDA device (or libbabelb200.so not built): the CUDA path has no CPU fallback')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
