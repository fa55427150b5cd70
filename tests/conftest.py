import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skip them (instead of failing in the first .cuda()) on a box without one, so a
    plain `pytest tests` works everywhere.  On a GPU box nothing is skipped: a missing libmagb200.so must FAIL there."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:            # pragma: no cover
        reason = f"torch not importable: {e}"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
