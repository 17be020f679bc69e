import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) where there is no CUDA device or the library has not been built."""
    import torch
    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    else:
        from tulip_b200._lib import lib_path
        if not os.path.exists(lib_path()):
            reason = "libtulip_b200.so not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
