import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lct-vqa_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _stock_ops_off_by_default():
    """pcd_ops.allow_stock_ops is process-global: cases that opt in (toy dimensions) must not leak into the next test."""
    yield
    mod = sys.modules.get("pcd_ops")
    if mod is not None:
        mod.allow_stock_ops(False)
