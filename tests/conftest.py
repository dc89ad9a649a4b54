import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly rather than silently pass: the kernels are the product and there is
    # no CPU fallback behind them.
    expr = (config.getoption("-m") or "").strip()
    if expr == "gpu" and any(it.get_closest_marker("gpu") for it in items):
        import torch
        if not torch.cuda.is_available():
            raise pytest.UsageError("`-m gpu` selected but no CUDA device is visible: the GPU tests cannot run here")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
