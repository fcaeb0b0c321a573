import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _metric_max_label():
    """The reference's data loader sets RankingMetricKey.MAX_LABEL (the ERR normaliser) when a data set is read; tests
    build their data in memory, so the plugin's copy of that setting gets the value of the 5-level test labels here
    (every test file on its own, in any order)."""
    try:
        from ultra_pytorch_b200 import metrics as b200_metrics
    except Exception:
        yield
        return
    saved = b200_metrics.MAX_LABEL
    if saved is None:
        b200_metrics.MAX_LABEL = 4.0
    yield
    b200_metrics.MAX_LABEL = saved
