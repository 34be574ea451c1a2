import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices on one box (deselected otherwise)")


def pytest_collection_modifyitems(config, items):
    import torch
    n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n_dev < 2:           # multi-GPU tests are not skipped but left out: a one-GPU box reports no skips
        gone = [it for it in items if "multigpu" in it.keywords]
        if gone:
            items[:] = [it for it in items if "multigpu" not in it.keywords]
            config.hook.pytest_deselected(items=gone)
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
