import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("FL4H_LOG_LEVEL", "WARNING")


def pytest_configure(config):  # noqa: ANN001, ANN201
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on a B200 box)")
    config.addinivalue_line("markers", "multigpu: test needs >1 CUDA device")


def pytest_collection_modifyitems(config, items):  # noqa: ANN001, ANN201
    import torch

    if torch.cuda.is_available():
        return
    skip_gpu = pytest.mark.skip(reason="no CUDA device available")
    for item in items:
        if "gpu" in item.keywords or "multigpu" in item.keywords:
            item.add_marker(skip_gpu)
