import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu; when someone runs the whole suite on a CPU box they skip loudly.
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        path = os.path.join(GOLDEN, name)
        if name.endswith(".json"):
            with open(path) as f:
                return json.load(f)
        import numpy as np
        return np.load(path, allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree libva_b200.so (built here if missing; nvcc cross-compiles without a GPU)."""
    from video_analytics_b200.build import build_library
    return build_library()
