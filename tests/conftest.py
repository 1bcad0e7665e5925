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
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(GOLDEN, name), weights_only=False)
    return load


@pytest.fixture(scope="session")
def hps():
    """The product package, with the CUDA library built (fails loudly if it cannot be)."""
    import importlib
    import __graft_entry__
    __graft_entry__.build()
    return importlib.import_module("hierarchical-point-slam_b200")
