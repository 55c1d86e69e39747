"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol export (no device calls).
`-m gpu`      : parity tests proper — the CUDA path through the C-ABI against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        # a hung kernel must be a failure with a stack trace, not a lost box: every GPU test gets a
        # timeout (pytest-timeout, thread method: a blocked ctypes call cannot be interrupted by a signal)
        if config.pluginmanager.hasplugin("timeout"):
            for item in items:
                if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                    item.add_marker(pytest.mark.timeout(1200 if "sharded" in item.name else 300, method="thread"))
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load
