"""pytest configuration: the `gpu` marker (tests that need a B200) and import paths.

`python -m pytest tests -m "not gpu"` runs here on CPU (oracle vs golden vectors, host logic, C-ABI
symbol check); `-m gpu` are the parity tests proper and call the CUDA library through its C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device of compute capability 10.x (B200)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
