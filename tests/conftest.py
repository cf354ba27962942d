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


import pytest  # noqa: E402


@pytest.fixture(params=["in_place", "into"])
def prune_mode(request, monkeypatch):
    """Runs a prune test twice: through bnn_prune (in place, two sweeps) and through bnn_prune_into (one out-of-place sweep,
    results copied back so that the test's in-place assertions apply unchanged).  Requests for keys_out stay on bnn_prune
    (the general path is the only one that stores keys)."""
    if request.param == "into":
        from bayesianneuralnetworks_b200 import _C
        real = _C.prune

        def via_into(entries, flags=0):
            keyed = [e for e in entries if e[4] is not None]
            rest = [e for e in entries if e[4] is None]
            if keyed:
                real(keyed, flags=flags)
            if rest:
                before = [(e[0].clone(), e[1].clone()) for e in rest]
                outs = _C.prune_into([(mu, rho, k, mask) for mu, rho, k, mask, _ in rest], flags=flags)
                for (mu, rho, *_), (mo, ro), (m0, r0) in zip(rest, outs, before):
                    assert bool((mu == m0).all() | (mu != mu).any()) and mo.data_ptr() != mu.data_ptr()   # inputs untouched
                    mu.copy_(mo)
                    rho.copy_(ro)
        monkeypatch.setattr(_C, "prune", via_into)
    return request.param
