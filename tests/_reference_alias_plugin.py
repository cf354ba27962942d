"""pytest plugin used by tests/test_reference_suite_gpu.py: runs the REFERENCE's own test-suite (baseline/_ref/tests,
unmodified) against this package — `pytorch_bayesian` is aliased to `bayesianneuralnetworks_b200`, new
tensors default to the CUDA device (the reference's tests build everything on the default device; the drop-in has no CPU
path), and torch.distributions argument validation is off (reference tests/test_prune.py passes an int to log_prob,
which torch >= 1.8 rejects otherwise — SURVEY §0)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bayesianneuralnetworks_b200 as bnn  # noqa: E402

for name in ("", ".nn", ".prune", ".utils"):
    sys.modules["pytorch_bayesian" + name] = sys.modules["bayesianneuralnetworks_b200" + name]
torch.distributions.Distribution.set_default_validate_args(False)
torch.set_default_device("cuda")
bnn.set_precision("fp32")
