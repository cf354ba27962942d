"""PruneNormal (mirror of pytorch_bayesian/prune/prune.py:5-22).

Per variational tensor, independently: key = log N(0; mean, stddev), the k = int(percentage * numel)
largest keys get mean <- 0, scale <- -30, in place and without autograd.  All tensors of the model
go through the same launches of libbnn_b200's exact select (keys bit-identical to torch's
Normal.log_prob on the device; ties at the k-th key resolved towards the lowest index).
"""
import torch

from .. import _C
from ..nn.variational import WeightNormal
from ..utils.traversal import apply_wb


class PruneNormal():

    def __call__(self, module, percentage=0.5):
        self.prune(module, percentage)

    def prune_param(self, param, percentage):
        """prune.py:10-17 for one tensor."""
        self._launch([param], percentage)

    @staticmethod
    def _launch(params, percentage):
        entries = []
        for p in params:
            if not isinstance(p, WeightNormal):
                raise NotImplementedError(f"PruneNormal: unsupported variational tensor {p.__class__.__name__}")
            _C.require_cuda(p.mean, p.scale)
            n = p.mean.numel()
            k = int(percentage * n)          # prune.py:13 — float32 arithmetic when percentage is a tensor
            if k < 0 or k > n:
                raise RuntimeError(f"selected index k out of range: k={k}, numel={n}")
            if not (p.mean.is_contiguous() and p.scale.is_contiguous()):
                raise ValueError("PruneNormal needs contiguous mean/scale parameters")
            entries.append((p.mean.data, p.scale.data, k, None, None))
        _C.prune(entries)

    def prune(self, module, percentage=0.5):
        """prune.py:19-22 — every tensor the traversal finds, weights and biases alike."""
        with torch.no_grad():
            found = module.traverse(lambda m: apply_wb(m, lambda p: p))
            if found:
                self._launch(found, percentage)
