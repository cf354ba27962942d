"""KLDivergence and Entropy (mirror of pytorch_bayesian/nn/loss.py:11-51).

KLDivergence visits the same tensors in the same order as the reference (traverse / apply_wb) and
returns the same scalar — mean over elements per tensor, mean over tensors, / n_batches — but all
tensors are reduced by ONE bandwidth-bound launch, and the priors are read as host scalars (no
torch.distributions object, hence no validation sync, on the hot path).
"""
import warnings

import torch
from torch.nn import Module

from ..functional import KLSum
from ..utils.traversal import apply_wb
from .variational import WeightNormal


def _scalar_prior(prior, what):
    loc, scale = getattr(prior, 'loc', None), getattr(prior, 'scale', None)
    if loc is None or scale is None or torch.as_tensor(loc).numel() != 1 or torch.as_tensor(scale).numel() != 1:
        raise NotImplementedError(
            f"KLDivergence: the fused kernel needs a scalar Normal(loc, scale) prior for {what}; got {prior!r}")
    return float(loc), float(scale)


class KLDivergence(Module):
    def __init__(self, number_of_batches=1):
        super(KLDivergence, self).__init__()
        self.n_batches = number_of_batches

    def compute_kl(self, param, module, type):
        """loss.py:16-28 — here it only gathers (tensor, prior); the arithmetic happens in one launch."""
        if not isinstance(param, WeightNormal):
            raise NotImplementedError(f"KLDivergence: unsupported variational tensor {param.__class__.__name__}")
        prior = module.weight_prior if type == 'w' else module.bias_prior
        return (param, _scalar_prior(prior, f"{module.__class__.__name__}.{'weight' if type == 'w' else 'bias'}"))

    def forward(self, model):
        found = model.traverse(lambda m: apply_wb(m, self.compute_kl, pass_module=True, pass_type=True))
        if found is None:
            raise ValueError('KLDivergence was not able to find BayasianModules')    # loss.py:34-36
        n = len(found)
        priors = [p for _, p in found]
        coeffs = [1.0 / (w.mean.numel() * n * self.n_batches) for w, _ in found]
        flat = [t for w, _ in found for t in (w.mean, w.scale)]
        return KLSum.apply(priors, coeffs, *flat)


class Entropy(Module):
    """loss.py:41-51 (a metric on the mean prediction; plain torch)."""

    def __init__(self, dim=0):
        super(Entropy, self).__init__()
        self.dim = dim

    def forward(self, x):
        if (x == 0).all():
            warnings.warn('Entropy received a tensor containing all zeros', RuntimeWarning)
        return (-x * torch.log(x + 1e-10)).sum(dim=self.dim).mean()
