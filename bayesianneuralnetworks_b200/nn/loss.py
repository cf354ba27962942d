"""KLDivergence and Entropy (mirror of pytorch_bayesian/nn/loss.py:11-51).

KLDivergence visits the same tensors in the same order as the reference (traverse / apply_wb) and
returns the same scalar — mean over elements per tensor, mean over tensors, / n_batches — but all
tensors are reduced by ONE bandwidth-bound launch, and the priors are read as host scalars (no
torch.distributions object, hence no validation sync, on the hot path).
"""
import warnings

import torch
from torch.nn import Module

from torch.distributions import MultivariateNormal
from torch.distributions.kl import kl_divergence

from ..functional import KLSum
from ..utils.traversal import apply_wb
from .mvn import WeightMultivariateNormal
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
        """loss.py:16-28 — factorised Gaussians are only gathered here as (tensor, prior) and reduced by one launch;
        the full-covariance tensors of MultivariateNormalLinear (loss.py:24-26, out of the hot path) go through
        torch.distributions and come back as their mean KL."""
        prior = module.weight_prior if type == 'w' else module.bias_prior
        if isinstance(param, WeightMultivariateNormal):
            prior = MultivariateNormal(prior.mean.to(param.device), scale_tril=prior.scale_tril.to(param.device))
            return kl_divergence(param.dist, prior).mean()
        if not isinstance(param, WeightNormal):
            raise NotImplementedError(f"KLDivergence: unsupported variational tensor {param.__class__.__name__}")
        return (param, _scalar_prior(prior, f"{module.__class__.__name__}.{'weight' if type == 'w' else 'bias'}"))

    def forward(self, model):
        found = model.traverse(lambda m: apply_wb(m, self.compute_kl, pass_module=True, pass_type=True))
        if found is None:
            raise ValueError('KLDivergence was not able to find BayasianModules')    # loss.py:34-36
        n = len(found)                          # loss.py:38: mean over ALL listed tensors, / n_batches
        fused = [f for f in found if isinstance(f, tuple)]
        other = [f for f in found if not isinstance(f, tuple)]
        total = None
        if fused:
            priors = [p for _, p in fused]
            coeffs = [1.0 / (w.mean.numel() * n * self.n_batches) for w, _ in fused]
            flat = [t for w, _ in fused for t in (w.mean, w.scale)]
            total = KLSum.apply(priors, coeffs, *flat)
        if other:
            rest = torch.stack(other).sum() / (n * self.n_batches)
            total = rest if total is None else total + rest
        return total


class Entropy(Module):
    """loss.py:41-51 (a metric on the mean prediction; plain torch)."""

    def __init__(self, dim=0):
        super(Entropy, self).__init__()
        self.dim = dim

    def forward(self, x):
        if (x == 0).all():
            warnings.warn('Entropy received a tensor containing all zeros', RuntimeWarning)
        return (-x * torch.log(x + 1e-10)).sum(dim=self.dim).mean()
