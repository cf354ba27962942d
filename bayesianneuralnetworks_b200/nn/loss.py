"""KLDivergence and Entropy (mirror of pytorch_bayesian/nn/loss.py:11-51).

KLDivergence visits the same tensors in the same order as the reference (traverse / apply_wb) and
returns the same scalar — mean over elements per tensor, mean over tensors, / n_batches — but all
tensors are reduced by ONE bandwidth-bound launch, and the priors are read as host scalars (no
torch.distributions object, hence no validation sync, on the hot path).
"""
import warnings

import torch
from torch.nn import Module

from torch.distributions import MultivariateNormal, Normal
from torch.distributions.kl import kl_divergence

from ..functional import KLSum
from ..utils.traversal import apply_wb
from .mvn import WeightMultivariateNormal
from .variational import WeightNormal


def _is_scalar_normal(prior):
    loc, scale = getattr(prior, 'loc', None), getattr(prior, 'scale', None)
    return (isinstance(prior, Normal) and loc is not None and scale is not None
            and torch.as_tensor(loc).numel() == 1 and torch.as_tensor(scale).numel() == 1)


def _scalar_prior(prior, what):
    if not _is_scalar_normal(prior):
        raise NotImplementedError(
            f"the fused kernel needs a scalar Normal(loc, scale) prior for {what}; got {prior!r}")
    return float(prior.loc), float(prior.scale)


def fused_kl_entries(model):
    """[(WeightNormal, prior_loc, prior_scale)] of every variational tensor the traversal finds, in its order (weights
    before biases, loss.py:17-20) — for callers that need the KL of each tensor from a fused sweep (prune.kl_and_prune).
    Raises NotImplementedError for tensors without a fused form (full covariance, tensor-valued priors)."""
    def visit(param, module, type):
        if not isinstance(param, WeightNormal):
            raise NotImplementedError(f"no fused KL for {param.__class__.__name__}")
        loc, scale = _scalar_prior(module.weight_prior if type == 'w' else module.bias_prior,
                                   f"{module.__class__.__name__}.{'weight' if type == 'w' else 'bias'}")
        return (param, loc, scale)
    found = model.traverse(lambda m: apply_wb(m, visit, pass_module=True, pass_type=True))
    if found is None:
        raise ValueError('KLDivergence was not able to find BayasianModules')    # loss.py:34-36
    return found


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
            # the prior's tensors live on the CPU: keep one device copy per (module, tensor) instead of
            # an H2D copy per call, and skip argument validation (a device->host sync per object) —
            # both are what makes this term capturable in a CUDA graph; the arithmetic is torch's
            cache = module.__dict__.setdefault('_bnn_prior_cache', {})
            key = (type, param.device)
            hit = cache.get(key)
            if hit is None or hit[0] is not prior:
                standard = bool((prior.mean == 0).all()) and bool(
                    (prior.scale_tril == torch.eye(prior.scale_tril.shape[-1]).expand_as(prior.scale_tril)).all())
                hit = (prior, MultivariateNormal(prior.mean.to(param.device), scale_tril=prior.scale_tril.to(param.device),
                                                 validate_args=False), standard)
                cache[key] = hit
            if hit[2]:
                # standard-normal prior (the layer's default, dense.py:91-96): _kl_multivariatenormal_multivariatenormal
                # (torch/distributions/kl.py) reduces to  0.5 (tr(L L^T) + |mu|^2 - k) - sum_i log L_ii  — the same value
                # without the batched triangular solves (C3 step: ~40 small launches fewer)
                tril = param.variance
                k = tril.shape[-1]
                half = 0.5 * (tril.pow(2).sum((-2, -1)) + param.mean.pow(2).sum(-1) - k)
                return (half - tril.diagonal(dim1=-2, dim2=-1).log().sum(-1)).mean()
            posterior = MultivariateNormal(param.mean, scale_tril=param.variance, validate_args=False)
            return kl_divergence(posterior, hit[1]).mean()
        if not isinstance(param, WeightNormal):
            raise NotImplementedError(f"KLDivergence: unsupported variational tensor {param.__class__.__name__}")
        if not _is_scalar_normal(prior):
            # a tensor-valued (per-element) or non-Normal prior, which the reference accepts: torch.distributions, as
            # loss.py:28 does (the prior's tensors follow the parameter's device; no validation sync)
            if isinstance(prior, Normal):
                prior = Normal(torch.as_tensor(prior.loc).to(param.device), torch.as_tensor(prior.scale).to(param.device),
                               validate_args=False)
            return kl_divergence(Normal(param.mean, param.stddev, validate_args=False), prior).mean()
        return (param, _scalar_prior(prior, f"{module.__class__.__name__}.{'weight' if type == 'w' else 'bias'}"))

    def _gather(self, model, which):
        """(n, fused entries, torch-composite terms) of the traversal; `which` = 'fused' skips evaluating the composite
        terms, 'composite' skips nothing but returns no fused entries' work."""
        def visit(param, module, type):
            fusable = isinstance(param, WeightNormal) and _is_scalar_normal(
                module.weight_prior if type == 'w' else module.bias_prior)
            if which == 'fused' and not fusable:
                return 0.0                                  # counted in n, not evaluated
            if which == 'composite' and fusable:
                return 0.0
            return self.compute_kl(param, module, type)
        found = model.traverse(lambda m: apply_wb(m, visit, pass_module=True, pass_type=True))
        if found is None:
            raise ValueError('KLDivergence was not able to find BayasianModules')    # loss.py:34-36
        return len(found), [f for f in found if isinstance(f, tuple)], [f for f in found if torch.is_tensor(f)]

    def _fused_total(self, n, fused):
        priors = [p for _, p in fused]
        coeffs = [1.0 / (w.mean.numel() * n * self.n_batches) for w, _ in fused]
        flat = [t for w, _ in fused for t in (w.mean, w.scale)]
        return KLSum.apply(priors, coeffs, *flat)

    def fused_part(self, model):
        """The share of forward() that the fused kernel evaluates (factorised Gaussians with scalar priors), or None."""
        n, fused, _ = self._gather(model, 'fused')
        return self._fused_total(n, fused) if fused else None

    def composite_part(self, model):
        """The share of forward() that goes through torch.distributions (full-covariance tensors, tensor-valued priors),
        or None.  forward() == fused_part() + composite_part(); training.ElboTrainer differentiates only this share
        through autograd — the fused share's gradient is a closed form added by optim.ELBOAdam."""
        n, _, other = self._gather(model, 'composite')
        return torch.stack(other).sum() / (n * self.n_batches) if other else None

    def forward(self, model):
        n, fused, other = self._gather(model, 'all')        # loss.py:38: mean over ALL listed tensors, / n_batches
        total = self._fused_total(n, fused) if fused else None
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
