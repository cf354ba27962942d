"""Flipout layers (Wen et al. 2018) — SURVEY §8(f) rank 1: `FlipoutNormalLinear` (pytorch_bayesian/nn/dense.py:63-83)
and `FlipOutNormalConvNd/1d/2d/3d` (conv.py:145-251).

The perturbation is sigma itself with random sign flips,
    y = op(x, mean) + op(x * S, stddev) * R,        R, S in {-1, +1}.
Reference quirks kept: no bias; the Linear signs are per feature and shared by the whole batch (dense.py:71-75), the conv
signs are per example (conv.py:154-161).

* FlipoutNormalLinear (CUDA): because R and S are shared by the batch,  (x * S) sigma^T * R = x (sigma o (R S^T))^T, so the
  layer IS a sampled contraction  y = x (mean + sigma o eps)^T  with the rank-one sign noise eps = R S^T: ONE launch of the
  sample-and-contract kernels (bnn_rng.row_sign / col_sign: the generator warps read the two sign vectors instead of
  drawing Philox normals), where the reference runs two GEMMs and two elementwise passes; the backward kernels form
  d mean and d scale with the same eps.  R / S are still drawn by torch and kept as attributes (`sampled`).
* The conv layers' signs differ per example, which makes the perturbed filter example specific: those stay torch
  composites (two cuDNN convolutions with unsampled weights).  CPU tensors take the composite everywhere.
Their variational tensor is a WeightNormal, hence KLDivergence and PruneNormal treat them like every other Bayesian layer
(the fused KL / prune kernels).
"""
import torch
import torch.nn.functional as F
from torch.distributions.normal import Normal

from .. import runtime
from ..utils.traversal import _pair, _single, _triple
from .layers import NormalConvNd, NormalLinear


def _signs(*shape, device):
    return (torch.rand(*shape, device=device) - .5).sign()


_FUSED_LINEAR = {"on": True}


def set_fused_flipout_linear(flag=True):
    """False: FlipoutNormalLinear evaluates the reference's two-contraction composite on CUDA too (A/B comparisons)."""
    _FUSED_LINEAR["on"] = bool(flag)


class FlipoutNormalLinear(NormalLinear):
    _fused = False        # not a Philox-noise layer: the container treats it as a composite ...
    _mc_composite = True  # ... that takes part in the batched Monte-Carlo forward (S sign draws in one call)

    def __init__(self, in_features, out_features, prior=Normal(0, .1)):
        super(FlipoutNormalLinear, self).__init__(in_features, out_features, False, prior)

    def sample(self):
        self.R = _signs(self.weight.size(0), device=self.weight.device)
        self.S = _signs(self.weight.size(1), device=self.weight.device)
        self._signs = None

    @property
    def sampled(self):
        return (self.R, self.S)

    def _forward_fused(self, x, sample):
        """One sampled contraction with eps = R S^T per Monte-Carlo sample (module doc)."""
        from ..functional import DrawSpec, SampledLinear
        ctx = runtime.current_mc()
        S = 1 if ctx is None else ctx.samples
        shared = ctx is None or not ctx.expanded
        if ctx is not None and shared and x.shape[0] != ctx.rows:
            raise RuntimeError("batched Monte-Carlo forward: the first Bayesian layer must see the network "
                               f"input rows ({ctx.rows}), got {x.shape[0]}")
        dev = self.weight.device
        if sample:
            R = _signs(S, self.weight.size(0), device=dev)
            Sg = _signs(S, self.weight.size(1), device=dev)
            self._signs = (R, Sg)
            self.R, self.S = R[-1], Sg[-1]
        elif S == 1:              # dense.py:77-79: reuse the attributes (a previous draw, sample(), or set by the caller)
            self._signs = (self.R.reshape(1, -1).to(dev, torch.float32).contiguous(),
                           self.S.reshape(1, -1).to(dev, torch.float32).contiguous())
        elif getattr(self, "_signs", None) is None or self._signs[0].shape[0] != S:
            raise RuntimeError("forward(sample=False) needs a previous draw with the same number of MC samples")
        spec = DrawSpec(0, 0, 0, signs=self._signs)
        lead = x.shape[:-1]
        y = SampledLinear.apply(x.reshape(-1, x.shape[-1]), self.weight.mean, self.weight.scale, None, None, S, shared,
                                spec, None, runtime.precision())
        if ctx is not None:
            ctx.expanded = True
        if shared and S > 1:
            return y.view((S * lead[0],) + tuple(lead[1:]) + (y.shape[-1],))
        return y.view(tuple(lead) + (y.shape[-1],))

    def forward(self, x, sample=True):
        if x.is_cuda and _FUSED_LINEAR["on"] and x.dtype == torch.float32:
            return self._forward_fused(x, sample)
        S, x, ctx = runtime.mc_expand_rows(x)
        if ctx is not None and S > 1 and sample:
            # S Monte-Carlo passes at once: one sign pair per sample, shared by that sample's B rows (dense.py:71-75)
            R = _signs(S, 1, self.weight.size(0), device=self.weight.device)
            Sg = _signs(S, 1, self.weight.size(1), device=self.weight.device)
            self.R, self.S = R[-1, 0], Sg[-1, 0]
            x3 = x.reshape(S, -1, x.shape[-1])
            perturbation = ((x3 * Sg).matmul(self.weight.stddev.t()) * R).reshape(x.shape[:-1] + (self.weight.size(0),))
            return F.linear(x, self.weight.mean) + perturbation
        if sample:
            self.sample()
        perturbation = (x * self.S).matmul(self.weight.stddev.t()) * self.R
        return F.linear(x, self.weight.mean, perturbation)       # dense.py:81-83: the perturbation rides in as the bias


class FlipOutNormalConvNd(NormalConvNd):
    _op = None            # set by the dimensional subclasses
    _fused = False
    _mc_composite = True  # the signs are per example (conv.py:154-161): S*B rows simply draw S*B sign sets

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation, transposed, groups, prior):
        super(FlipOutNormalConvNd, self).__init__(in_channels, out_channels, _single(kernel_size), stride, padding,
                                                  dilation, transposed, groups, False, prior)

    def sample(self, batch_size=1, additional_dims=()):
        self.R = _signs(batch_size, self.weight.size(0), *additional_dims, device=self.weight.device)
        self.S = _signs(batch_size, self.weight.size(1), *additional_dims, device=self.weight.device)

    @property
    def sampled(self):
        return (self.R, self.S)

    def _flipout(self, x, nd, sample):
        _, x, _ = runtime.mc_expand_rows(x)
        if sample:
            self.sample(x.size(0), (1,) * nd)
        op = type(self)._op
        args = (self.stride, self.padding, self.dilation, self.groups)
        out = op(x, self.weight.mean, self.bias, *args)
        return out + op(x * self.S.expand_as(x), self.weight.stddev, self.bias, *args) * self.R.expand_as(out)


class FlipOutNormalConv1d(FlipOutNormalConvNd):
    _op = staticmethod(F.conv1d)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior=Normal(0, .1)):
        super(FlipOutNormalConv1d, self).__init__(in_channels, out_channels, _single(kernel_size), _single(stride),
                                                  _single(padding), _single(dilation), False, groups, prior)

    def forward(self, x, sample=True):
        return self._flipout(x, 1, sample)


class FlipOutNormalConv2d(FlipOutNormalConvNd):
    _op = staticmethod(F.conv2d)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior=Normal(0, .1)):
        super(FlipOutNormalConv2d, self).__init__(in_channels, out_channels, _pair(kernel_size), _pair(stride),
                                                  _pair(padding), _pair(dilation), False, groups, prior)

    def forward(self, x, sample=True):
        return self._flipout(x, 2, sample)


class FlipOutNormalConv3d(FlipOutNormalConvNd):
    _op = staticmethod(F.conv3d)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior=Normal(0, .1)):
        super(FlipOutNormalConv3d, self).__init__(in_channels, out_channels, _triple(kernel_size), _triple(stride),
                                                  _triple(padding), _triple(dilation), False, groups, prior)

    def forward(self, x, sample=True):
        return self._flipout(x, 3, sample)
