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
* The conv layers' signs differ per example, which makes the perturbed filter example specific: there is no common
  sampled weight, the layer really is TWO convolutions with unsampled weights, conv(x, mean) + conv(x * S, stddev) * R.
  FlipOutNormalConv1d / 2d (CUDA, groups == 1, in_channels % 32 == 0) run both through the library's implicit-GEMM
  contraction kernels (the sample-and-contract kernels with injected eps = 0, i.e. W = the given tensor; TMA im2col
  operand, tcgen05 MMAs; forward, input gradient and weight gradients), with the sign products and d stddev / d scale
  left to torch element-wise ops; everything else (3-d, grouped, odd channel counts, CPU tensors) evaluates the
  reference's composite with torch's convolutions.
Their variational tensor is a WeightNormal, hence KLDivergence and PruneNormal treat them like every other Bayesian layer
(the fused KL / prune kernels).
"""
import torch
import torch.nn.functional as F
from torch.distributions.normal import Normal

from .. import runtime
from ..functional import conv_implicit_eligible
from ..utils.traversal import _pair, _single, _triple
from .layers import NormalConvNd, NormalLinear


def _signs(*shape, device):
    return (torch.rand(*shape, device=device) - .5).sign()


_FUSED_LINEAR = {"on": True}
_KERNEL_CONV = {"on": True}


def set_fused_flipout_linear(flag=True):
    """False: FlipoutNormalLinear evaluates the reference's two-contraction composite on CUDA too (A/B comparisons)."""
    _FUSED_LINEAR["on"] = bool(flag)


def set_flipout_conv_kernels(flag=True):
    """False: FlipOutNormalConv1d / 2d evaluate the reference's composite with torch's convolutions on CUDA too."""
    _KERNEL_CONV["on"] = bool(flag)


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

    def _kernel_path(self, x, nd):
        return (_KERNEL_CONV["on"] and nd <= 2 and x.is_cuda and x.dtype == torch.float32 and x.dim() == nd + 2
                and not self.transposed and conv_implicit_eligible(self.in_channels, self.groups))

    def _contract(self, x4, w4, geometry):
        """conv2d(x4, w4) through the implicit-GEMM contraction kernels: the sample-and-contract path with ONE sample and
        injected eps = 0 (W = w4 exactly); autograd returns d x4 and d w4 from the library's input- / weight-gradient
        kernels."""
        from ..functional import DrawSpec, SampledConv2dImplicit
        zero = getattr(self, "_zero_eps", None)
        if zero is None or zero.device != w4.device or zero.numel() != w4.numel():
            zero = torch.zeros(1, w4.numel(), device=w4.device, dtype=torch.float32)
            self._zero_eps = zero
        spec = DrawSpec(0, 0, 0, eps=zero)
        return SampledConv2dImplicit.apply(x4, w4, self.weight.scale.detach().view(w4.shape), None, None, 1, True, spec,
                                           None, runtime.precision(), *geometry)

    def _flipout(self, x, nd, sample):
        _, x, _ = runtime.mc_expand_rows(x)
        if sample:
            self.sample(x.size(0), (1,) * nd)
        if self._kernel_path(x, nd):
            xs = x * self.S.expand_as(x)
            mean, stddev = self.weight.mean, self.weight.stddev
            if nd == 1:            # a 2-d convolution with a height of one
                shape4 = (mean.shape[0], mean.shape[1], 1, mean.shape[2])
                geometry = ((1, tuple(self.stride)[0]), (0, tuple(self.padding)[0]), (1, tuple(self.dilation)[0]))
                x, xs, mean, stddev = x.unsqueeze(2), xs.unsqueeze(2), mean.view(shape4), stddev.view(shape4)
            else:
                geometry = (tuple(self.stride), tuple(self.padding), tuple(self.dilation))
            out = self._contract(x, mean, geometry)
            pert = self._contract(xs, stddev, geometry)
            if nd == 1:
                out, pert = out.squeeze(2), pert.squeeze(2)
            return out + pert * self.R.expand_as(out)
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
