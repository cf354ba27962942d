"""The reference's layers that are NOT on the variational hot path, provided so that `pytorch_bayesian.nn` can be
aliased to this package as a whole (out of scope as kernels; plain torch, CPU or CUDA):

  MCDropoutLinear / MCDropoutConv{Nd,1d,2d,3d}   deterministic torch layer + dropout that stays on while `sample` is
                                                  true (dense.py:165-179, conv.py:254-326)
  NormalInverseGaussianLinear                     evidential-regression head: four softplus-constrained outputs
                                                  (dense.py:141-162)
  NormalInverseGaussianLoss / ...Uncertainty      its loss and the aleatoric / epistemic split (loss.py:54-79)

They hold no `WeightNormal`, so KLDivergence / PruneNormal fail on them exactly as the reference does.
"""
import math

import torch
import torch.nn.functional as F
from torch.distributions.normal import Normal
from torch.nn import Module

from .container import BayesianModule


def _positive(t):
    return 1e-10 + F.softplus(t)


class NormalInverseGaussianLinear(BayesianModule):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, None)
        self.linear = torch.nn.Linear(in_features, 4 * out_features, bias)

    def forward(self, x, sample=False):
        gamma, upsilon, alpha, beta = self.linear(x).split(self.out_channels, dim=-1)
        upsilon, alpha, beta = _positive(upsilon), 1 + _positive(alpha), _positive(beta)
        if not sample:
            return (gamma, upsilon, alpha, beta)
        return Normal(gamma.clone(), torch.sqrt(beta / (upsilon * (alpha - 1))))


class MCDropoutLinear(BayesianModule):
    def __init__(self, in_features, out_features, bias=True, drop_prob=0.5):
        super().__init__(in_features, out_features, None)
        self.drop_prob = drop_prob
        self.linear = torch.nn.Linear(in_features, out_features, bias)

    def forward(self, x, sample=True):
        return F.dropout(self.linear(x), self.drop_prob, sample, False)


class MCDropoutConvNd(BayesianModule):
    def __init__(self, in_channels, out_channels, drop_prob):
        super().__init__(in_channels, out_channels, None)
        self.drop_prob = drop_prob


def _mc_dropout_conv(name, conv_cls):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 drop_prob=0.5):
        MCDropoutConvNd.__init__(self, in_channels, out_channels, drop_prob)
        self.conv = conv_cls(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.weight = self.conv.weight            # plain Parameters, as in the reference (no `.dist`)
        self.bias = self.conv.bias

    def forward(self, x, sample=True):
        return F.dropout(self.conv(x), self.drop_prob, sample, False)

    return type(name, (MCDropoutConvNd,), {"__init__": __init__, "forward": forward, "__module__": __name__})


MCDropoutConv1d = _mc_dropout_conv("MCDropoutConv1d", torch.nn.Conv1d)
MCDropoutConv2d = _mc_dropout_conv("MCDropoutConv2d", torch.nn.Conv2d)
MCDropoutConv3d = _mc_dropout_conv("MCDropoutConv3d", torch.nn.Conv3d)


class NormalInverseGaussianLoss(Module):
    def __init__(self, reg_lambda=1e-2):
        super().__init__()
        self.reg_lambda = reg_lambda

    def nll(self, y, gamma, upsilon, alpha, beta):
        omega = 2 * beta * (1 + upsilon)
        return (0.5 * torch.log(math.pi / upsilon) - alpha * torch.log(omega)
                + (alpha + 0.5) * torch.log(upsilon * (y - gamma) ** 2 + omega)
                + torch.lgamma(alpha) - torch.lgamma(alpha + 0.5))

    def forward(self, gamma, upsilon, alpha, beta, y):
        regularizer = ((y - gamma).abs() * (2 * upsilon + alpha)).mean()
        return self.nll(y, gamma, upsilon, alpha, beta).mean() + self.reg_lambda * regularizer


class NormalInverseGaussianUncertainty(Module):
    def forward(self, upsilon, alpha, beta):
        aleatoric = beta / (alpha - 1)
        return (aleatoric, aleatoric / upsilon)
