"""Full-covariance Gaussian layers — SURVEY §8(f) rank 4: `WeightMultivariateNormal` (pytorch_bayesian/nn/core.py:48-92)
and `MultivariateNormalLinear` (dense.py:86-138), the last layer of examples/CIFAR10/model.py.

Not on the `mu + softplus(rho) * eps` path (a lower-triangular scale per output row, a Cholesky-based KL): torch
composites, kept so that the example model definitions run unchanged.  Reference quirks preserved on purpose: the
noise is UNIFORM (`torch.rand_like`, core.py:91), `stddev` is the elementwise square root of the triangular matrix
(core.py:68-69), the upper triangle of `scale` is initialised to -100 (dense.py:106-109).
"""
import math

import torch
from torch.distributions import MultivariateNormal
from torch.nn import Module, init
from torch.nn.parameter import Parameter

from .. import runtime
from .layers import BayesianLinear


class WeightMultivariateNormal(Module):

    def __init__(self, *channels):
        super(WeightMultivariateNormal, self).__init__()
        self.mean = Parameter(torch.empty(*channels))
        self.scale = Parameter(torch.eye(channels[-1]).repeat(*channels[:-1], 1, 1))
        self.sample()

    @property
    def device(self):
        return self.mean.device

    @property
    def requires_grad(self):
        return self.mean.requires_grad

    @property
    def variance(self):
        tril = torch.tril(torch.nn.functional.softplus(self.scale))
        return torch.eye(self.size(-1), device=self.device) * 1e-10 + tril

    @property
    def stddev(self):
        return self.variance.sqrt()

    @property
    def dist(self):
        return MultivariateNormal(self.mean, scale_tril=self.variance)

    @property
    def shape(self):
        return self.size()

    def size(self, *dims):
        return self.mean.size(*dims)

    def sample(self):
        noise = torch.rand_like(self.mean).unsqueeze(-1)
        self.sampled = self.mean + torch.matmul(self.stddev, noise).squeeze(-1)

    def sample_many(self, count):
        """[count, *shape]: `count` independent draws of sample() in one call; `.sampled` keeps the last one."""
        noise = torch.rand((count,) + tuple(self.mean.shape), device=self.device, dtype=self.mean.dtype).unsqueeze(-1)
        draws = self.mean + torch.matmul(self.stddev, noise).squeeze(-1)
        self.sampled = draws[-1]
        return draws


class MultivariateNormalLinear(BayesianLinear):
    _mc_composite = True      # takes part in the batched Monte-Carlo forward (S weight draws, one batched matmul)

    def __init__(self, in_features, out_features, bias=True, weight_prior=None, bias_prior=None):
        if not weight_prior:
            weight_prior = MultivariateNormal(torch.zeros(out_features, in_features),
                                              torch.eye(in_features).repeat(out_features, 1, 1))
        if bias and not bias_prior:
            bias_prior = MultivariateNormal(torch.zeros(out_features), torch.eye(out_features))
        super(MultivariateNormalLinear, self).__init__(in_features, out_features, bias, WeightMultivariateNormal,
                                                       weight_prior, bias_prior)

    @staticmethod
    def _mask_upper(scale):
        with torch.no_grad():
            scale[torch.triu(torch.ones_like(scale), 1).to(torch.bool)] = -100

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight.mean, a=math.sqrt(5))
        init.normal_(self.weight.scale, -2.0, 0.15)
        self._mask_upper(self.weight.scale)
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight.mean)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias.mean, -bound, bound)
            init.normal_(self.bias.scale, -2.0, 0.15)
            self._mask_upper(self.bias.scale)
        self.sample()

    def sample(self):
        self.weight.sample()
        if self.bias is not None:
            self.bias.sample()
        self.sampled = (self.weight.sampled, self.bias.sampled if self.bias is not None else None)

    def forward(self, x, sample=True):
        S, x, ctx = runtime.mc_expand_rows(x)
        if ctx is not None and S > 1 and sample:
            w = self.weight.sample_many(S)                                   # [S, out, in]
            b = self.bias.sample_many(S) if self.bias is not None else None  # [S, out]
            self.sampled = (self.weight.sampled, self.bias.sampled if self.bias is not None else None)
            x3 = x.reshape(S, -1, x.shape[-1])
            y = torch.matmul(x3, w.transpose(1, 2))
            if b is not None:
                y = y + b.unsqueeze(1)
            return y.reshape(x.shape[:-1] + (w.shape[1],))
        if sample:
            self.sample()
        return torch.nn.functional.linear(x, *self.sampled)
