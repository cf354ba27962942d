"""WeightNormal: the factorised-Gaussian variational tensor (mirror of pytorch_bayesian/nn/core.py:7-45).

`mean` and `scale` (= rho) are the Parameters, so state_dict keys stay `...weight.mean`,
`...weight.scale` (SURVEY §5 checkpoint).  Sampling does not materialise anything: `sample()`
advances this tensor's draw counter in its Philox stream, the contraction kernels generate
`mean + stddev * eps` on the fly, and `.sampled` materialises the last draw on demand (CUDA only).
"""
import torch
from torch.distributions import Normal
from torch.nn import Module
from torch.nn.parameter import Parameter

from .. import _C, runtime
from ..functional import DrawSpec, Materialize


class WeightNormal(Module):
    # (Cout, Cg, taps) for conv weights whose eps stream is keyed by the element index in (o, kh, kw, c) order — the order
    # the implicit-GEMM conv kernels walk the tensor in (functional.SampledConv2dImplicit); None: storage order.  eps is
    # i.i.d., so this only decides WHICH standard-normal draw meets which weight; `.sampled`, the kernels and injected eps
    # (given in storage order, like the reference's randn_like tensors) all go through the same mapping.
    _eps_layout = None

    def __init__(self, *channels):
        super(WeightNormal, self).__init__()
        self.mean = Parameter(torch.empty(*channels))
        self.scale = Parameter(torch.empty(*channels))
        # Philox stream identity and draw bookkeeping: plain attributes, not buffers, so that the
        # state_dict key set equals the reference's
        self._tensor_id = runtime.next_tensor_id()
        self._draw = 0            # draws consumed so far
        self._last = None         # (first draw index, count) of the most recent sample()
        self.sample()             # core.py:15 (a counter bump here: construction happens on the CPU)

    # ---- core.py:17-42 -------------------------------------------------------------------------
    @property
    def device(self):
        return self.mean.device

    @property
    def requires_grad(self):
        return self.mean.requires_grad

    @property
    def stddev(self):
        """1e-10 + softplus(scale) — core.py:25-27 (autograd-tracked, for user code; the kernels
        compute it themselves)."""
        return 1e-10 + torch.nn.functional.softplus(self.scale)

    @property
    def variance(self):
        return self.stddev.pow(2)

    @property
    def dist(self):
        return Normal(self.mean, self.stddev)

    @property
    def shape(self):
        return self.size()

    def size(self, *dims):
        return self.mean.size(*dims)

    # ---- sampling ------------------------------------------------------------------------------
    def advance(self, count=1, offset=0, total=None):
        """Reserve `total` (default `count`) draws; this process uses [begin+offset, begin+offset+count)."""
        begin = self._draw
        self._draw += count if total is None else total
        self._last = (begin + offset, count)
        return begin + offset

    def draw_spec(self, begin, count):
        eps = runtime.injected_for(self)
        if eps is not None:
            if eps.shape[0] != count or eps[0].numel() != self.mean.numel():
                raise ValueError(f"injected eps must have shape [{count}, {tuple(self.mean.shape)}], got "
                                 f"{tuple(eps.shape)}")
            eps = eps.to(device=self.mean.device, dtype=torch.float32).contiguous()
            if self._eps_layout is not None:
                O, Cg, taps = self._eps_layout
                eps = eps.reshape(count, O, Cg, taps).transpose(2, 3).contiguous()
        return DrawSpec(runtime.seed(), self._tensor_id, begin, eps, runtime.step_counter(self.mean.device)
                        if self.mean.is_cuda else None)

    def sample(self):
        """core.py:44-45 — draws a fresh eps (by advancing the counter)."""
        self.advance(1)

    def materialize(self, begin=None, count=None):
        """[count, *shape] sampled tensors of draws [begin, begin+count) (default: the last draw)."""
        if begin is None:
            begin, count = self._last
        _C.require_cuda(self.mean)
        if self._eps_layout is not None:
            O, Cg, taps = self._eps_layout
            def perm(t):
                return t.reshape(O, Cg, taps).transpose(1, 2).contiguous()
            out = Materialize.apply(perm(self.mean), perm(self.scale), count, self.draw_spec(begin, count))
            return out.reshape(count, O, taps, Cg).transpose(2, 3).reshape((count,) + tuple(self.mean.shape))
        return Materialize.apply(self.mean, self.scale, count, self.draw_spec(begin, count))

    @property
    def sampled(self):
        """mean + stddev * eps of the most recent draw (of its last MC sample when the draw was a
        batch), autograd-tracked like the reference's attribute."""
        begin, count = self._last
        return self.materialize(begin + count - 1, 1)[0]
