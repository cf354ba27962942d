"""BayesianModule / BayesianNetworkModule (mirror of pytorch_bayesian/nn/container.py:6-37).

BayesianNetworkModule.forward keeps the reference contract — a Python list of S predictions, or the
bare tensor for S == 1 (utils.py:10-11) — but, where it is provably equivalent, runs `_forward`
ONCE for all S Monte-Carlo samples: the deterministic trunk in front of the first Bayesian layer is
evaluated once on the B input rows (the reference recomputes it S times with identical results),
the first Bayesian layer expands to S*B rows with a different eps stream per sample, and every
later layer processes the S independent row blocks in one launch.
"""
import torch
from torch.nn import Module

from .. import runtime
from ..utils.traversal import _item_or_list, traverse
from .elbo import MCSamples

# leaf modules that act row by row along dim 0 (safe to see S*B rows instead of S passes of B rows)
_ROWWISE = [
    torch.nn.Linear, torch.nn.Conv1d, torch.nn.Conv2d, torch.nn.Conv3d, torch.nn.Flatten, torch.nn.Identity,
    torch.nn.ELU, torch.nn.ReLU, torch.nn.ReLU6, torch.nn.LeakyReLU, torch.nn.GELU, torch.nn.SiLU, torch.nn.Tanh,
    torch.nn.Sigmoid, torch.nn.Softplus, torch.nn.SELU, torch.nn.CELU, torch.nn.Hardtanh, torch.nn.PReLU,
    torch.nn.MaxPool1d, torch.nn.MaxPool2d, torch.nn.MaxPool3d, torch.nn.AvgPool1d, torch.nn.AvgPool2d,
    torch.nn.AvgPool3d, torch.nn.AdaptiveAvgPool1d, torch.nn.AdaptiveAvgPool2d, torch.nn.AdaptiveMaxPool2d,
    torch.nn.LayerNorm, torch.nn.GroupNorm, torch.nn.Unflatten, torch.nn.ZeroPad2d,
]
_BATCHNORM = (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)


def register_rowwise_module(cls):
    """Declare a user module class as acting independently on every row of dim 0 (e.g. a custom
    Flatten), so that networks containing it qualify for the batched Monte-Carlo forward."""
    if cls not in _ROWWISE:
        _ROWWISE.append(cls)
    return cls


class BayesianModule(Module):
    """container.py:6-14 — stores the priors and the channel counts."""

    def __init__(self, in_channels, out_channels, prior, bias_prior=None):
        super(BayesianModule, self).__init__()
        self.weight_prior = prior
        self.bias_prior = bias_prior if bias_prior else prior
        self.in_channels = in_channels
        self.out_channels = out_channels


class BayesianNetworkModule(Module):
    """container.py:17-37."""

    def __init__(self, in_channels, out_channels, samples=10):
        super(BayesianNetworkModule, self).__init__()
        self.samples = samples
        self.in_channels = in_channels
        self.out_channels = out_channels

    def _forward(self, x, *args, **kwargs):
        raise NotImplementedError('self._forward() not implemented')

    def traverse(self, fn, *args, **kwargs):
        return traverse(self, fn, *args, **kwargs)

    # ------------------------------------------------------------------ batched Monte-Carlo forward
    def _mc_plan(self):
        """(foldable, [BatchNorm modules that update running statistics]) from the module tree."""
        from .layers import _FusedBayesianLayer
        from .mvn import WeightMultivariateNormal
        from .variational import WeightNormal
        key = (self.training, len(_ROWWISE))
        cached = self.__dict__.get('_mc_plan_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        ok, bns, n_bayes = True, [], 0
        rowwise = tuple(_ROWWISE)
        for m in self.modules():
            if m is self or isinstance(m, (WeightNormal, WeightMultivariateNormal, torch.nn.Sequential,
                                           torch.nn.ModuleList, torch.nn.ModuleDict)):
                continue
            if (isinstance(m, _FusedBayesianLayer) and m._fused) or getattr(m, '_mc_composite', False):
                n_bayes += 1        # fused sample-and-contract layers and the torch composites (Flipout, full covariance)
            elif isinstance(m, _BATCHNORM):
                if m.training and m.track_running_stats:
                    ok = ok and m.momentum is not None
                    bns.append(m)
            elif isinstance(m, (torch.nn.Softmax, torch.nn.LogSoftmax)):
                ok = ok and m.dim not in (0, None)
            elif isinstance(m, (torch.nn.Dropout, torch.nn.Dropout1d, torch.nn.Dropout2d, torch.nn.Dropout3d,
                                torch.nn.AlphaDropout)):
                ok = ok and not m.training      # a shared trunk would reuse one mask for all samples
            elif not isinstance(m, rowwise):
                ok = False                      # unknown code (incl. composite modules and nested networks)
        plan = (ok and n_bayes > 0, bns)
        self.__dict__['_mc_plan_cache'] = (key, plan)
        return plan

    def _forward_batched(self, x, samples, args, kwargs):
        foldable, bns = self._mc_plan()
        if not foldable or not torch.is_tensor(x) or x.dim() < 1 or not x.is_cuda:
            return None
        rank, world = runtime.sample_partition()
        if samples % world != 0:
            raise ValueError(f"{samples} Monte-Carlo samples do not split over {world} sample-parallel ranks")
        local = samples // world
        rows = x.shape[0]
        saved = [(bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone(), bn.momentum)
                 for bn in bns]
        ctx = runtime.MCContext(local, rows, sample_offset=rank * local, total_samples=samples)
        # BatchNorm running statistics: the reference updates them once per MC pass with identical batch
        # statistics, i.e. S momentum steps r <- (1-m) r + m stat; that equals ONE step with momentum
        # 1 - (1-m)^S, which the single batched pass uses (num_batches_tracked advances by S).
        hooks = []
        for bn in bns:
            hooks.append(bn.register_forward_pre_hook(_bn_guard))
            bn.momentum = 1.0 - (1.0 - bn.momentum) ** samples
        try:
            with runtime.mc_batch(ctx):
                out = self._forward(x, *args, **kwargs)
            if not (torch.is_tensor(out) and ctx.expanded and out.dim() >= 1 and out.shape[0] == local * rows):
                raise _NotRowwise()
        except _NotRowwise:
            # a training-mode BatchNorm after the first Bayesian layer (per-sample batch statistics), or a
            # result that is not [S*B, ...]: not equivalent to S passes -> undo and use the reference loop
            with torch.no_grad():
                for bn, (m0, v0, n0, _) in zip(bns, saved):
                    bn.running_mean.copy_(m0), bn.running_var.copy_(v0), bn.num_batches_tracked.copy_(n0)
            self.__dict__['_mc_plan_cache'] = ((self.training, len(_ROWWISE)), (False, []))
            return None
        finally:
            for h in hooks:
                h.remove()
            for bn, sv in zip(bns, saved):
                bn.momentum = sv[3]
        if samples > 1:
            with torch.no_grad():
                for bn in bns:
                    bn.num_batches_tracked += samples - 1
        result = MCSamples(out.view((local, rows) + tuple(out.shape[1:])).unbind(0))
        result.batched = out           # nn.mc_mean_loss evaluates a row-mean criterion on it in one call
        return result

    def forward(self, x, samples=None, *args, **kwargs):
        if samples is None:
            samples = self.samples
        mode = runtime.mc_batching()
        if samples > 1 and mode != 'never' and runtime.current_mc() is None:
            outs = self._forward_batched(x, samples, args, kwargs)
            if outs is not None:
                return _item_or_list(outs)
            if mode == 'always':
                raise RuntimeError("set_mc_batching('always'): this network does not qualify for the batched forward")
        # the reference loop (container.py:36-37); under sample parallelism rank r takes the global
        # sample indices r, r + world, ... of this pass
        rank, world = runtime.sample_partition()
        if world == 1 or runtime.current_mc() is not None:
            return _item_or_list([self._forward(x, *args, **kwargs) for _ in range(samples)])
        if samples % world != 0:
            raise ValueError(f"{samples} Monte-Carlo samples do not split over {world} sample-parallel ranks")
        outs = []
        for _ in range(samples // world):
            rows = x.shape[0] if torch.is_tensor(x) and x.dim() >= 1 else 0
            with runtime.mc_batch(runtime.MCContext(1, rows, sample_offset=rank, total_samples=world)):
                outs.append(self._forward(x, *args, **kwargs))
        return _item_or_list(outs)


class _NotRowwise(Exception):
    pass


def _bn_guard(module, inputs):
    ctx = runtime.current_mc()
    if ctx is not None and ctx.expanded and ctx.samples > 1 and module.training:
        raise _NotRowwise()
