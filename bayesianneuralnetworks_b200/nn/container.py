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
    def _mc_key(self):
        """Everything the plan depends on: the training flag of every submodule (a per-submodule .eval() / .train(),
        e.g. an MC-dropout toggle, changes it) and the registered row-wise classes."""
        return (tuple(m.training for m in self.modules()), len(_ROWWISE))

    def _mc_plan(self):
        """(foldable, [BatchNorm modules using batch statistics], [unknown leaf modules to probe], verified) from the
        module tree."""
        from .layers import _FusedBayesianLayer
        from .mvn import WeightMultivariateNormal
        from .variational import WeightNormal
        key = self._mc_key()
        cached = self.__dict__.get('_mc_plan_cache')
        if cached is not None and cached[0] == key:
            return cached[1]
        ok, bns, probes, n_bayes = True, [], [], 0
        rowwise = tuple(_ROWWISE)
        for m in self.modules():
            if m is self or isinstance(m, (WeightNormal, WeightMultivariateNormal, torch.nn.Sequential,
                                           torch.nn.ModuleList, torch.nn.ModuleDict)):
                continue
            if (isinstance(m, _FusedBayesianLayer) and m._fused) or getattr(m, '_mc_composite', False):
                n_bayes += 1        # fused sample-and-contract layers and the torch composites (Flipout, full covariance)
            elif isinstance(m, _BATCHNORM):
                if m.training or not m.track_running_stats:        # batch statistics: pooled over S*B rows after expansion
                    if m.training and m.track_running_stats:
                        ok = ok and m.momentum is not None
                    bns.append(m)
            elif isinstance(m, (torch.nn.Softmax, torch.nn.LogSoftmax)):
                ok = ok and m.dim not in (0, None)
            elif isinstance(m, (torch.nn.Dropout, torch.nn.Dropout1d, torch.nn.Dropout2d, torch.nn.Dropout3d,
                                torch.nn.AlphaDropout)):
                ok = ok and not m.training      # a shared trunk would reuse one mask for all samples
            elif isinstance(m, rowwise):
                pass
            elif (next(m.children(), None) is None and next(m.parameters(recurse=False), None) is None
                  and next(m.buffers(recurse=False), None) is None):
                probes.append(m)                # unknown stateless leaf (e.g. the examples' own Flatten): probed at run time
            else:
                ok = False                      # unknown code with state or children (composite modules, nested networks)
        plan = _McPlan(ok and n_bayes > 0, bns, probes)
        self.__dict__['_mc_plan_cache'] = (key, plan)
        return plan

    def _disable_batching(self, why):
        self.__dict__['_mc_plan_cache'] = (self._mc_key(), _McPlan(False, [], [], why))

    def _forward_batched(self, x, samples, args, kwargs):
        plan = self._mc_plan()
        if not plan.ok or not torch.is_tensor(x) or x.dim() < 1 or not x.is_cuda:
            return None
        bns = plan.bns
        rank, world = runtime.sample_partition()
        if samples % world != 0:
            raise ValueError(f"{samples} Monte-Carlo samples do not split over {world} sample-parallel ranks")
        local = samples // world
        rows = x.shape[0]
        tracked = [bn for bn in bns if bn.training and bn.track_running_stats]
        saved = [(bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone(), bn.momentum)
                 for bn in tracked]
        draws = _draw_counters(self)
        ctx = runtime.MCContext(local, rows, sample_offset=rank * local, total_samples=samples)
        # BatchNorm running statistics: the reference updates them once per MC pass with identical batch
        # statistics, i.e. S momentum steps r <- (1-m) r + m stat; that equals ONE step with momentum
        # 1 - (1-m)^S, which the single batched pass uses (num_batches_tracked advances by S).
        hooks = [bn.register_forward_pre_hook(_bn_guard) for bn in bns]
        hooks += [m.register_forward_hook(_probe_rowwise) for m in plan.probes]
        for bn in tracked:
            bn.momentum = 1.0 - (1.0 - bn.momentum) ** samples

        def undo(why):
            with torch.no_grad():
                for bn, (m0, v0, n0, _) in zip(tracked, saved):
                    bn.running_mean.copy_(m0), bn.running_var.copy_(v0), bn.num_batches_tracked.copy_(n0)
            for w, state in draws:
                w._draw, w._last = state
            self._disable_batching(why)

        try:
            with runtime.mc_batch(ctx):
                out = self._forward(x, *args, **kwargs)
            if not (torch.is_tensor(out) and ctx.expanded and out.dim() >= 1 and out.shape[0] == local * rows):
                raise _NotRowwise("the result is not a tensor of S*B rows")
        except Exception as exc:      # noqa: BLE001
            # `_forward` is arbitrary user code: whatever goes wrong in the S*B-row attempt (a training-mode BatchNorm
            # after the first Bayesian layer, a skip connection adding B rows to S*B rows, two Bayesian branches fed the
            # same input, ...) the reference loop below still evaluates the model the reference's way
            undo(f"{type(exc).__name__}: {exc}")
            return None
        finally:
            for h in hooks:
                h.remove()
            for bn, sv in zip(tracked, saved):
                bn.momentum = sv[3]
        if (not plan.verified and not torch.cuda.is_current_stream_capturing()
                and getattr(runtime._tls, "eps", None) is None):      # (injected eps tensors are sized for S samples)
            # once per plan: ONE reference-loop pass with the first sample's draws must reproduce the first row block
            # (catches functional code in `_forward` that mixes rows, which no module inspection can see)
            if not self._verify_batched(x, out, rows, draws, tracked, args, kwargs):
                undo("the batched pass does not reproduce a reference-loop pass (row-mixing code in _forward?)")
                return None
            plan.verified = True
        if samples > 1:
            with torch.no_grad():
                for bn in tracked:
                    bn.num_batches_tracked += samples - 1
        result = MCSamples(out.view((local, rows) + tuple(out.shape[1:])).unbind(0))
        result.batched = out           # nn.mc_mean_loss evaluates a row-mean criterion on it in one call
        return result

    def _verify_batched(self, x, out, rows, draws_before, tracked, args, kwargs):
        """Re-evaluates Monte-Carlo sample 0 of the batched pass the reference's way (one `_forward` over the B rows,
        same Philox draw indices, hence the same eps) and compares it with the first row block of `out`."""
        rank, world = runtime.sample_partition()
        after = [(w, (w._draw, w._last)) for w, _ in draws_before]
        composites = [m for m in self.modules() if getattr(m, '_mc_composite', False)]
        if composites:
            return True         # torch-RNG composites (Flipout signs, full covariance) cannot replay their draws
        # the check pass must not touch the running statistics IN PLACE: the batched pass saved them for its backward
        # (an in-place write would invalidate that graph) — it runs on throw-away copies of the buffers instead
        real = [(bn, bn.running_mean, bn.running_var, bn.num_batches_tracked) for bn in tracked]
        for bn, m0, v0, n0 in real:
            bn.running_mean, bn.running_var, bn.num_batches_tracked = m0.clone(), v0.clone(), n0.clone()
        for w, state in draws_before:
            w._draw, w._last = state
        try:
            with torch.no_grad(), runtime.mc_batch(runtime.MCContext(1, rows, sample_offset=(out.shape[0] // rows) * rank,
                                                                       total_samples=1)):
                one = self._forward(x, *args, **kwargs)
            tol = 1e-4 if runtime.precision_name() == "fp32" else 1e-2
            first = out[:rows].detach()
            same = (torch.is_tensor(one) and one.shape == first.shape
                    and bool(((one - first).abs().max() <= tol * first.abs().max().clamp_min(1e-30)).item()))
        except Exception:      # noqa: BLE001
            same = False
        finally:
            for w, state in after:
                w._draw, w._last = state
            for bn, m0, v0, n0 in real:
                bn.running_mean, bn.running_var, bn.num_batches_tracked = m0, v0, n0
        return same

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


class _McPlan:
    """What `_mc_plan` found: `ok` (the module tree qualifies for the batched pass), the BatchNorm modules that use batch
    statistics, the unknown stateless leaf modules probed at run time, whether one batched pass has been verified
    against a reference-loop pass, and — when batching was switched off at run time — why.  Indexable as
    (ok, bns) for brevity."""
    __slots__ = ("ok", "bns", "probes", "verified", "why")

    def __init__(self, ok, bns, probes, why=None):
        self.ok, self.bns, self.probes, self.verified, self.why = ok, bns, probes, False, why

    def __getitem__(self, i):
        return (self.ok, self.bns)[i]


def _draw_counters(model):
    """[(WeightNormal, (draws consumed, last draw))] of every variational tensor below `model`."""
    from .variational import WeightNormal
    return [(w, (w._draw, w._last)) for w in model.modules() if isinstance(w, WeightNormal)]


def _bn_guard(module, inputs):
    """A BatchNorm that normalises with BATCH statistics (training mode, or no running statistics) after the rows were
    expanded would pool the statistics of all S samples: not what S separate passes compute."""
    ctx = runtime.current_mc()
    if ctx is not None and ctx.expanded and ctx.samples > 1:
        raise _NotRowwise("BatchNorm with batch statistics after the first Bayesian layer")


def _probe_rowwise(module, inputs, output):
    """Forward hook on an unknown stateless leaf module during a batched pass: it qualifies when it keeps the leading
    dimension and evaluating the first half of the rows alone gives the first half of the result, bit for bit."""
    ok = module.__dict__.get('_bnn_rowwise')
    if ok is None:
        x = inputs[0] if inputs else None
        ok = False
        if (len(inputs) == 1 and torch.is_tensor(x) and torch.is_tensor(output) and x.dim() >= 1 and output.dim() >= 1
                and output.shape[0] == x.shape[0]):
            if x.shape[0] < 2 or torch.cuda.is_current_stream_capturing():
                return              # nothing to compare / cannot synchronise: decided by a later call
            half = x.shape[0] // 2
            with torch.no_grad():
                part = module.forward(x[:half])
            ok = torch.is_tensor(part) and part.shape == output[:half].shape and torch.equal(part, output[:half].detach())
        module.__dict__['_bnn_rowwise'] = ok
    if not ok:
        raise _NotRowwise(f"{type(module).__name__} does not act row by row")
