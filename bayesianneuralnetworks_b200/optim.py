"""ELBOAdam — the optimizer half of the ELBO tail (additive API).

The reference trains with `torch.optim.Adam(model.parameters())` on `loss = likelihood + KLDivergence(model)`
(examples/MNIST/train.py:43,57-65).  The KL term is a closed form of the variational parameters alone, so its gradient
needs no autograd: ELBOAdam applies, in ONE pass over every (mean, scale) pair the KLDivergence of the reference would
visit, `grad = d likelihood + d KL` and Adam's update (bnn_adam_kl_step).  The training loop then back-propagates the
likelihood only:

    opt = bnn.optim.ELBOAdam(model, number_of_batches=len(loader), lr=1e-3)
    preds = model(x)
    likelihood = bnn.nn.mc_mean_loss(criterion, preds, y)
    likelihood.backward()                    # (data parallel: all-reduce the gradients here, the KL part is added after)
    opt.step()                               # same parameters as Adam on likelihood + KLDivergence(n_batches)(model)
    kl = bnn.nn.KLDivergence(len(loader))(model)     # only if the value is wanted for logging (no gradient needed)

Parameters that are not variational tensors seen by `model.traverse` (deterministic layers, Bayesian layers nested in
plain modules — the reference's KL does not see those either) are updated by torch's own fused Adam.
"""
import torch

from . import _C
from .nn.loss import _scalar_prior
from .nn.variational import WeightNormal
from .utils.traversal import apply_wb


class ELBOAdam:
    def __init__(self, model, number_of_batches=1, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        def visit(param, module=None, type=None):          # apply_wb passes module / type by keyword (utils.py:30-47)
            return (param, module.weight_prior if type == 'w' else module.bias_prior)
        found = model.traverse(lambda m: apply_wb(m, visit, pass_module=True, pass_type=True)) or []
        for p, _ in found:
            if not isinstance(p, WeightNormal):
                raise NotImplementedError(f"ELBOAdam: the KL gradient of {p.__class__.__name__} has no fused form; use "
                                          "torch.optim.Adam with KLDivergence in the loss for this model")
        n = len(found)
        self._var = []
        seen = set()
        for w, prior in found:
            loc, scale = _scalar_prior(prior, w.__class__.__name__)
            coeff = 1.0 / (w.mean.numel() * n * number_of_batches)      # loss.py:28,38: mean of means / n_batches
            if id(w.mean) in seen:                                       # a tensor listed twice: its KL counts twice
                for e in self._var:
                    if e["w"] is w:
                        e["coeff"] += coeff
                continue
            seen.update((id(w.mean), id(w.scale)))
            self._var.append({"w": w, "loc": loc, "scale": scale, "coeff": coeff, "state": None})
        other = [p for p in model.parameters() if id(p) not in seen]
        self._other = torch.optim.Adam(other, lr=lr, betas=betas, eps=eps, fused=all(p.is_cuda for p in other),
                                       capturable=capturable) if other else None
        self._step_dev = None
        self._params = [p for e in self._var for p in (e["w"].mean, e["w"].scale)] + other

    @property
    def param_groups(self):
        return self._other.param_groups if self._other is not None else []

    def zero_grad(self, set_to_none=True):
        for p in self._params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_()
                p.grad.zero_()

    def step(self):
        if self._other is not None:
            self._other.step()
        if not self._var:
            return
        dev = self._var[0]["w"].mean.device
        if self._step_dev is None:
            self._step_dev = torch.zeros((), device=dev, dtype=torch.float32)
        self._step_dev += 1                                   # device-resident: a captured graph replays it
        entries = []
        for e in self._var:
            w = e["w"]
            if e["state"] is None:
                e["state"] = tuple(torch.zeros_like(w.mean) for _ in range(4))
            m_mu, v_mu, m_rho, v_rho = e["state"]
            g_mu = None if w.mean.grad is None else w.mean.grad.contiguous()
            g_rho = None if w.scale.grad is None else w.scale.grad.contiguous()
            entries.append((w.mean.data, w.scale.data, g_mu, g_rho, m_mu, v_mu, m_rho, v_rho, e["loc"], e["scale"],
                            e["coeff"]))
        _C.adam_kl_step(entries, self.lr, self.betas[0], self.betas[1], self.eps, step_dev=self._step_dev)
