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

It is a `torch.optim.Optimizer`: `param_groups[0]` holds the variational tensors, `param_groups[1]` every other
parameter (deterministic layers, Bayesian layers nested in plain modules — the reference's KL does not see those
either); lr / betas / eps are read from the groups at every step, so `for g in opt.param_groups: g['lr'] = ...` and
torch's LR schedulers work; `state_dict()` / `load_state_dict()` carry the moments and the step count.  CUDA float32
parameters of the second group ride in the same kernel launch (plain Adam, `rho == NULL`); anything else goes to
torch's own Adam.  Tensors with `requires_grad=False` are skipped, as torch's Adam skips them.

Multi-GPU: `attach_peers(PeerGradients)` makes `step()` average the gradients over the ranks INSIDE the optimizer kernel
(bnn_adam_kl_step_peers: every rank reads all ranks' flat gradient buffers over NVLink), bracketed by two flag barriers —
no NCCL call, so the whole training step stays one capturable CUDA graph.
"""
import torch

from . import _C
from .nn.loss import _is_scalar_normal, _scalar_prior
from .nn.variational import WeightNormal
from .utils.traversal import apply_wb


class ELBOAdam(torch.optim.Optimizer):
    def __init__(self, model, number_of_batches=1, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        def visit(param, module=None, type=None):          # apply_wb passes module / type by keyword (utils.py:30-47)
            return (param, module.weight_prior if type == 'w' else module.bias_prior)
        found = model.traverse(lambda m: apply_wb(m, visit, pass_module=True, pass_type=True)) or []
        n = len(found)                    # loss.py:38 averages over ALL listed tensors, fused or not
        self._var = []
        self.composite = []               # tensors whose KL has no fused form: their KL gradient must come from autograd
        seen = set()
        for w, prior in found:
            if not (isinstance(w, WeightNormal) and _is_scalar_normal(prior)):
                # full-covariance tensors / tensor-valued priors: plain Adam here; add KLDivergence.composite_part(model)
                # to the loss that is back-propagated (training.ElboTrainer does)
                self.composite.append(w)
                continue
            loc, scale = _scalar_prior(prior, w.__class__.__name__)
            coeff = 1.0 / (w.mean.numel() * n * number_of_batches)      # loss.py:28,38: mean of means / n_batches
            if id(w.mean) in seen:                                       # a tensor listed twice: its KL counts twice
                for e in self._var:
                    if e["w"] is w:
                        e["coeff"] += coeff
                continue
            seen.update((id(w.mean), id(w.scale)))
            self._var.append({"w": w, "loc": loc, "scale": scale, "coeff": coeff})
        other = [p for p in model.parameters() if id(p) not in seen]
        groups = [{"params": [p for e in self._var for p in (e["w"].mean, e["w"].scale)], "variational": True}]
        if other:
            groups.append({"params": other, "variational": False})
        if not groups[0]["params"]:
            groups = groups[1:]
        if not groups:
            raise ValueError("ELBOAdam: the model has no parameters")
        super().__init__(groups, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps)))
        self.capturable = capturable
        self._step_dev = {}               # device -> 0-dim float32 step count (device-resident: graph replays advance it)
        self._torch_adam = None           # for parameters the kernel does not take (non-CUDA / non-float32)
        self._peers = None
        self._views = None
        self.two_hop_above = 2            # peer exchange: ranks above this count average the buffers in two hops first

    # ---------------------------------------------------------------------------------------------------------------
    def attach_peers(self, peers, views=None):
        """`peers`: a parallel.PeerGradients whose flat buffer holds every gradient of this optimizer (see module doc).
        `views` ({id(param): view of peers.flat}): where each parameter's gradient sits in the flat buffer when the
        gradients are gathered into it after the backward pass (training.ElboTrainer) instead of being views of it."""
        self._peers = peers
        self._views = views

    def _moments(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st["exp_avg"], st["exp_avg_sq"]

    def _kernel_ok(self, p):
        return p.is_cuda and p.dtype == torch.float32 and _C._dense(p)

    def _grad(self, p):
        if self._peers is not None and self._views is not None:
            return self._views[id(p)]
        g = p.grad
        if g is None:
            return None
        if g.stride() != p.stride():
            g = g.contiguous() if p.is_contiguous() else torch.empty_like(p).copy_(g)
        return g

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_device = {}            # device -> (hyper-parameters, entries)
        leftovers = []
        for group in self.param_groups:
            hyper = (group["lr"], group["betas"][0], group["betas"][1], group["eps"])
            if group.get("variational"):
                for e in self._var:
                    w = e["w"]
                    if not (w.mean.requires_grad or w.scale.requires_grad):
                        continue                      # frozen tensors are left alone, as torch's Adam leaves them
                    if w.mean.requires_grad != w.scale.requires_grad:
                        raise NotImplementedError("ELBOAdam: freeze mean and scale of a variational tensor together")
                    m_mu, v_mu = self._moments(w.mean)
                    m_rho, v_rho = self._moments(w.scale)
                    by_device.setdefault((w.mean.device, hyper), []).append(
                        (w.mean.data, w.scale.data, self._grad(w.mean), self._grad(w.scale), m_mu, v_mu, m_rho, v_rho,
                         e["loc"], e["scale"], e["coeff"]))
            else:
                for p in group["params"]:
                    if not p.requires_grad or (p.grad is None and not (self._peers is not None and self._views is not None)):
                        continue
                    if self._kernel_ok(p):
                        m, v = self._moments(p)
                        by_device.setdefault((p.device, hyper), []).append(
                            (p.data, None, self._grad(p), None, m, v, None, None, 0.0, 1.0, 0.0))
                    else:
                        leftovers.append((p, group))
        for (device, hyper), entries in by_device.items():
            step_dev = self._step_dev.get(device)
            if step_dev is None:
                step_dev = self._step_dev[device] = torch.zeros((), device=device, dtype=torch.float32)
        for step_dev in self._step_dev.values():
            step_dev += 1                                 # device-resident: a captured graph replays it
        peers = self._peers
        two_hop = peers is not None and peers.world > self.two_hop_above
        if peers is not None:
            peers.barrier()                               # every rank's gradients are complete
        if two_hop:
            # more than two ranks: average the flat buffers in place (every rank reduces its slice and writes it back to
            # all: (R-1)/R of the buffer per direction instead of R-1 whole buffers inbound), then the local kernel
            _C.peer_average(peers.bases, peers.rank, peers.flat.numel(), peers.device)
            peers.barrier()                               # every slice has landed in every buffer
        for (device, hyper), entries in by_device.items():
            _C.adam_kl_step(entries, *hyper, step_dev=self._step_dev[device],
                            peers=None if (peers is None or two_hop) else (peers.rank, peers.bases))
        if peers is not None and not two_hop:
            peers.barrier()                               # every rank has read: the buffers may be overwritten
        if leftovers:
            self._step_leftovers(leftovers)
        return loss

    def _step_leftovers(self, leftovers):
        if self._peers is not None:
            raise NotImplementedError("ELBOAdam with peer gradients needs CUDA float32 parameters")
        if self._torch_adam is None:
            groups = {}
            for p, g in leftovers:
                groups.setdefault(id(g), (g, []))[1].append(p)
            self._torch_adam = torch.optim.Adam([{"params": ps, "lr": g["lr"], "betas": g["betas"], "eps": g["eps"]}
                                                 for g, ps in groups.values()], capturable=self.capturable)
            self._torch_adam_groups = [g for g, _ in groups.values()]
        for tg, g in zip(self._torch_adam.param_groups, self._torch_adam_groups):
            tg["lr"], tg["betas"], tg["eps"] = g["lr"], g["betas"], g["eps"]
        self._torch_adam.step()

    # ---------------------------------------------------------------------------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        steps = [float(s) for s in self._step_dev.values()]
        sd["bnn_step"] = max(steps) if steps else 0.0
        if self._torch_adam is not None:
            sd["bnn_torch_adam"] = self._torch_adam.state_dict()
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        step = state_dict.pop("bnn_step", 0.0)
        leftover = state_dict.pop("bnn_torch_adam", None)
        super().load_state_dict(state_dict)
        devices = {p.device for g in self.param_groups for p in g["params"] if p.is_cuda}
        self._step_dev = {d: torch.full((), float(step), device=d, dtype=torch.float32) for d in devices}
        if leftover is not None and self._torch_adam is not None:
            self._torch_adam.load_state_dict(leftover)
