"""PruneNormal (mirror of pytorch_bayesian/prune/prune.py:5-22).

Per variational tensor, independently: key = log N(0; mean, stddev), the k = int(percentage * numel)
largest keys get mean <- 0, scale <- -30, without autograd.  All tensors of the model go through the
same launches of libbnn_b200's exact select (keys bit-identical to torch's Normal.log_prob on the
device; ties at the k-th key resolved towards the lowest index).

Large tensors take the ONE-SWEEP out-of-place path (bnn_prune_into: 8 bytes read + 8 written per pair
instead of two reads and a write): the pruned values land in fresh tensors whose storage then replaces
the Parameters' (`param.data = out`; the Parameter objects, their `.grad` and every optimizer state keyed
by them stay).  That is observably the reference's in-place update unless other tensors alias the
parameter's storage — views taken earlier, or a captured CUDA graph that bakes the address in; for those
`PruneNormal(in_place=True)` (or `bnn.prune.set_in_place(True)`) keeps the strictly in-place kernel.
"""
import torch

from .. import _C
from ..nn.variational import WeightNormal
from ..utils.traversal import apply_wb


_IN_PLACE = {"always": False}
_SWAP_MIN_NUMEL = 1 << 20       # below this the in-place kernel's second read comes out of L2 anyway


def set_in_place(flag=True):
    """Process-wide default of PruneNormal(in_place=None)."""
    _IN_PLACE["always"] = bool(flag)


class PruneNormal():

    def __init__(self, in_place=None):
        self.in_place = in_place

    def __call__(self, module, percentage=0.5):
        self.prune(module, percentage)

    def prune_param(self, param, percentage):
        """prune.py:10-17 for one tensor."""
        self._launch([param], percentage)

    def kl_and_prune(self, module, percentage=0.5, number_of_batches=1):
        """KLDivergence(number_of_batches)(module) of the model BEFORE pruning and the pruning itself in one sweep over the
        parameters (north_star: "the pruning mask reuses that same pass"): tensors on the one-sweep out-of-place path
        get their KL element sums as a by-product of the key arithmetic (bnn_prune_into, kl_sum_out), the others go through
        bnn_kl.  Same reduction as nn.KLDivergence: mean over tensors of the per-tensor means, divided by
        number_of_batches (loss.py:28,38).  Returns the 0-dim float32 divergence."""
        from ..nn.loss import fused_kl_entries
        with torch.no_grad():
            priors = fused_kl_entries(module)       # [(WeightNormal, loc, scale)] in traversal order
            sums = self._launch([w for w, _, _ in priors], percentage, kl_priors=[(loc, sc) for _, loc, sc in priors])
            means = torch.stack([s / w.mean.numel() for s, (w, _, _) in zip(sums, priors)])
            return (means.mean() / number_of_batches).to(torch.float32)

    def _launch(self, params, percentage, kl_priors=None):
        in_place = _IN_PLACE["always"] if self.in_place is None else self.in_place
        entries, swap, small, order = [], [], [], []
        for p in params:
            if not isinstance(p, WeightNormal):
                raise NotImplementedError(f"PruneNormal: unsupported variational tensor {p.__class__.__name__}")
            _C.require_cuda(p.mean, p.scale)
            n = p.mean.numel()
            k = int(percentage * n)          # prune.py:13 — float32 arithmetic when percentage is a tensor
            if k < 0 or k > n:
                raise RuntimeError(f"selected index k out of range: k={k}, numel={n}")
            if not (p.mean.is_contiguous() and p.scale.is_contiguous()):
                raise ValueError("PruneNormal needs contiguous mean/scale parameters")
            if not in_place and n >= _SWAP_MIN_NUMEL:
                swap.append((p, k, len(order)))
            else:
                entries.append((p.mean.data, p.scale.data, k, None, None))
                small.append(len(order))
            order.append(p)
        sums = [None] * len(order)
        if kl_priors is not None and small:       # KL of the tensors that stay on the in-place kernel: before it runs
            ks = _C.kl([(order[i].mean.data, order[i].scale.data, None, None, kl_priors[i][0], kl_priors[i][1], 1.0)
                        for i in small])
            for j, i in enumerate(small):
                sums[i] = ks[j]
        _C.prune(entries)
        if swap:
            res = _C.prune_into([(p.mean.data, p.scale.data, k, None) for p, k, _ in swap],
                                kl_priors=None if kl_priors is None else [kl_priors[i] for _, _, i in swap])
            outs, ks = res if kl_priors is not None else (res, None)
            for j, ((p, _, i), (mu_out, rho_out)) in enumerate(zip(swap, outs)):
                p.mean.data, p.scale.data = mu_out, rho_out
                if ks is not None:
                    sums[i] = ks[j]
        return sums

    def prune(self, module, percentage=0.5):
        """prune.py:19-22 — every tensor the traversal finds, weights and biases alike."""
        with torch.no_grad():
            found = module.traverse(lambda m: apply_wb(m, lambda p: p))
            if found:
                self._launch(found, percentage)
