"""The likelihood term of the ELBO over S Monte-Carlo predictions (additive API).

The reference's loop body averages a row-mean criterion over the list of S predictions,
`torch.stack([criterion(pred, y) for pred in preds]).mean()` (examples/MNIST/train.py:59-61): 3 S launches forward and
as many backward for S tiny tensors.  When the predictions are the row blocks of ONE [S*B, ...] tensor — which is what
the batched Monte-Carlo forward of `BayesianNetworkModule` produces — the mean of the S block means is the mean over
all S*B rows, i.e. one criterion call.  `mc_mean_loss` does that and falls back to the reference loop otherwise.  Plain
cross-entropy on CUDA (the examples' criterion) goes to the fused kernels bnn_mc_cross_entropy_fwd / _bwd.
"""
import torch
import torch.nn.functional as F

_MEAN_FUNCTIONS = (F.cross_entropy, F.nll_loss, F.mse_loss, F.l1_loss, F.binary_cross_entropy,
                   F.binary_cross_entropy_with_logits, F.smooth_l1_loss, F.huber_loss)
_MEAN_MODULES = (torch.nn.CrossEntropyLoss, torch.nn.NLLLoss, torch.nn.MSELoss, torch.nn.L1Loss, torch.nn.BCELoss,
                 torch.nn.BCEWithLogitsLoss, torch.nn.SmoothL1Loss, torch.nn.HuberLoss)


class MCSamples(list):
    """The reference's return value — a Python list of the S per-sample outputs (container.py:36-37) — that also
    remembers the [S*B, ...] tensor its entries are views of (`batched`, sample-major rows)."""

    batched = None


def _row_mean(criterion):
    """True when `criterion(input, target)` is known to average over the rows of dim 0 with equal weight per row
    (or per-row weights that depend on the row's target only, as CrossEntropyLoss(weight=...) does)."""
    if isinstance(criterion, _MEAN_MODULES):
        return getattr(criterion, 'reduction', None) == 'mean'
    return any(criterion is f for f in _MEAN_FUNCTIONS)


def _plain_cross_entropy(criterion):
    """ignore_index when `criterion` is cross-entropy with mean reduction, class-index targets and no class weights or
    label smoothing (the criterion of the reference examples, train.py:40); None otherwise."""
    if criterion is F.cross_entropy:
        return -100
    if (type(criterion) is torch.nn.CrossEntropyLoss and criterion.weight is None and criterion.reduction == 'mean'
            and criterion.label_smoothing == 0.0):
        return criterion.ignore_index
    return None


def mc_mean_loss(criterion, preds, target):
    """mean_s criterion(preds[s], target) for the list of Monte-Carlo predictions `preds` (a bare tensor when S == 1,
    as the reference returns it)."""
    if torch.is_tensor(preds):
        return criterion(preds, target)
    base = getattr(preds, 'batched', None)
    n = len(preds)
    if (base is not None and n > 1 and _row_mean(criterion) and torch.is_tensor(target) and target.dim() >= 1
            and base.shape[0] == n * target.shape[0]):
        ignore_index = _plain_cross_entropy(criterion)
        if (ignore_index is not None and base.is_cuda and base.dtype == torch.float32 and base.dim() == 2
                and target.dim() == 1 and target.dtype == torch.int64 and target.device == base.device):
            from ..functional import MCCrossEntropy       # one fused kernel each way, labels not replicated
            return MCCrossEntropy.apply(base, target, ignore_index)
        return criterion(base, target.repeat((n,) + (1,) * (target.dim() - 1)))
    return torch.stack([criterion(p, target) for p in preds]).mean()
