"""Host-side partitioning for one-process-per-GPU runs.  The path shards by independent
units — batch rows, Monte-Carlo samples, variational tensors — so the only exchange steps are one
all-reduce of the flat gradient buffer per training step and one scalar (or small vector) all-reduce
for a tensor-sharded KL.  Collectives go through torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def sample_range(total_samples, rank, world):
    """Global MC sample indices [begin, end) of `rank` (contiguous, equal blocks)."""
    if total_samples % world != 0:
        raise ValueError(f"{total_samples} Monte-Carlo samples do not split over {world} ranks")
    per = total_samples // world
    return rank * per, (rank + 1) * per


def grid_coordinates(rank, world, sample_groups):
    """rank -> (data index, sample-group index) for a data x sample grid (sample index fastest)."""
    if world % sample_groups != 0:
        raise ValueError(f"{sample_groups} sample groups do not divide {world} ranks")
    return rank // sample_groups, rank % sample_groups


def round_robin(items, rank, world):
    """The items (variational tensors) owned by `rank` when sharded round-robin (KL / prune sweeps)."""
    return [it for i, it in enumerate(items) if i % world == rank]


def allreduce_gradients(params, group=None, average=True):
    """ONE all-reduce over the flat buffer of every parameter gradient (missing gradients count as zeros
    so that all ranks reduce the same layout).  Equal-sized shards => plain average (train.py:59-61 is a
    mean over samples and rows)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    for p, g, f in zip(params, grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        if p.grad is None:
            p.grad = f.clone()
        else:
            g.copy_(f)


def allreduce_kl_sums(per_tensor_sums, owners, numels, n_batches, group=None):
    """Tensor-sharded KL: every rank holds the element sums of the tensors it owns (0 elsewhere); one
    all-reduce of the small per-tensor vector, then the reference reduction mean-of-means / n_batches
    (loss.py:28,38) applied with exact per-tensor 1/numel."""
    vec = per_tensor_sums.clone()
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    inv = torch.tensor([1.0 / n for n in numels], dtype=vec.dtype, device=vec.device)
    return (vec * inv).sum() / (len(numels) * n_batches)
