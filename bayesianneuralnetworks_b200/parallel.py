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


def default_grid(world):
    """(data groups, sample groups) of the data x sample grid used when none is given: 2 = 1 x 2, 4 = 2 x 2, 8 = 2 x 4
    (sample groups need no BatchNorm caveat; more than four of them leave few samples per rank)."""
    sample_groups = {1: 1, 2: 2, 4: 2, 8: 4}.get(world, 1)
    return world // sample_groups, sample_groups


# ------------------------------------------------------------------------------------------------ peer gradients
class PeerGradients:
    """This rank's flat gradient buffer, allocated so that every rank of the node maps every other rank's buffer (torch
    symmetric memory: CUDA virtual-memory handles exchanged through the process group's store; NVLink peer access), plus
    the flag blocks of the barrier kernel.  Consumed by optim.ELBOAdam.attach_peers: the optimizer kernel reads all
    ranks' copies of a gradient element and averages them (bnn_adam_kl_step_peers), bnn_peer_barrier orders the steps.

    `flat` is the local buffer (float32 [numel]); make the parameters' `.grad` views of it (training.ElboTrainer does)."""

    def __init__(self, numel, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _C
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _C.MAX_PEERS:
            raise ValueError(f"peer gradient exchange supports up to {_C.MAX_PEERS} ranks of one node, got {self.world}")
        self.device = torch.device(device)
        padded = (numel + 3) // 4 * 4
        self._flag_words = 64                       # one 256-byte line per rank's flag block
        # one allocation: [flags | gradients]; symmetric memory hands back the peers' base pointers
        self._buf = symm.empty(self._flag_words + padded, dtype=torch.float32, device=self.device)
        self._buf.zero_()
        torch.cuda.synchronize(self.device)
        self._handle = symm.rendezvous(self._buf, group)
        ptrs = [int(p) for p in self._handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self._buf.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned unexpected peer pointers")
        self.flag_ptrs = ptrs
        self.bases = [p + 4 * self._flag_words for p in ptrs]
        self.flat = self._buf[self._flag_words:self._flag_words + numel]
        self._epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=group)                   # every rank's flags are zero before anyone signals
        torch.cuda.synchronize(self.device)

    def barrier(self):
        """bnn_peer_barrier on the current stream (capturable)."""
        from . import _C
        _C.peer_barrier(self.flag_ptrs, self.rank, self._epoch, self.device)


# ------------------------------------------------------------------------------------------------ bucketed exchange
class BucketedAllReduce:
    """Gradient all-reduce in buckets launched WHILE the backward pass is still running ("bucketed
    per layer and overlapped with backward"): every parameter's gradient is a view of one flat buffer laid out in REVERSE
    registration order (roughly the order in which autograd finishes them); a post-accumulate hook counts a bucket's
    parameters down and, when the bucket is complete, launches its NCCL all-reduce (AVG) on a side stream that waits for
    the producing stream.  `finish()` makes the consumer stream wait for the exchange.  For large gradients (C4: 537 MB,
    one bucket per layer), where the exchange would otherwise sit exposed between backward and the optimizer."""

    def __init__(self, params, bucket_bytes=64 << 20, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        device = order[0].device
        self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        self.buckets = []                      # [begin, end, pending, n_params]
        self._bucket_of = {}
        off, begin, count = 0, 0, 0
        for p in order:
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), self.flat.storage_offset() + off)
            self._bucket_of[id(p)] = len(self.buckets)
            off += p.numel()
            count += 1
            if (off - begin) * 4 >= bucket_bytes:
                self.buckets.append([begin, off, count, count])
                begin, count = off, 0
        if count:
            self.buckets.append([begin, off, count, count])
        self.stream = torch.cuda.Stream(device=device)
        self._hooks = [p.register_post_accumulate_grad_hook(self._ready) for p in self.params]
        self._launched = 0

    def _ready(self, p):
        b = self.buckets[self._bucket_of[id(p)]]
        b[2] -= 1
        if b[2] == 0:
            b[2] = b[3]
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(self.flat[b[0]:b[1]], op=dist.ReduceOp.AVG, group=self.group)
            self._launched += 1

    def zero(self):
        self.flat.zero_()

    def finish(self):
        """Call after backward(): the current stream waits for every bucket's all-reduce."""
        torch.cuda.current_stream().wait_stream(self.stream)
        self._launched = 0

    def close(self):
        for h in self._hooks:
            h.remove()
