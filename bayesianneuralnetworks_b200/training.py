"""ElboTrainer — the body of the reference's training loop (examples/MNIST/train.py:55-65) on this package's API, as one
replayable unit:

    zero_grad -> preds = model(x) (S Monte-Carlo predictions, one batched pass) -> KLDivergence(model)
    -> likelihood = mean_s criterion(pred_s, y) -> backward -> [gradient exchange] -> Adam step

Single GPU: the whole step is captured into ONE CUDA graph (the Philox streams advance through a device-side step
counter, runtime.graph_safe_rng, so every replay draws fresh eps).  Several GPUs (one process per GPU; data x sample
grid) — `exchange`:

  'peer'      the flat gradient buffer lives in peer-mapped memory and the optimizer kernel itself averages the ranks'
              gradients over NVLink (bnn_adam_kl_step_peers between two bnn_peer_barrier launches): still ONE graph, no
              NCCL call.  The choice for the examples' small models (gradients of a few MB: pure latency).
  'bucketed'  per-bucket NCCL all-reduces launched from autograd hooks while the backward pass is still running
              (parallel.BucketedAllReduce); eager launches.  The choice for large gradients (C4: 537 MB).
  'flat'      ONE NCCL all-reduce of the flat buffer between a forward+backward graph and an optimizer graph.
  'auto'      'peer' below 32 MB of gradients when symmetric memory is available, else 'bucketed'.

Partition: rank (d, s) of an R_data x R_sample grid takes batch slice d (the caller feeds it) and the global MC samples
[s * S_local, (s+1) * S_local) of S_local * R_sample (runtime.set_sample_partition); every rank's loss is the mean over
its rows and samples, shards are equal-sized, so the global gradient is the plain average over ALL ranks.  The KL
gradient is data independent and identical on every rank: ELBOAdam adds it after the average.
"""
import torch
import torch.nn.functional as F

from . import _C, optim, parallel, runtime
from .nn import KLDivergence, mc_mean_loss


class ElboTrainer:
    def __init__(self, model, number_of_batches, lr=1e-3, criterion=F.cross_entropy, graph=True, optimizer="elbo-adam",
                 loss_tail="batched", exchange="auto", group=None, sample_groups=1, peer_limit_bytes=32 << 20):
        import torch.distributed as dist
        self.model = model
        self.criterion = criterion
        self.loss_tail = loss_tail
        self.use_graph = bool(graph)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.group = group
        self.device = next(model.parameters()).device
        self.params = [p for p in model.parameters() if p.requires_grad]
        runtime.graph_safe_rng(self.use_graph)
        if self.world > 1:
            if self.world % sample_groups != 0:
                raise ValueError(f"{sample_groups} sample groups do not divide {self.world} ranks")
            self.data_index, self.sample_index = parallel.grid_coordinates(self.rank, self.world, sample_groups)
            runtime.set_sample_partition(self.sample_index, sample_groups)
        else:
            self.data_index = self.sample_index = 0
        self.sample_groups = sample_groups
        self.kld = KLDivergence(number_of_batches=number_of_batches)
        self.optimizer_kind = optimizer
        if optimizer == "elbo-adam":      # KL gradient + Adam in one pass, likelihood-only backward
            self.opt = optim.ELBOAdam(model, number_of_batches=number_of_batches, lr=lr, capturable=self.use_graph)
        else:                             # torch's fused Adam on likelihood + KL (the reference's optimizer, train.py:43)
            self.opt = torch.optim.Adam(model.parameters(), lr=lr, capturable=self.use_graph, fused=True)
        grad_bytes = 4 * sum(p.numel() for p in self.params)
        self.exchange = None
        self.exchange_note = "single GPU"
        self.flat = None
        self.peers = None
        self.buckets = None
        if self.world > 1:
            self.exchange_note = ""
            mode = exchange
            if mode == "auto":
                mode = "peer" if (optimizer == "elbo-adam" and grad_bytes <= peer_limit_bytes) else "bucketed"
            if mode == "peer":
                if optimizer != "elbo-adam":
                    raise ValueError("exchange='peer' is implemented by ELBOAdam's kernel")
                try:
                    self.peers = parallel.PeerGradients(sum(p.numel() for p in self.params), self.device, group)
                except Exception as exc:      # noqa: BLE001 — no symmetric memory on this system: NCCL instead
                    if exchange == "peer":
                        raise
                    self.exchange_note = f"peer memory unavailable ({type(exc).__name__}: {exc}); "
                    mode = "flat"
            if mode == "peer":
                # gradients stay where autograd puts them (set_to_none each step, no accumulation kernels); ONE gather
                # launch copies them into the peer-visible flat buffer after the backward pass
                self._peer_views, off = {}, 0
                for p in self.params:
                    self._peer_views[id(p)] = torch.as_strided(self.peers.flat, p.size(), p.stride(),
                                                               self.peers.flat.storage_offset() + off)
                    off += p.numel()
                self.opt.attach_peers(self.peers, self._peer_views)
                if self.world > self.opt.two_hop_above:
                    self.exchange_note = ("flat gradient buffers averaged in place over NVLink peer memory in two hops (every "
                                          "rank reduces its slice and writes it to all: bnn_peer_average between two "
                                          "bnn_peer_barrier launches), then the local optimizer kernel; no NCCL call; one CUDA "
                                          "graph")
                else:
                    self.exchange_note = ("gradients averaged inside the optimizer kernel over NVLink peer memory "
                                          "(bnn_adam_kl_step_peers + bnn_peer_barrier); no NCCL call; one CUDA graph")
            elif mode == "bucketed":
                self.buckets = parallel.BucketedAllReduce(self.params, group=group)
                self.flat = self.buckets.flat
                self.use_graph = False
                runtime.graph_safe_rng(False)
                self.exchange_note += (f"{len(self.buckets.buckets)} NCCL all-reduce buckets launched from autograd hooks "
                                       "on a side stream, overlapping the backward pass; eager launches")
            elif mode == "flat":
                self.flat = torch.zeros(sum(p.numel() for p in self.params), device=self.device, dtype=torch.float32)
                self._bind_flat()
                self.exchange_note += ("one NCCL all-reduce of the flat gradient buffer between a forward+backward graph "
                                       "and an optimizer graph")
            else:
                raise ValueError(f"unknown exchange mode {exchange!r}")
            self.exchange = mode
        self.graph = None
        self.graph_opt = None
        self.static_loss = None
        self.launches_per_step = None
        self._copy_stream = None
        self._prefetched = None

    def _bind_flat(self):
        """Gradients become views of ONE flat buffer (same strides as their parameter: channels_last trunk weights are
        dense permutations of their storage), so that the exchange sees static addresses and a single range."""
        off = 0
        for p in self.params:
            p.grad = torch.as_strided(self.flat, p.size(), p.stride(), self.flat.storage_offset() + off)
            off += p.numel()

    # ------------------------------------------------------------------------------------------------ one step
    def _forward_backward(self, x, y):
        if self.use_graph:
            runtime.advance_rng_step(x.device)
        if self.flat is not None:
            self.flat.zero_()                 # gradients are views of one flat buffer (static addresses)
        else:
            self.opt.zero_grad(set_to_none=True)
        preds = self.model(x)
        elbo = self.optimizer_kind == "elbo-adam"
        composite = None
        if elbo:
            with torch.no_grad():
                divergence = self.kld.fused_part(self.model)
            if self.opt.composite:            # KL terms without a fused form (full-covariance head): through autograd
                composite = self.kld.composite_part(self.model)
        else:
            divergence = self.kld(self.model)
        if self.loss_tail == "loop":        # the reference loop body verbatim (train.py:59-61)
            likelihood = torch.stack([self.criterion(p, y) for p in preds]).mean()
        else:                               # the same mean as ONE cross-entropy over the S*B rows
            likelihood = mc_mean_loss(self.criterion, preds, y)
        if elbo:                            # the optimizer adds the closed-form KL gradient; the value is still reported
            (likelihood if composite is None else likelihood + composite).backward()
            value = likelihood.detach() if divergence is None else likelihood.detach() + divergence
            return value if composite is None else value + composite.detach()
        loss = likelihood + divergence
        loss.backward()
        return loss

    def _pack(self):
        """'peer' exchange: every gradient -> its slot of the peer-visible flat buffer, one launch."""
        items, off = [], 0
        for p in self.params:
            g = p.grad
            if g is not None and g.stride() != p.stride():      # not in the parameter's memory order: through the view
                self._peer_views[id(p)].copy_(g)
            else:
                items.append((g, off, p.numel()))
            off += p.numel()
        _C.pack_gradients(items, self.peers.flat)

    def _exchange(self):
        import torch.distributed as dist
        if self.exchange == "peer":
            self._pack()
        elif self.exchange == "flat":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        elif self.exchange == "bucketed":
            self.buckets.finish()
        # 'peer': the averaging itself happens inside opt.step()

    def _body(self, x, y):
        loss = self._forward_backward(x, y)
        if self.world > 1:
            self._exchange()
        self.opt.step()
        return loss

    def capture(self, x, y):
        """Warm up eagerly on a side stream, then capture the step on static input buffers.  The warm-up steps are real
        steps (every rank runs the same number, the peer barriers pair up)."""
        if not self.use_graph:
            return
        # warm-up runs on a side stream, capture on the capture stream: the gradient accumulators legitimately see two
        # streams (neither is the legacy default stream), so silence torch's advisory about it
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self.sx, self.sy = x.clone(), y.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(self.sx, self.sy)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = _C.launch_count
        self.graph = torch.cuda.CUDAGraph()
        two_graphs = self.exchange == "flat"
        with torch.cuda.graph(self.graph):
            self.static_loss = self._forward_backward(self.sx, self.sy)
            if not two_graphs:
                if self.world > 1:
                    self._exchange()
                self.opt.step()
        if two_graphs:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self.opt.step()
        self.launches_per_step = _C.launch_count - before

    def prefetch(self, x, y):
        """Start the host -> device copy of the NEXT step's (pinned) batch on a copy stream, so that it overlaps the step
        that is running — what a DataLoader with pin_memory + non_blocking copies does for the reference loop
        (train.py:52-56).  `step(x, y)` with the same tensors then only waits for that copy.  Two staging buffers
        alternate; a buffer is rewritten only after the step that consumed it has picked it up."""
        if x.is_cuda:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staging = [None, None]
            self._consumed = [None, None]
            self._turn = 0
        k = self._turn
        self._turn ^= 1
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[k] is not None:
                self._copy_stream.wait_event(self._consumed[k])
            if self._staging[k] is None or self._staging[k][0].shape != x.shape or self._staging[k][1].shape != y.shape:
                self._staging[k] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                                    torch.empty(y.shape, dtype=y.dtype, device=self.device))
            sx, sy = self._staging[k]
            sx.copy_(x, non_blocking=True)
            sy.copy_(y, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        self._prefetched = (x, y, k, ready)

    def _take_prefetched(self, x, y):
        """The staged device copy of (x, y) if prefetch() was called with exactly these tensors, else None."""
        pf = self._prefetched
        if pf is None or pf[0] is not x or pf[1] is not y:
            return None
        self._prefetched = None
        _, _, k, ready = pf
        torch.cuda.current_stream().wait_event(ready)
        return k, self._staging[k]

    def _release_staging(self, k):
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        self._consumed[k] = done

    def step(self, x, y):
        """One training step; returns the (device) loss tensor.  With a captured graph, x / y are copied into the
        graph's static inputs (they may be pinned host tensors; `prefetch` overlaps that copy with the previous step)."""
        staged = self._take_prefetched(x, y)
        if self.graph is None:
            if staged is not None:
                k, (dx, dy) = staged
                loss = self._body(dx, dy)
                self._release_staging(k)
                return loss
            if not x.is_cuda:
                x, y = x.to(self.device, non_blocking=True), y.to(self.device, non_blocking=True)
            return self._body(x, y)
        if staged is not None:
            k, (dx, dy) = staged
            self.sx.copy_(dx, non_blocking=True)      # device -> device: microseconds
            self.sy.copy_(dy, non_blocking=True)
            self._release_staging(k)
        else:
            self.sx.copy_(x, non_blocking=True)
            self.sy.copy_(y, non_blocking=True)
        self.graph.replay()
        if self.graph_opt is not None:
            self._exchange()
            self.graph_opt.replay()
        return self.static_loss

    def release(self):
        """Drop graphs, hooks and partition state (bench.py runs several workloads in one process)."""
        self.graph = self.graph_opt = None
        if self.buckets is not None:
            self.buckets.close()
        runtime.set_sample_partition(0, 1)
        runtime.graph_safe_rng(False)
