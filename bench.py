#!/usr/bin/env python
"""bench.py — ELBO training throughput of the variational hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4]

Metric (BASELINE.json): ELBO train samples*MC/sec = B*S*N / step time, where one step is the body
of the reference's training loop, examples/MNIST/train.py:55-65: zero_grad -> model(x) (S Monte-Carlo
predictions) -> KLDivergence(model) -> mean cross-entropy over the S predictions -> backward ->
[gradient all-reduce] -> Adam step.  Default workload = BASELINE.json configs[1] ("c2"): the
examples/MNIST/model.py topology (what the FashionMNIST topology is with NormalConv2d/NormalLinear,
SURVEY §0-4), 28x28 inputs, batch 256 per GPU, S=8.  Synthetic data, reference initialisation.

One JSON line on stdout (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step fed from
pinned host memory with the loss read back every step; `roofline`: the dominant hot-path kernel timed
live with CUDA events; `cpu_baseline`: the oracle's restatement of the reference step on the host
cores (bounded sample); `kl_prune`: the bandwidth-bound KL / prune sweeps (C5 shape, bounded size).
`--impl reference` times the reference's CPU path (oracle port) on the same config.

Torch-side settings of the deterministic trunk (Conv2d / BatchNorm2d / ELU / Linear of the example models, outside the
hot path but inside the measured step): cuDNN autotuning (`--no-cudnn-benchmark`), torch.channels_last for the trunk
modules (`--nchw-trunk`), allow_tf32 in TF32 mode.  The likelihood term goes through nn.mc_mean_loss
(`--loss-tail loop` = the reference loop verbatim), the optimizer is bnn.optim.ELBOAdam (`--optimizer adam`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ["NCCL_DEBUG"] = "WARN"        # NCCL's version banner goes to stdout; this program prints ONE JSON line there

N_BATCHES = 469          # ceil(60000 / 128), examples/MNIST/train.py:38 (any constant)
WORKLOADS = {
    "c2": dict(name="C2 examples/MNIST BCNN topology (NormalConv2d 64x64x3x3 s2 + NormalLinear 576x10), 28x28",
               batch=256, samples=8),
    "c3": dict(name="C3 examples/CIFAR10 BCNN topology (NormalConv2d 128x128x3x3 on 4x4 maps; the full-covariance "
                    "MultivariateNormalLinear head, out of the hot path, replaced by NormalLinear(128,10)), 32x32x3",
               batch=512, samples=16),
    "c4": dict(name="C4 wide Bayesian MLP 4x NormalLinear(4096,4096)", batch=1024, samples=32),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ models
def build_model(workload, samples):
    import bayesianneuralnetworks_b200 as bnn
    from torch.nn import BatchNorm2d, Conv2d, ELU, Flatten, Sequential, Softmax

    class Net(bnn.nn.BayesianNetworkModule):
        def __init__(self, layers, cin, cout):
            super().__init__(cin, cout, samples)
            self.layers = layers

        def _forward(self, x):
            return self.layers(x)

    if workload == "c2":        # examples/MNIST/model.py:20-33
        layers = Sequential(Conv2d(1, 32, 5, padding=2, stride=2), BatchNorm2d(32), ELU(),
                            Conv2d(32, 32, 3, padding=1, stride=1), ELU(),
                            Conv2d(32, 64, 3, padding=0, stride=2), ELU(),
                            bnn.nn.NormalConv2d(64, 64, 3, padding=1, stride=2), ELU(), Flatten(),
                            bnn.nn.NormalLinear(576, 10), Softmax(dim=-1))
        return Net(layers, 1, 10)
    if workload == "c3":        # examples/CIFAR10/model.py:20-39
        from torch.nn import Linear
        layers = Sequential(Conv2d(3, 64, 5, padding=2, stride=2), BatchNorm2d(64), ELU(),
                            Conv2d(64, 128, 5, padding=2, stride=2), ELU(),
                            Conv2d(128, 128, 5, padding=2, stride=2), ELU(),
                            Conv2d(128, 128, 3, padding=1), ELU(), Conv2d(128, 128, 3, padding=1), ELU(),
                            bnn.nn.NormalConv2d(128, 128, 3, padding=1), ELU(), Flatten(),
                            Linear(2048, 128), ELU(), bnn.nn.NormalLinear(128, 10), Softmax(dim=-1))
        return Net(layers, 3, 10)
    layers = Sequential(bnn.nn.NormalLinear(4096, 4096), ELU(), bnn.nn.NormalLinear(4096, 4096), ELU(),
                        bnn.nn.NormalLinear(4096, 4096), ELU(), bnn.nn.NormalLinear(4096, 4096), Softmax(dim=-1))
    return Net(layers, 4096, 4096)


def synthetic_batch(workload, batch, gen):
    if workload == "c2":
        return torch.rand(batch, 1, 28, 28, generator=gen), torch.randint(0, 10, (batch,), generator=gen)
    if workload == "c3":
        return torch.rand(batch, 3, 32, 32, generator=gen), torch.randint(0, 10, (batch,), generator=gen)
    return torch.randn(batch, 4096, generator=gen), torch.randint(0, 4096, (batch,), generator=gen)


def hot_flops_per_step(workload, batch, samples):
    """Algorithmic flops of the hot-path contractions per step: fwd 2MNK, bwd 4MNK (2MNK for a
    first layer without dX)."""
    if workload == "c2":
        conv = 2 * 9 * 64 * 576         # per sample*MC row, forward
        lin = 2 * 10 * 576
        return batch * samples * 3 * (conv + lin)
    if workload == "c3":
        return batch * samples * 3 * (2 * 16 * 128 * 1152 + 2 * 10 * 128)
    return batch * samples * (4 * 6 - 2) * 4096 * 4096


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe).  The sampler
    runs from before the warm-up (nvidia-smi needs ~100 ms to deliver its first line) and every line is stamped on
    arrival; `stop()` reports the samples that fall inside the marked windows — or, for a window shorter than the
    sampling period, the samples nearest to it."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [(t, r) for t, r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        inside = [r for t, r in rows if any(a - 0.05 <= t <= b + 0.05 for a, b in self.windows)]
        note = "inside the timed regions"
        if not inside and rows and self.windows:
            mid = sum(a + b for a, b in self.windows) / (2 * len(self.windows))
            inside = [r for _, r in sorted(rows, key=lambda tr: abs(tr[0] - mid))[:3]]
            note = "nearest to the timed regions (shorter than the sampling period)"
        sm = sorted(float(r[1]) for r in inside)
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None, "window": note}


# ------------------------------------------------------------------------------------------------ the step
class Trainer:
    """The reference training-loop body (train.py:55-65) on this repo's public API.  With `graph=True` the whole
    step (RNG advance, zero_grad, S-sample forward, KL, CE, backward, gradient all-reduce, Adam) is captured
    once into a CUDA graph and replayed: the Philox streams advance through a device-side step counter
    (bnn.graph_safe_rng), so every replay draws fresh eps."""

    def __init__(self, workload, device, world, samples, graph, loss_tail="batched", optimizer="elbo-adam",
                 channels_last=False):
        import bayesianneuralnetworks_b200 as bnn
        self.bnn = bnn
        self.loss_tail = loss_tail
        torch.manual_seed(0)
        bnn.graph_safe_rng(graph)
        self.model = build_model(workload, samples).to(device)
        if channels_last:                 # torch-side knob for the deterministic trunk: cuDNN's NHWC kernels without
            for m in self.model.modules():     # layout conversions around each call; (mu, rho) stay row-major OIHW
                if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
                    m.to(memory_format=torch.channels_last)
        self.kld = bnn.nn.KLDivergence(number_of_batches=N_BATCHES)
        self.optimizer = optimizer
        if optimizer == "elbo-adam":      # KL gradient + Adam in one pass, likelihood-only backward
            self.opt = bnn.optim.ELBOAdam(self.model, number_of_batches=N_BATCHES, lr=1e-3, capturable=graph)
        else:                             # torch's single-pass fused Adam (the reference's torch.optim.Adam, train.py:43)
            self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-3, capturable=graph, fused=True)
        self.world = world
        self.params = [p for p in self.model.parameters()]
        self.graph = None
        self.graph_opt = None
        self.flat = None
        self.use_graph = graph
        self.launches_per_step = None

    def _forward_backward(self, x, y):
        if self.use_graph:
            self.bnn.advance_rng_step(x.device)
        if self.flat is not None:
            self.flat.zero_()                 # gradients are views of one flat buffer (static addresses)
        else:
            self.opt.zero_grad(set_to_none=True)
        preds = self.model(x)
        if self.optimizer == "elbo-adam":
            with torch.no_grad():
                divergence = self.kld(self.model)
        else:
            divergence = self.kld(self.model)
        if self.loss_tail == "loop":        # the reference loop body verbatim (train.py:59-61)
            likelihood = torch.stack([F.cross_entropy(p, y) for p in preds]).mean()
        else:                               # the same mean as ONE cross-entropy over the S*B rows
            likelihood = self.bnn.nn.mc_mean_loss(F.cross_entropy, preds, y)
        if self.optimizer == "elbo-adam":   # the optimizer adds the closed-form KL gradient; the value is still reported
            likelihood.backward()
            return likelihood.detach() + divergence
        loss = likelihood + divergence
        loss.backward()
        return loss

    def _allreduce(self):
        """Data parallel: ONE all-reduce (avg) of the flat gradient buffer."""
        import torch.distributed as dist
        if self.flat is not None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            from bayesianneuralnetworks_b200 import parallel
            parallel.allreduce_gradients(self.params)

    def _body(self, x, y):
        loss = self._forward_backward(x, y)
        if self.world > 1:
            self._allreduce()
        self.opt.step()
        return loss

    def capture(self, x, y):
        """Warm up eagerly on a side stream, then capture the step on static input buffers.  One GPU: one graph.
        Several GPUs: forward+backward and the optimizer step are two graphs with the NCCL all-reduce of the flat
        gradient buffer launched eagerly between them (collectives stay outside the captured region)."""
        from bayesianneuralnetworks_b200 import _C
        # warm-up runs on a side stream, capture on the capture stream: the gradient accumulators legitimately see two
        # streams (neither is the legacy default stream), so silence torch's advisory about it
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self.sx, self.sy = x.clone(), y.clone()
        if self.world > 1:
            # the gradients are views of ONE flat buffer: a single all-reduce, static addresses for both graphs
            total = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(total, device=x.device, dtype=torch.float32)
            off = 0
            for p in self.params:
                # same strides as the parameter (channels_last trunk weights are dense permutations of their storage):
                # torch's fused Adam wants gradients in the parameter's layout
                p.grad = torch.as_strided(self.flat, p.size(), p.stride(), off)
                off += p.numel()
        # one GPU: zero_grad(set_to_none=True) inside the captured step — autograd then hands the freshly computed
        # gradient tensors (static addresses in the graph's pool) to .grad without an accumulation pass
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._body(self.sx, self.sy)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        before = _C.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._forward_backward(self.sx, self.sy)
            if self.world == 1:
                self.opt.step()
        if self.world > 1:
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt):
                self.opt.step()
        self.launches_per_step = _C.launch_count - before

    def step(self, x, y):
        if self.graph is None:
            return self._body(x, y)
        self.sx.copy_(x, non_blocking=True)
        self.sy.copy_(y, non_blocking=True)
        self.graph.replay()
        if self.world > 1:
            self._allreduce()
            self.graph_opt.replay()
        return self.static_loss


def run_b200(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from bayesianneuralnetworks_b200 import _C
    import bayesianneuralnetworks_b200 as bnn
    _C.lib()                          # fails loudly when the CUDA library is missing
    wl = WORKLOADS[args.workload]
    B, S = wl["batch"], wl["samples"]
    bnn.set_precision(args.precision)
    sample_parallel = args.parallel == "sample" and world > 1
    if sample_parallel:          # rank r evaluates the global MC samples [r*S/R, (r+1)*S/R) of the SAME batch
        if S % world != 0:
            raise SystemExit(f"{S} MC samples do not split over {world} ranks")
        bnn.set_sample_partition(rank, world)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    if args.precision == "tf32":      # TF32 hot path, "deterministic trunk through torch with allow_tf32"
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    trainer = Trainer(args.workload, device, world, S, graph=not args.no_graph, loss_tail=args.loss_tail,
                      optimizer=args.optimizer, channels_last=not args.nchw_trunk)
    gen = torch.Generator().manual_seed(1 if sample_parallel else 1 + rank)
    n_host = 8
    host = [tuple(t.pin_memory() for t in synthetic_batch(args.workload, B, gen)) for _ in range(n_host)]
    dev = [(x.to(device), y.to(device)) for x, y in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, feed_from_host):
        """K steps, each bracketed by CUDA events on the current stream; L2 flushed (untimed) between steps."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        barrier()
        t0 = time.perf_counter()
        last = None
        for i in range(n_steps):
            flush.zero_()
            starts[i].record()
            if feed_from_host and trainer.graph is not None:
                x, y = host[i % n_host]       # pinned host -> the graph's static input buffers, one async copy each
            elif feed_from_host:
                hx, hy = host[i % n_host]
                x, y = hx.to(device, non_blocking=True), hy.to(device, non_blocking=True)
            else:
                x, y = dev[i % n_host]
            loss = trainer.step(x, y)
            if feed_from_host:
                last = loss.item()            # device -> host read of the step's result, every step
            stops[i].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops))
        return ms / 1e3, wall, last

    clocks = ClockSampler(local)
    clocks.start()
    graph_note = "eager launches"
    if trainer.use_graph:
        try:
            trainer.capture(*dev[0])
            graph_note = ("whole step captured in one CUDA graph (device-side Philox step counter), replayed per step"
                          if world == 1 else "forward+backward and optimizer captured as two CUDA graphs, the NCCL "
                          "all-reduce of the flat gradient buffer launched eagerly between them")
        except Exception as exc:      # noqa: BLE001 — fall back to eager launches, say so in the result
            sys.stderr.write(f"CUDA graph capture failed ({exc!r}); running eagerly\n")
            trainer.graph, trainer.graph_opt, trainer.use_graph = None, None, False
            bnn.graph_safe_rng(False)
            graph_note = f"eager launches (graph capture failed: {type(exc).__name__})"
    for i in range(max(args.warmup, 3)):
        trainer.step(*dev[i % n_host])
    launches0 = _C.launch_count
    t_a = time.time()
    dev_s, dev_wall, _ = timed(args.steps, False)
    clocks.mark(t_a, time.time())
    launches = _C.launch_count - launches0
    if trainer.graph is not None:
        launches = trainer.launches_per_step * args.steps      # replays launch the captured kernels
    timed(2, True)                                    # untimed: first-touch of the staging allocations of the host-fed path
    t_a = time.time()
    e2e_s, e2e_wall, last_loss = timed(args.steps, True)
    clocks.mark(t_a, time.time())
    clock_info = clocks.stop()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    dev_s, e2e_s = max_over_ranks(dev_s), max_over_ranks(e2e_s)
    units = B * S * (1 if sample_parallel else world) * args.steps
    pk = peaks()

    # ---- roofline of the dominant hot-path kernel, timed live with CUDA events on the launching stream
    _C.set_kernel_timing(True)
    n_prof = min(args.steps, 10)
    for i in range(n_prof):
        flush.zero_()
        trainer._body(*dev[i % n_host])          # eager launches so that every library call can be bracketed
    per_kernel = _C.kernel_timing_summary(n_prof)
    _C.set_kernel_timing(False)
    roof = roofline(args.workload, B, S, per_kernel, pk)

    out = None
    if rank == 0:
        x0, y0 = host[0]
        out = {
            "metric": "ELBO train samples*MC/sec", "value": units / dev_s, "unit": "samples*MC/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "strong" if sample_parallel else "weak",
            "vs_baseline": None,
            "dtype": "tf32" if args.precision == "tf32" else "fp32 (3xTF32 split on tcgen05)",
            "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "mc_samples": S,
                       "global_batch": B if sample_parallel else B * world,
                       "parallelism": (f"sp{world} (MC samples sharded, {S // world} per GPU)" if sample_parallel
                                       else f"dp{world}") if world > 1 else "single", "n_batches": N_BATCHES,
                       "optimizer": "Adam (torch fused)" if args.optimizer == "adam" else "bnn.optim.ELBOAdam (KL gradient + Adam in one pass; torch fused Adam for the deterministic layers)", "cudnn_benchmark": not args.no_cudnn_benchmark, "trunk_allow_tf32": args.precision == "tf32", "trunk_memory_format": "contiguous (NCHW)" if args.nchw_trunk else "channels_last (torch Conv2d / BatchNorm2d modules only)", "launch": graph_note, "l2": "flushed between steps (256 MiB write, untimed); each step "
                       "timed with its own CUDA event pair", "step": "zero_grad+forward(S)+KL+CE+backward+Adam" + (" (KL gradient applied inside the optimizer pass)" if args.optimizer == "elbo-adam" else ""),
                       "loss_tail": ("nn.mc_mean_loss: mean of the S per-sample cross-entropies evaluated as one call over "
                                     "the S*B rows (identical value and gradients, tests/test_modules_gpu.py)"
                                     if args.loss_tail == "batched" else "reference loop: S cross-entropy calls")},
            "e2e": {"value": units / e2e_s, "unit": "samples*MC/s",
                    "h2d_bytes_per_step": x0.numel() * x0.element_size() + y0.numel() * y0.element_size(),
                    "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * e2e_s / args.steps, "last_loss": last_loss},
            "gpu_launches": launches,
            "wall_s": {"device_resident": dev_wall, "e2e": e2e_wall},
            "clocks": clock_info,
            "roofline": roof,
            "hot_path": {"algorithmic_tflops_per_s": hot_flops_per_step(args.workload, B, S) * (1 if sample_parallel else world) /
                         (dev_s / args.steps) / 1e12, "kernels_ms_per_step": per_kernel},
            "peaks": pk,
        }
    if not args.no_extras:
        kl_prune = bench_kl_prune(device, pk, world=world)          # every rank sweeps its shard of the tensors
        if rank == 0:
            out["kl_prune"] = kl_prune
    if rank == 0 and world == 1 and not args.no_extras:
        out["cpu_baseline"] = cpu_baseline(args.workload, bounded_seconds=20.0)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def measured_traffic(workload, name):
    """DRAM bytes per launch of `name` from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get(workload, {}).get(name, {}).get("bytes_per_launch")
    except (OSError, ValueError):
        return None


# per workload: the Bayesian layers as (rows per (batch row, MC sample), N, K, input elements per (row, sample))
# — the conv layers' K is Cin*kh*kw of the implicit GEMM, their algorithmic input is the un-expanded NCHW tensor
HOT_LAYERS = {
    "c2": [(9, 64, 576, 64 * 6 * 6), (1, 10, 576, 576)],
    "c3": [(16, 128, 1152, 128 * 4 * 4), (1, 10, 128, 128)],
    "c4": [(1, 4096, 4096, 4096)] * 4,
}


def roofline(workload, B, S, per_kernel, pk):
    """Dominant hot-path kernel = the libbnn_b200 CONTRACTION entry point with the largest time share of the step.
    Its roof follows from its arithmetic intensity: algorithmic flops / algorithmic bytes (operands read once, results
    written once, fp32) against the ridge point peak_tensor / peak_hbm — the narrow layers of C2/C3 (N = 64 / 128 / 10)
    sit on the bandwidth side, the 4096-wide layers of C4 on the tensor side.  Both fractions are reported."""
    if not per_kernel:
        return None
    gemms = {k: v for k, v in per_kernel.items() if k.startswith("bnn_sampled_gemm")}
    name = max(gemms or per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    k = per_kernel[name]
    base = {"kernel": name, "launches_per_step": k["launches_per_step"],
            "avg_launch_us": 1e3 * k["ms_per_step"] / k["launches_per_step"]}
    if name not in gemms:
        return dict(base, bound="hbm", achieved=None, peak=pk["hbm"], unit="GB/s", frac=None, traffic=None)
    layers = HOT_LAYERS[workload]
    if name == "bnn_sampled_gemm_dgrad" and workload == "c4":
        layers = layers[1:]                      # the first layer needs no input gradient
    rows = B * S
    flops = sum(2 * r * n * kk for r, n, kk, _ in layers) * rows
    if name == "bnn_sampled_gemm_fwd":           # x, (mu, sigma) -> y (+ bias)
        nbytes = sum(rows * (x_in + r * n) + 2 * n * kk for r, n, kk, x_in in layers) * 4
    elif name == "bnn_sampled_gemm_dgrad":       # dy, (mu, sigma) -> dx
        nbytes = sum(rows * (r * n + x_in) + 2 * n * kk for r, n, kk, x_in in layers) * 4
    else:                                        # dy, x, rho -> dmu, drho
        nbytes = sum(rows * (r * n + x_in) + 3 * n * kk for r, n, kk, x_in in layers) * 4
    seconds = k["ms_per_step"] * 1e-3
    tensor_peak = pk["bf16_sustained"] / 2.0      # TF32 dense = half the bf16 rate on this tensor pipe
    tflops, gbs = flops / seconds / 1e12, nbytes / seconds / 1e9
    intensity, ridge = flops / nbytes, tensor_peak * 1e12 / (pk["hbm"] * 1e9)
    out = dict(base, arithmetic_intensity=intensity, ridge_point=ridge,
               tensor={"achieved": tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tflops / tensor_peak},
               hbm={"achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]},
               traffic=measured_traffic(workload, name), algorithmic_bytes_per_step=nbytes,
               algorithmic_flops_per_step=flops,
               peak_note="TF32 dense peak taken as half of the measured sustained bf16 cuBLAS rate "
                         f"({pk['source']}); fp32 mode issues 3 TF32 MMAs per product; timed eagerly with CUDA events "
                         "around the C-ABI call (includes its launch latency)")
    side = "tensor" if intensity >= ridge else "hbm"
    out.update(bound=side, achieved=out[side]["achieved"], peak=out[side]["peak"], unit=out[side]["unit"],
               frac=out[side]["frac"])
    return out


def bench_kl_prune(device, pk, world=1, tensors=64):
    """C5 (BASELINE.json configs[4]): KL forward, KL forward+grad and the pruning sweep over 2^30
    (mu, rho) pairs — 64 tensors of 4096x4096, 8 GiB — sharded round-robin by tensor over the GPUs (64 / N tensors per
    GPU; one GPU sweeps all of them); the working set per GPU is 8 GiB / N >> L2.  Algorithmic bytes: KL
    fwd 8 B/pair, fwd+grad 16 B/pair, prune 8 + 8p B/pair.  With several GPUs the KL legs include the one scalar
    all-reduce of the sharded sum, prune needs no exchange; times are the max over ranks and GB/s the
    aggregate."""
    from bayesianneuralnetworks_b200 import _C
    n_t = max(1, tensors // world)
    pairs = n_t * 4096 * 4096
    gen = torch.Generator(device=device).manual_seed(5)
    mus = [(torch.rand(4096, 4096, device=device, generator=gen) * 2 - 1) / 64 for _ in range(n_t)]
    rhos = [torch.randn(4096, 4096, device=device, generator=gen) * 0.15 - 2.0 for _ in range(n_t)]
    gm = [torch.empty_like(m) for m in mus]
    gr = [torch.empty_like(m) for m in mus]
    fwd = [(m, r, None, None, 0.0, 0.1, 1.0) for m, r in zip(mus, rhos)]
    both = [(m, r, a, b, 0.0, 0.1, 1.0) for m, r, a, b in zip(mus, rhos, gm, gr)]

    def over_ranks(seconds):
        if world == 1:
            return seconds
        import torch.distributed as dist
        t = torch.tensor([seconds], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def kl_leg(entries):
        sums = _C.kl(entries)
        if world > 1:                       # tensor-sharded KL: one small all-reduce
            import torch.distributed as dist
            total = sums.sum()
            dist.all_reduce(total)

    def time_it(fn, reps=5, inner=8):
        """Average duration of one call: `inner` back-to-back calls between one event pair (the host-side cost of
        preparing a launch overlaps the previous kernel, as it does in a training loop; every call streams the whole
        2 GiB working set, so no call finds its data in L2), best of `reps`."""
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(inner):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / inner)
        return over_ranks(best * 1e-3)

    def entry(bytes_per_pair, t):
        gbps = bytes_per_pair * pairs * world / t / 1e9
        return {"GBps": gbps, "frac_of_measured_hbm": gbps / (pk["hbm"] * world), "ms": t * 1e3}

    res = {"workload": "C5: 64 tensors of 4096x4096 (2^30 pairs, 8 GiB of mu/rho), sharded by tensor over the GPUs",
           "pairs_per_gpu": pairs, "tensors_per_gpu": n_t, "n_gpus": world,
           "l2": f"working set {pairs * 8 / 2 ** 30:.0f} GiB per GPU, larger than L2",
           "timing": "CUDA events; KL legs: average of 8 back-to-back calls (best of 5); prune: one call per measurement "
                     "(it modifies its input, restored untimed), best of 3"}
    res["kl_fwd"] = entry(8, time_it(lambda: kl_leg(fwd)))
    res["kl_fwd_grad"] = entry(16, time_it(lambda: kl_leg(both)))
    p = 0.75
    k = int(p * 4096 * 4096)
    del gm, gr
    saved = [(m.clone(), r.clone()) for m, r in zip(mus, rhos)]

    def prune_once():
        _C.prune([(m, r, k, None, None) for m, r in zip(mus, rhos)])
    prune_once()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        for (m, r), (sm, sr) in zip(zip(mus, rhos), saved):
            m.copy_(sm), r.copy_(sr)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prune_once()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    res["prune_p0.75"] = entry(8 + 8 * p, over_ranks(best))
    # the example's sweep (examples/MNIST/prune.py:49-50): successive levels on the SAME tensors — what was pruned at
    # one level (mu = 0, rho = -30: the largest key there is) is re-selected first at the next
    for (m, r), (sm, sr) in zip(zip(mus, rhos), saved):
        m.copy_(sm), r.copy_(sr)
    torch.cuda.synchronize()
    sweep = []
    for level in torch.linspace(.75, 1, 6).tolist():
        kk = int(level * 4096 * 4096)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _C.prune([(m, r, kk, None, None) for m, r in zip(mus, rhos)])
        b.record()
        torch.cuda.synchronize()
        t = over_ranks(a.elapsed_time(b) * 1e-3)
        sweep.append(dict(entry(8 if kk == 4096 * 4096 else 8 + 8 * level, t), p=round(level, 2)))      # p = 1 writes only
    pruned = sum(int((r == -30).sum()) for r in rhos)
    res["prune_sweep"] = {"levels": sweep, "all_pruned_after_p1": pruned == n_t * 4096 * 4096}
    return res


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_step_fn(workload, B, S):
    """The reference step restated on CPU tensors by oracle/variational_oracle.py (ElboStepOracle)."""
    from oracle import variational_oracle as orc
    torch.manual_seed(0)
    model = build_model(workload, S)      # used only as a container of identically initialised parameters
    stages, cur = [], []
    for m in model.layers:
        kind = type(m).__name__
        if kind in ("NormalConv2d", "NormalLinear"):
            if cur:
                stages.append(('torch', torch.nn.Sequential(*cur)))
                cur = []
            ps = [m.weight.mean, m.weight.scale, m.bias.mean, m.bias.scale]
            loc, scale = float(m.weight_prior.loc), float(m.weight_prior.scale)
            if kind == "NormalLinear":
                stages.append(('linear', *ps, loc, scale))
            else:
                stages.append(('conv2d', *ps, loc, scale, m.stride, m.padding, m.dilation, m.groups))
        else:
            cur.append(m)
    if cur:
        stages.append(('torch', torch.nn.Sequential(*cur)))
    step = orc.ElboStepOracle(stages, S, N_BATCHES)
    opt = torch.optim.Adam(step.parameters(), lr=1e-3)
    gen = torch.Generator().manual_seed(1)
    x, y = synthetic_batch(workload, B, gen)

    def run():
        opt.zero_grad()
        loss, _ = step.loss(x, y)
        loss.backward()
        opt.step()
        return float(loss)
    return run


def cpu_baseline(workload, bounded_seconds):
    wl = WORKLOADS[workload]
    B, S = wl["batch"], wl["samples"]
    sample = "full step (B=%d, S=%d)" % (B, S)
    if workload == "c4":            # a full C4 step is ~12 TFLOP on the CPU: time a 1-sample slice
        S, sample = 1, "one MC sample of the step (B=1024, S=1 of 32), scaled by samples*MC"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = oracle_step_fn(workload, B, S)
    run()
    t0, n = time.perf_counter(), 0
    while n < 3 or (time.perf_counter() - t0 < bounded_seconds and n < 200):
        run()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": B * S / dt, "unit": "samples*MC/s", "cores": cores, "kind": "port",
            "sample": f"{sample}, {n} steps, torch {torch.__version__} CPU fp32, {torch.get_num_threads()} threads",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (the oracle port: the reference is
    Python on torch ops and cannot be vendored) on the host cores, same config / metric / unit."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    wl = WORKLOADS[args.workload]
    B, S = wl["batch"], wl["samples"]
    sample = f"full step (B={B}, S={S})"
    if args.workload == "c4":
        S, sample = 1, "one MC sample of the step (S=1 of 32)"
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run = oracle_step_fn(args.workload, B, S)
    steps, warmup = min(args.steps, 20), min(max(args.warmup, 1), 3)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    value = B * S / dt
    print(json.dumps({
        "impl": "reference", "metric": "ELBO train samples*MC/sec", "value": value, "unit": "samples*MC/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": wl["name"], "batch_per_gpu": B, "mc_samples": S, "n_batches": N_BATCHES,
                   "optimizer": "Adam", "step": "zero_grad+forward(S)+KL+CE+backward+Adam"},
        "cpu_baseline": {"value": value, "unit": "samples*MC/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}, {steps} steps, torch {torch.__version__} CPU fp32, "
                                   f"{torch.get_num_threads()} threads"},
        "e2e": {"value": value, "unit": "samples*MC/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"],
                    help="hot-path contraction mode: tf32 (2e-3 parity class) or fp32 = 3xTF32 split (1e-5 class)")
    ap.add_argument("--parallel", default="data", choices=["data", "sample"],
                    help="N > 1: shard the batch (weak scaling, default) or the MC samples of one batch (strong scaling)")
    ap.add_argument("--loss-tail", default="batched", choices=["batched", "loop"],
                    help="likelihood term: nn.mc_mean_loss (one CE over the S*B rows) or the reference's per-sample loop")
    ap.add_argument("--no-cudnn-benchmark", action="store_true",
                    help="leave torch.backends.cudnn.benchmark off for the deterministic torch trunk (the examples' setting; "
                         "the default run lets cuDNN pick its kernels by measurement, 0.73 -> 0.69 ms per C2 step)")
    ap.add_argument("--optimizer", default="elbo-adam", choices=["adam", "elbo-adam"],
                    help="elbo-adam: bnn.optim.ELBOAdam (KL gradient + Adam in one pass, likelihood-only backward; same "
                         "trajectory); adam: torch's fused Adam on likelihood + KL, the reference loop verbatim")
    ap.add_argument("--nchw-trunk", action="store_true",
                    help="leave the deterministic torch trunk (Conv2d / BatchNorm2d / ELU) in torch's default NCHW memory "
                         "format (the examples' setting).  Default: torch.channels_last for those modules — cuDNN's NHWC "
                         "kernels without a layout conversion around each call (C2 0.64 -> 0.54 ms, C3 2.19 -> 1.91 ms); "
                         "the Bayesian layers take either layout and keep (mu, rho) row-major OIHW")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the kl_prune and cpu_baseline legs (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
