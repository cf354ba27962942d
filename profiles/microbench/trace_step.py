"""Per-kernel device time of the captured training step (CUDA-graph replays) from CUPTI activity records
(torch.profiler): warm, in-graph durations — what the ncu launch list (cold caches, serialised) cannot show.

    python profiles/microbench/trace_step.py [c2|c3|c4] [--nchw] [--steps N]
"""
import argparse
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="c2")
    ap.add_argument("--nchw", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import bayesianneuralnetworks_b200 as bnn
    from bayesianneuralnetworks_b200.training import ElboTrainer
    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS[args.workload]
    bnn.set_precision("tf32")
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    model = bench.build_model(args.workload, wl["samples"]).to(dev)
    if not args.nchw:
        for m in model.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
                m.to(memory_format=torch.channels_last)
    tr = ElboTrainer(model, bench.N_BATCHES, graph=True)
    gen = torch.Generator().manual_seed(1)
    x, y = (t.to(dev) for t in bench.synthetic_batch(args.workload, wl["batch"], gen))
    tr.capture(x, y)
    for _ in range(5):
        tr.step(x, y)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            tr.step(x, y)
        torch.cuda.synchronize()
    tot, cnt = collections.Counter(), collections.Counter()
    first, last = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name.replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        if not name.startswith("bnn::"):
            name = re.sub(r"<.*", "", name)
        tot[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        cnt[name] += 1
        t0 = ev.time_range.start
        t1 = ev.time_range.end
        first = t0 if first is None else min(first, t0)
        last = t1 if last is None else max(last, t1)
    n = args.steps
    busy = sum(tot.values())
    print(f"{args.workload} {'NCHW' if args.nchw else 'channels_last'} trunk: {n} replays, span {(last - first) / n:.1f} us/step, "
          f"kernel time {busy / n:.1f} us/step, {sum(cnt.values()) / n:.1f} kernels/step")
    for name, t in tot.most_common(60):
        print(f"{t / n:9.2f} us/step  {cnt[name] / n:5.1f}x  {name[-90:]}")


if __name__ == "__main__":
    main()
