"""Profiling driver: bnn_prune_into (one out-of-place sweep) and bnn_prune (in place, two sweeps) on n tensors of 4096^2,
p = 0.75, CUDA-event time per call.  `python profiles/microbench/prof_prune_into.py [n_tensors]`.
Under ncu: `-k regex:prune_`."""
import sys

import torch

sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C  # noqa: E402

n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 16
gen = torch.Generator(device='cuda').manual_seed(5)
mus = [(torch.rand(4096, 4096, device='cuda', generator=gen) * 2 - 1) / 64 for _ in range(n_t)]
rhos = [torch.randn(4096, 4096, device='cuda', generator=gen) * 0.15 - 2.0 for _ in range(n_t)]
k = int(0.75 * 4096 * 4096)
pairs = n_t * 4096 * 4096
for name in ("into", "in_place"):
    best = 1e9
    for rep in range(3):
        ms = [m.clone() for m in mus]
        rs = [r.clone() for r in rhos]
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if name == "into":
            if rep == 0:
                out = [(torch.empty_like(m), torch.empty_like(r)) for m, r in zip(ms, rs)]      # caller-owned outputs
                _C.prune_into([(m, r, k, None) for m, r in zip(ms, rs)], out=out)              # workspace allocation
                torch.cuda.synchronize()
                a.record()
            outs = _C.prune_into([(m, r, k, None) for m, r in zip(ms, rs)], out=out)
        else:
            _C.prune([(m, r, k, None, None) for m, r in zip(ms, rs)])
        b.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, a.elapsed_time(b))
    print(f"{name}: {best:.3f} ms for {n_t} tensors -> {14.0 * pairs / best / 1e6:.0f} GB/s algorithmic (8 + 8p B/pair)")

if len(sys.argv) > 2 and sys.argv[2] == "--trace":
    # CUPTI timeline of ONE bnn_prune_into call (torch.profiler): start offset, duration and stream of every kernel
    import re
    import time
    from torch.profiler import ProfilerActivity, profile
    ms = [m.clone() for m in mus]
    rs = [r.clone() for r in rhos]
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        h0 = time.perf_counter()
        _C.prune_into([(m, r, k, None) for m, r in zip(ms, rs)], out=out)
        h1 = time.perf_counter()
        torch.cuda.synchronize()
    print(f"host time of the call: {(h1 - h0) * 1e6:.0f} us")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    t0 = min(e.time_range.start for e in evs)
    for e in sorted(evs, key=lambda e: e.time_range.start):
        name = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("bnn::", ""))
        dur = e.time_range.end - e.time_range.start
        if dur < 4 and "prune_" in name and "sweep" not in name and "sample" not in name and "resolve" not in name:
            continue
        print(f"{e.time_range.start - t0:9.1f} us  +{dur:8.1f} us  {name[:60]}")
