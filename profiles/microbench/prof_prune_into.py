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
            outs = _C.prune_into([(m, r, k, None) for m, r in zip(ms, rs)])
        else:
            _C.prune([(m, r, k, None, None) for m, r in zip(ms, rs)])
        b.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, a.elapsed_time(b))
    print(f"{name}: {best:.3f} ms for {n_t} tensors -> {14.0 * pairs / best / 1e6:.0f} GB/s algorithmic (8 + 8p B/pair)")
