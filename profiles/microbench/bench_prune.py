import sys, torch
sys.path.insert(0, '.')  # run from the repo root
from bayesianneuralnetworks_b200 import _C
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = 'cuda'
gen = torch.Generator(device=dev).manual_seed(5)
mus = [(torch.rand(4096, 4096, device=dev, generator=gen) * 2 - 1) / 64 for _ in range(n_t)]
rhos = [torch.randn(4096, 4096, device=dev, generator=gen) * 0.15 - 2.0 for _ in range(n_t)]
saved = [(m.clone(), r.clone()) for m, r in zip(mus, rhos)]
numel = 4096 * 4096
def run(p):
    k = int(p * numel)
    _C.prune([(m, r, k, None, None) for m, r in zip(mus, rhos)])
def restore():
    for (m, r), (sm, sr) in zip(zip(mus, rhos), saved):
        m.copy_(sm), r.copy_(sr)
run(0.75); torch.cuda.synchronize()
# correctness of tensor 0 against torch on the device
restore()
key = torch.distributions.Normal(mus[0], 1e-10 + torch.nn.functional.softplus(rhos[0]), validate_args=False).log_prob(torch.zeros((), device=dev))
kth = torch.topk(key.flatten(), int(0.75 * numel), sorted=True).values[-1]
run(0.75); torch.cuda.synchronize()
got = (rhos[0] == -30.0)
want = key >= kth
print("mask equal:", bool((got == want).all()), "count", int(got.sum()), int(want.sum()))
for p in (0.75,):
    best = 1e9
    for _ in range(reps):
        restore(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(p); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"prune p={p}: {best:.3f} ms  {(8 + 8 * p) * n_t * numel / best / 1e6:.0f} GB/s algorithmic")
# the example's sweep: successive levels on the same tensors
restore(); torch.cuda.synchronize()
tot = 0.0
for p in torch.linspace(.75, 1, 6).tolist():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(p); b.record(); torch.cuda.synchronize()
    print(f"  sweep level p={p:.2f}: {a.elapsed_time(b):.3f} ms, pruned now {int((rhos[0] == -30).sum())} (want {int(p * numel)})")
