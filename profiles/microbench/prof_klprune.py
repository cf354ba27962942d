import sys, torch
sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ONLY_PRUNE = len(sys.argv) > 2
dev = 'cuda'
gen = torch.Generator(device=dev).manual_seed(5)
mus = [(torch.rand(4096, 4096, device=dev, generator=gen) * 2 - 1) / 64 for _ in range(n_t)]
rhos = [torch.randn(4096, 4096, device=dev, generator=gen) * 0.15 - 2.0 for _ in range(n_t)]
gm = [torch.empty_like(m) for m in mus]; gr = [torch.empty_like(m) for m in mus]
for it in range(2):
    _C.kl([(m, r, None, None, 0.0, 0.1, 1.0) for m, r in zip(mus, rhos)])
    _C.kl([(m, r, a, b, 0.0, 0.1, 1.0) for m, r, a, b in zip(mus, rhos, gm, gr)])
k = int(0.75 * 4096 * 4096)
_C.prune([(m, r, k, None, None) for m, r in zip(mus, rhos)])
torch.cuda.synchronize()
print("ok")
