import sys, torch
sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C as C
from oracle import variational_oracle as orc
import numpy as np
base_mu = torch.tensor([0.0, 0.5, 1.0]).repeat(5000)
base_rho = torch.full((15000,), -2.0)
for k in (1, 4999, 5000, 5001, 7777, 14999):
    mu, rho = base_mu.clone().cuda(), base_rho.clone().cuda()
    mask = torch.empty(15000, dtype=torch.uint8, device="cuda")
    keys = torch.empty(15000, device="cuda")
    C.prune([(mu, rho, k, mask, keys)])
    ref = orc.prune_mask_lowest_index(base_mu, base_rho, k)
    m = mask.bool().cpu()
    diff = (m != ref).nonzero().flatten()
    print("k", k, "sum", int(m.sum()), "ndiff", diff.numel(), diff[:10].tolist(), "uniq keys", torch.unique(keys).tolist())
    if diff.numel():
        print("  got true idx sample", m.nonzero().flatten()[-5:].tolist(), "ref", ref.nonzero().flatten()[-5:].tolist())
def init_params(shape, gen, fan_in=None):
    fan_in = fan_in or (int(np.prod(shape[1:])) if len(shape) > 1 else shape[0])
    bound = 1.0 / np.sqrt(fan_in)
    mu = (torch.rand(shape, generator=gen) * 2 - 1) * bound
    rho = torch.randn(shape, generator=gen) * 0.15 - 2.0
    return mu, rho
g = torch.Generator().manual_seed(4)
shapes = [(64, 64, 3, 3), (64,), (10, 576), (10,)] * 8
params = [init_params(s, g) for s in shapes]
dev = [(m.cuda(), r.cuda()) for m, r in params]
for p in (0.5, 0.75):
    masks = [torch.empty(s, dtype=torch.uint8, device="cuda") for s in shapes]
    C.prune([(m, r, orc.prune_count(p, m.numel()), mk, None) for (m, r), mk in zip(dev, masks)])
    for i, ((m, r), mk, (cm, cr)) in enumerate(zip(dev, masks, params)):
        k = orc.prune_count(p, cm.numel())
        ref = orc.prune_mask_lowest_index(cm, cr, k)
        got = mk.bool().cpu()
        nd = int((got != ref).sum())
        if nd: print("p", p, "tensor", i, shapes[i], "k", k, "sum", int(got.sum()), "ndiff", nd)
        orc.prune_apply(cm, cr, ref)
        m.copy_(cm); r.copy_(cr)
print("done")
