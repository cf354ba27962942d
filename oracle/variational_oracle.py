"""CPU oracle of the variational hot path of pytorch_bayesian (Mirko-Nava/BayesianNeuralNetworks).

TEST INFRASTRUCTURE ONLY.  This module restates, on the CPU, the arithmetic the reference performs
on its hot path, so that the CUDA library (libbnn_b200.so) can be checked against it.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it; the product package never does.

The reference is pure Python on torch ops (it has no arithmetic of its own), so the restatement
uses the same ATen CPU ops in the same order; every function cites the reference file:line
(relative to the reference checkout) it follows.  The third-party arithmetic lives in torch
(`install_requires=['torch']`, unpinned; checked here against torch 2.11.0): F.softplus, F.linear,
F.conv2d, torch.distributions.kl._kl_normal_normal, Normal.log_prob, torch.topk.

Parity pin: `tests/golden/` holds vectors produced by the reference itself (imported from
/root/reference by tests/golden/make_golden.py); tests/test_oracle.py checks every function below
against them, plus the reference's own known-answer tests (tests/test_nn/test_dense.py:57-70,
test_conv.py:78-146, test_core.py:31-39).

The Philox4x32-10 / Box-Muller stream at the bottom has no counterpart in the reference (which
calls torch.randn_like); it restates the library's own counter-based generator (Salmon et al.,
SC'11; Random123 known-answer vectors pin it) so that in-kernel eps can be checked bit-exactly at
the integer level and to ~1e-6 after the float transform.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


# --------------------------------------------------------------------------- a2 / a3: WeightNormal
def stddev(rho):
    """sigma = 1e-10 + softplus(rho)  —  pytorch_bayesian/nn/core.py:25-27."""
    return 1e-10 + F.softplus(rho)


def sample(mu, rho, eps):
    """W = mean + stddev * eps  —  pytorch_bayesian/nn/core.py:44-45 (eps = the randn_like draw)."""
    return mu + stddev(rho) * eps


# --------------------------------------------------------------------------- a4: NormalLinear
def linear_forward(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None):
    """One MC sample of NormalLinear.forward — pytorch_bayesian/nn/dense.py:46-60:
    draw W, then b (dense.py:47,51), then F.linear(x, W, b) (dense.py:60)."""
    w = sample(mu_w, rho_w, eps_w)
    b = sample(mu_b, rho_b, eps_b) if mu_b is not None else None
    return F.linear(x, w, b)


# --------------------------------------------------------------------------- a5: NormalConv2d
def conv2d_forward(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None, stride=1, padding=0,
                   dilation=1, groups=1):
    """One MC sample of NormalConv2d.forward — pytorch_bayesian/nn/conv.py:65-73,112-119."""
    w = sample(mu_w, rho_w, eps_w)
    b = sample(mu_b, rho_b, eps_b) if mu_b is not None else None
    return F.conv2d(x, w, b, stride, padding, dilation, groups)


def conv1d_forward(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None, stride=1, padding=0,
                   dilation=1, groups=1):
    """pytorch_bayesian/nn/conv.py:65-73,89-96."""
    w = sample(mu_w, rho_w, eps_w)
    b = sample(mu_b, rho_b, eps_b) if mu_b is not None else None
    return F.conv1d(x, w, b, stride, padding, dilation, groups)


# --------------------------------------------------------------------------- a6: MC loop
def mc_forward(forward_one, x, n_samples):
    """BayesianNetworkModule.forward — pytorch_bayesian/nn/container.py:32-37 with
    utils/utils.py:10-11: S sequential passes; a list for S > 1, the bare tensor for S == 1."""
    outs = [forward_one(x, s) for s in range(n_samples)]
    return outs[0] if len(outs) == 1 else outs


# --------------------------------------------------------------------------- a7: KLDivergence
def kl_normal_elementwise(mu, rho, prior_loc, prior_scale):
    """KL(N(mu, sigma) || N(loc, scale)) per element — torch/distributions/kl.py:468-471
    (_kl_normal_normal) applied to WeightNormal.dist (core.py:33-35), as called at loss.py:28."""
    sigma = stddev(rho)
    var_ratio = (sigma / prior_scale).pow(2)
    t1 = ((mu - prior_loc) / prior_scale).pow(2)
    return 0.5 * (var_ratio + t1 - 1 - var_ratio.log())


def kl_tensor_sums(tensors):
    """Per-tensor element sums in float64 (what bnn_kl returns); tensors = [(mu, rho, loc, scale)]."""
    return [float(kl_normal_elementwise(mu.double(), rho.double(), loc, scale).sum())
            for (mu, rho, loc, scale) in tensors]


def kl_divergence(tensors, n_batches=1):
    """KLDivergence.forward — pytorch_bayesian/nn/loss.py:30-38: the mean over elements of each
    tensor (loss.py:28), then the mean over the list of tensors, divided by n_batches (loss.py:38).
    `tensors` is the traversal-ordered list [(mu, rho, prior_loc, prior_scale)], weights and
    biases as separate entries (utils/utils.py:36)."""
    if not tensors:
        raise ValueError('KLDivergence was not able to find BayasianModules')   # loss.py:34-36
    per = [kl_normal_elementwise(mu, rho, loc, scale).mean() for (mu, rho, loc, scale) in tensors]
    return torch.stack(per).mean() / n_batches


# --------------------------------------------------------------------------- a9: PruneNormal
def prune_key(mu, rho):
    """Normal(mean, stddev).log_prob(0) — pytorch_bayesian/prune/prune.py:11 through
    torch/distributions/normal.py:87-102: -((v - loc)^2) / (2 var) - log(scale) - log(sqrt(2 pi))."""
    sigma = stddev(rho)
    var = sigma ** 2
    value = torch.zeros((), dtype=mu.dtype)
    return -((value - mu) ** 2) / (2 * var) - sigma.log() - LOG_SQRT_2PI


def prune_count(percentage, numel):
    """k = int(percentage * numel) — prune.py:13; float32 arithmetic when `percentage` is a tensor."""
    return int(percentage * numel)


def prune_mask(mu, rho, percentage):
    """Boolean mask of PruneNormal.prune_param — prune.py:10-15 (top-k of the flattened keys,
    scattered into a mask).  Ties at the k-th key follow torch.topk (unspecified order)."""
    keys = prune_key(mu, rho).flatten()
    k = prune_count(percentage, keys.size(0))
    _, idx = torch.topk(keys, k)
    mask = torch.zeros_like(keys).scatter(0, idx, 1).bool().view(mu.shape)
    return mask


def prune_mask_from_keys(keys, k):
    """The library's documented tie rule applied to given keys: the k largest, ties at the threshold
    broken towards the lowest element index (stable descending sort).  Equals torch.topk's set
    whenever the k-th key is unique.  Tests pass keys computed by torch ON THE SAME DEVICE as the
    kernel (the contract of SURVEY 7.3): torch's CPU kernels may round identical inputs differently
    in the vectorised body and the scalar tail, which breaks exact ties."""
    flat = keys.flatten()
    order = torch.sort(flat, descending=True, stable=True).indices[:k]
    mask = torch.zeros(flat.numel(), dtype=torch.bool, device=flat.device)
    mask[order] = True
    return mask.view(keys.shape)


def prune_mask_lowest_index(mu, rho, k):
    """prune_mask_from_keys on the CPU keys of prune_key."""
    return prune_mask_from_keys(prune_key(mu, rho), k)


def prune_apply(mu, rho, mask):
    """mean[mask] = 0; scale[mask] = -30 — prune.py:16-17 (in place)."""
    mu[mask] = 0
    rho[mask] = -30
    return mu, rho


# --------------------------------------------------------------------------- a11: gradients
def linear_grads(x, mu_w, rho_w, eps_w, mu_b, rho_b, eps_b, dy):
    """Autograd of one NormalLinear sample (the reference has no backward code):
    returns (dx, dmu_w, drho_w, dmu_b, drho_b) for upstream gradient dy."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (x, mu_w, rho_w)]
    bl = [t.detach().clone().requires_grad_(True) for t in (mu_b, rho_b)] if mu_b is not None else []
    y = linear_forward(leaves[0], leaves[1], leaves[2], eps_w, *(bl + [eps_b] if bl else []))
    grads = torch.autograd.grad(y, leaves + bl, dy)
    out = list(grads) + [None] * (5 - len(grads))
    return tuple(out)


# --------------------------------------------------------------------------- ELBO training step
class ElboStepOracle:
    """The body of the reference training loop — examples/MNIST/train.py:55-65 — restated on plain
    tensors for a network given as a list of stages.  Used as the CPU baseline (`bench.py`) and as
    the end-to-end parity oracle.  Stages:
        ('torch', module)                               deterministic torch module
        ('linear', mu_w, rho_w, mu_b, rho_b, loc, scale)                         NormalLinear
        ('conv2d', mu_w, rho_w, mu_b, rho_b, loc, scale, stride, padding, dilation, groups)
    Parameters are leaf tensors with requires_grad; eps is drawn with torch.randn_like in the
    reference's order (per sample: per layer: W then b) unless `eps_fn` supplies it.
    """

    def __init__(self, stages, n_samples, n_batches):
        self.stages = stages
        self.S = n_samples
        self.n_batches = n_batches

    def parameters(self):
        ps = []
        for st in self.stages:
            if st[0] == 'torch':
                ps += list(st[1].parameters())
            else:
                ps += [p for p in st[1:5] if p is not None]
        return ps

    def forward_one(self, x, eps_fn=None):
        draw = eps_fn if eps_fn is not None else torch.randn_like
        for st in self.stages:
            if st[0] == 'torch':
                x = st[1](x)
            elif st[0] == 'linear':
                _, mw, rw, mb, rb, _, _ = st
                ew = draw(mw)
                eb = draw(mb) if mb is not None else None
                x = linear_forward(x, mw, rw, ew, mb, rb, eb)
            else:
                _, mw, rw, mb, rb, _, _, stride, padding, dilation, groups = st
                ew = draw(mw)
                eb = draw(mb) if mb is not None else None
                x = conv2d_forward(x, mw, rw, ew, mb, rb, eb, stride, padding, dilation, groups)
        return x

    def kl(self):
        tensors = []
        for st in self.stages:
            if st[0] == 'torch':
                continue
            mw, rw, mb, rb, loc, scale = st[1:7]
            tensors.append((mw, rw, loc, scale))
            if mb is not None:
                tensors.append((mb, rb, loc, scale))
        return kl_divergence(tensors, self.n_batches)

    def loss(self, x, y, eps_fn=None):
        """train.py:57-63: preds = model(x); divergence; mean CE over the S predictions; sum."""
        preds = [self.forward_one(x, eps_fn) for _ in range(self.S)]
        divergence = self.kl()
        likelihood = torch.stack([F.cross_entropy(p, y) for p in preds]).mean()
        return likelihood + divergence, preds


# --------------------------------------------------------------------------- Philox4x32-10
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
    SC'11; Random123 philox.h).  ctr: uint32 array [..., 4]; key: uint32 array [..., 2] (broadcast).
    Returns uint32 [..., 4]."""
    ctr = np.array(ctr, dtype=np.uint32, copy=True)
    key = np.broadcast_to(np.array(key, dtype=np.uint32), ctr.shape[:-1] + (2,)).copy()
    c0, c1, c2, c3 = (ctr[..., i].copy() for i in range(4))
    k0, k1 = key[..., 0].copy(), key[..., 1].copy()
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def philox_eps(seed, step, tensor_id, sample, numel, elem_offset=0, return_bits=False):
    """eps[0:numel] of (seed, step, tensor_id, sample) exactly as include/bnn_b200.h defines it:
    key = (seed_lo, seed_hi ^ step_hi), counter = ((element + elem_offset) / 4, sample, tensor_id,
    step_lo); words (x, y) -> Box-Muller pair for elements 4g, 4g+1, words (z, w) for 4g+2, 4g+3;
    u = r * 2^-32 + 2^-33; radius = sqrt(-2 ln u1); angle = 2 pi u2 - pi; (radius cos, radius sin).
    Evaluated in float64 (the kernel uses fp32 fast intrinsics: compare at ~1e-5)."""
    first = elem_offset // 4
    last = (elem_offset + numel + 3) // 4
    g = np.arange(first, last, dtype=np.uint64)
    ctr = np.zeros((g.size, 4), dtype=np.uint32)
    ctr[:, 0] = g.astype(np.uint32)
    ctr[:, 1] = np.uint32(sample)
    ctr[:, 2] = np.uint32(tensor_id)
    ctr[:, 3] = np.uint32(step & 0xFFFFFFFF)
    key = np.array([seed & 0xFFFFFFFF, ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF], dtype=np.uint32)
    bits = philox4x32_10(ctr, key)
    # u01 as the kernel computes it: r -> fp32 (round to nearest even), then one fp32 fma
    r32 = bits.astype(np.float32).astype(np.float64)
    u = (r32 * 2.0 ** -32 + 2.0 ** -33).astype(np.float32).astype(np.float64)
    out = np.empty((g.size, 4), dtype=np.float64)
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        radius = np.sqrt(-2.0 * np.log(u[:, a]))
        angle = u[:, b] * (2.0 * np.pi) - np.pi
        out[:, o] = radius * np.cos(angle)
        out[:, o + 1] = radius * np.sin(angle)
    lo = elem_offset - first * 4
    eps = out.reshape(-1)[lo:lo + numel]
    if return_bits:
        return eps, bits
    return eps
