"""Generates the golden vectors under tests/golden/ by running the REFERENCE ITSELF
(/root/reference, pytorch_bayesian 0.0.4, imported unmodified) on CPU with torch fp32.

Run once in the build container:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, the vectors can.  Every random draw of the reference
(`torch.randn_like` in WeightNormal.sample, core.py:45) is recorded in call order so the same eps
can be injected into the CUDA path.

Outputs
  linear_case.npz   NormalLinear(20, 7): S=3 outputs, KL, gradients of sum(y*dy)+KL
  conv_case_*.npz   NormalConv2d configs incl. the reference's own test configs, stride/dilation/groups
  flipout_case.npz  FlipoutNormalLinear / FlipOutNormalConv2d outputs and gradients for recorded sign tensors
  mvn_case.npz      MultivariateNormalLinear output for recorded uniform draws; KLDivergence of a mixed model
  model_case.npz    small BayesianNetworkModule (Conv2d+ELU trunk, NormalConv2d, NormalLinear, Softmax):
                    S=3 ELBO loss (examples/MNIST/train.py:57-63) and all gradients
  mnist_ckpt_bayes_layers.npz  mean/scale of the Bayesian layers of examples/MNIST/mnist_pretrained.pth
  fmnist_ckpt_bayes_layers.npz same for examples/FashionMNIST/fmnist_pretrained.pth
  golden_values.json           KLDivergence values and PruneNormal mask fingerprints on the checkpoints
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
import pytorch_bayesian  # noqa: E402  (the reference)
from pytorch_bayesian.nn import (BayesianNetworkModule, KLDivergence, NormalConv2d, NormalLinear)  # noqa: E402
from pytorch_bayesian.prune import PruneNormal  # noqa: E402

assert pytorch_bayesian.__file__.startswith(REF)
torch.distributions.Distribution.set_default_validate_args(False)   # PruneNormal needs it on torch >= 1.8

_real_randn_like = torch.randn_like


class Recorder:
    """Wraps torch.randn_like (looked up at call time by core.py:45) and records every draw."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        def rec(t, *a, **k):
            e = _real_randn_like(t, *a, **k)
            self.draws.append(e.detach().clone())
            return e
        torch.randn_like = rec
        return self

    def __exit__(self, *exc):
        torch.randn_like = _real_randn_like


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def linear_case():
    torch.manual_seed(100)
    layer = NormalLinear(20, 7)
    x = torch.randn(5, 20)
    dy = torch.randn(3, 5, 7)
    with Recorder() as r:
        ys = [layer(x) for _ in range(3)]
    kl = KLDivergence(number_of_batches=4)(_wrap(layer))
    loss = sum((y * dy[s]).sum() for s, y in enumerate(ys)) + kl
    xg = x.clone().requires_grad_(True)
    loss.backward()
    # dx needs a second pass with the same eps
    it = iter(r.draws)
    torch.randn_like = lambda t, *a, **k: next(it)
    ys2 = [layer(xg) for _ in range(3)]
    torch.randn_like = _real_randn_like
    dx = torch.autograd.grad(sum((y * dy[s]).sum() for s, y in enumerate(ys2)), xg)[0]
    np.savez(os.path.join(HERE, "linear_case.npz"),
             x=np32(x), dy=np32(dy), w_mean=np32(layer.weight.mean), w_scale=np32(layer.weight.scale),
             b_mean=np32(layer.bias.mean), b_scale=np32(layer.bias.scale),
             eps_w=np.stack([np32(r.draws[2 * s]) for s in range(3)]),
             eps_b=np.stack([np32(r.draws[2 * s + 1]) for s in range(3)]),
             y=np.stack([np32(y) for y in ys]), kl=np32(kl), n_batches=4, prior_loc=0.0, prior_scale=0.1,
             g_w_mean=np32(layer.weight.mean.grad), g_w_scale=np32(layer.weight.scale.grad),
             g_b_mean=np32(layer.bias.mean.grad), g_b_scale=np32(layer.bias.scale.grad), g_x=np32(dx))


class _Net(BayesianNetworkModule):
    def __init__(self, seq, samples=1):
        super().__init__(1, 1, samples)
        self.layers = seq

    def _forward(self, x):
        return self.layers(x)


def _wrap(layer, samples=1):
    return _Net(torch.nn.Sequential(layer), samples)


def conv_case(name, cin, cout, k, stride, padding, dilation, groups, bias, hw, seed):
    torch.manual_seed(seed)
    layer = NormalConv2d(cin, cout, k, stride, padding, dilation, groups, bias)
    x = torch.randn(3, cin, *hw)
    with Recorder() as r:
        ys = [layer(x) for _ in range(2)]
    dy = torch.randn((2,) + tuple(ys[0].shape))
    kl = KLDivergence(number_of_batches=2)(_wrap(layer))
    loss = sum((y * dy[s]).sum() for s, y in enumerate(ys)) + kl
    loss.backward()
    xg = x.clone().requires_grad_(True)
    it = iter(r.draws)
    torch.randn_like = lambda t, *a, **kw: next(it)
    ys2 = [layer(xg) for _ in range(2)]
    torch.randn_like = _real_randn_like
    dx = torch.autograd.grad(sum((y * dy[s]).sum() for s, y in enumerate(ys2)), xg)[0]
    per = 2 if bias else 1
    out = dict(x=np32(x), dy=np32(dy), w_mean=np32(layer.weight.mean), w_scale=np32(layer.weight.scale),
               eps_w=np.stack([np32(r.draws[per * s]) for s in range(2)]),
               y=np.stack([np32(y) for y in ys]), kl=np32(kl), n_batches=2,
               g_w_mean=np32(layer.weight.mean.grad), g_w_scale=np32(layer.weight.scale.grad), g_x=np32(dx),
               cfg=np.array([cin, cout, k, stride, padding, dilation, groups, int(bias)]))
    if bias:
        out.update(b_mean=np32(layer.bias.mean), b_scale=np32(layer.bias.scale),
                   eps_b=np.stack([np32(r.draws[2 * s + 1]) for s in range(2)]),
                   g_b_mean=np32(layer.bias.mean.grad), g_b_scale=np32(layer.bias.scale.grad))
    np.savez(os.path.join(HERE, f"conv_case_{name}.npz"), **out)


def model_case():
    """A small network in the shape of examples/MNIST/model.py:20-33 and the loss of train.py:57-63."""
    torch.manual_seed(7)
    seq = torch.nn.Sequential(
        torch.nn.Conv2d(1, 8, 3, padding=1, stride=2), torch.nn.ELU(),
        NormalConv2d(8, 8, 3, padding=1, stride=2), torch.nn.ELU(),
        torch.nn.Flatten(), NormalLinear(8 * 3 * 3, 10), torch.nn.Softmax(dim=-1))
    net = _Net(seq, samples=3)
    x = torch.rand(6, 1, 12, 12)
    y = torch.randint(0, 10, (6,))
    with Recorder() as r:
        preds = net(x)
    kl = KLDivergence(number_of_batches=5)(net)
    ce = torch.nn.CrossEntropyLoss()
    likelihood = torch.stack([ce(p, y) for p in preds]).mean()
    loss = likelihood + kl
    loss.backward()
    out = dict(x=np32(x), y=y.numpy(), preds=np.stack([np32(p) for p in preds]), kl=np32(kl), loss=np32(loss),
               likelihood=np32(likelihood), n_batches=5)
    for k, v in net.state_dict().items():
        out["param." + k] = np32(v)
    for k, p in net.named_parameters():
        out["grad." + k] = np32(p.grad)
    # draw order per sample: conv W, conv b, linear W, linear b
    for i, nm in enumerate(["conv_w", "conv_b", "lin_w", "lin_b"]):
        out["eps." + nm] = np.stack([np32(r.draws[4 * s + i]) for s in range(3)])
    np.savez(os.path.join(HERE, "model_case.npz"), **out)


def flipout_case():
    """Flipout layers (dense.py:63-83, conv.py:145-221): outputs and gradients for recorded sign tensors R, S."""
    from pytorch_bayesian.nn import FlipoutNormalLinear, FlipOutNormalConv2d
    torch.manual_seed(300)
    lin = FlipoutNormalLinear(12, 5)
    x = torch.randn(6, 12)
    y = lin(x)
    dy = torch.randn_like(y)
    kl = KLDivergence(number_of_batches=3)(_wrap(lin))
    ((y * dy).sum() + kl).backward()
    conv = FlipOutNormalConv2d(3, 4, 3, 2, 1)
    xc = torch.randn(5, 3, 8, 8)
    yc = conv(xc)
    dyc = torch.randn_like(yc)
    (yc * dyc).sum().backward()
    np.savez(os.path.join(HERE, "flipout_case.npz"),
             lin_x=np32(x), lin_dy=np32(dy), lin_w_mean=np32(lin.weight.mean), lin_w_scale=np32(lin.weight.scale),
             lin_R=np32(lin.R), lin_S=np32(lin.S), lin_y=np32(y), lin_kl=np32(kl),
             lin_g_mean=np32(lin.weight.mean.grad), lin_g_scale=np32(lin.weight.scale.grad),
             conv_x=np32(xc), conv_dy=np32(dyc), conv_w_mean=np32(conv.weight.mean), conv_w_scale=np32(conv.weight.scale),
             conv_R=np32(conv.R), conv_S=np32(conv.S), conv_y=np32(yc),
             conv_g_mean=np32(conv.weight.mean.grad), conv_g_scale=np32(conv.weight.scale.grad))


def mvn_case():
    """MultivariateNormalLinear (dense.py:86-138, core.py:48-92) with recorded torch.rand_like draws (uniform noise),
    alone and next to a NormalLinear in KLDivergence (loss.py:24-28,38)."""
    from pytorch_bayesian.nn import MultivariateNormalLinear
    torch.manual_seed(400)
    mvn = MultivariateNormalLinear(6, 4)
    lin = NormalLinear(5, 6)
    draws = []
    real = torch.rand_like
    def rec(t, *a, **k):
        e = real(t, *a, **k)
        draws.append(e.detach().clone())
        return e
    torch.rand_like = rec
    x = torch.randn(3, 6)
    y = mvn(x)
    torch.rand_like = real
    net = _Net(torch.nn.Sequential(lin, mvn))
    kl = KLDivergence(number_of_batches=2)(net)
    kl.backward()
    np.savez(os.path.join(HERE, "mvn_case.npz"), x=np32(x), y=np32(y), u_w=np32(draws[0]), u_b=np32(draws[1]),
             w_mean=np32(mvn.weight.mean), w_scale=np32(mvn.weight.scale), b_mean=np32(mvn.bias.mean),
             b_scale=np32(mvn.bias.scale), lin_w_mean=np32(lin.weight.mean), lin_w_scale=np32(lin.weight.scale),
             lin_b_mean=np32(lin.bias.mean), lin_b_scale=np32(lin.bias.scale), kl=np32(kl),
             g_w_scale=np32(mvn.weight.scale.grad), g_lin_w_mean=np32(lin.weight.mean.grad))


def _load_example_model(example, ckpt):
    """examples/<example>/model.py BCNN with its shipped checkpoint (both example files define `model.BCNN`)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(f"_golden_{example}", os.path.join(REF, "examples", example, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = mod.BCNN(1, 10)
    model.load_state_dict(torch.load(os.path.join(REF, "examples", example, ckpt), map_location="cpu"))
    return model


def checkpoint_fixtures():
    """The Bayesian layers of the shipped checkpoints: KLDivergence(1) and PruneNormal mask fingerprints.
    MNIST: NormalConv2d + NormalLinear (weights and biases); FashionMNIST: the Flipout layers (weights only)."""
    values = {}
    for tag, example, ckpt in (("mnist", "MNIST", "mnist_pretrained.pth"), ("fmnist", "FashionMNIST", "fmnist_pretrained.pth")):
        model = _load_example_model(example, ckpt)
        conv, lin = model.layers[7], model.layers[10]
        arrays = dict(conv_w_mean=np32(conv.weight.mean), conv_w_scale=np32(conv.weight.scale),
                      lin_w_mean=np32(lin.weight.mean), lin_w_scale=np32(lin.weight.scale))
        if conv.bias is not None:
            arrays.update(conv_b_mean=np32(conv.bias.mean), conv_b_scale=np32(conv.bias.scale),
                          lin_b_mean=np32(lin.bias.mean), lin_b_scale=np32(lin.bias.scale))
        np.savez(os.path.join(HERE, f"{tag}_ckpt_bayes_layers.npz"), **arrays)
        values[tag] = {"kl_n_batches_1": float(KLDivergence(1)(model)), "prune": {}}
        for p in (0.75, 0.9):
            m = _load_example_model(example, ckpt)
            PruneNormal()(m, torch.tensor(p))           # p as a 0-dim tensor, like examples/MNIST/prune.py:49
            after = [t for l in (m.layers[7], m.layers[10]) for t in (l.weight, l.bias) if t is not None]
            masks = [(a.scale == -30) for a in after]
            values[tag]["prune"][str(p)] = {
                "counts": [int(mk.sum()) for mk in masks],
                "sha1_12": [hashlib.sha1(mk.numpy().tobytes()).hexdigest()[:12] for mk in masks]}
    with open(os.path.join(HERE, "golden_values.json"), "w") as f:
        json.dump(values, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    linear_case()
    conv_case("ref_cfg_a", 1, 1, 1, 1, 1, 1, 1, True, (10, 10), 201)     # reference tests/conftest.py:270-273
    conv_case("ref_cfg_b", 3, 4, 3, 1, 1, 1, 1, True, (10, 10), 202)     # reference tests/conftest.py:274-277
    conv_case("strided_grouped", 8, 6, 3, 2, 0, 2, 2, False, (9, 7), 203)
    conv_case("c2_shape", 64, 64, 3, 2, 1, 1, 1, True, (6, 6), 204)      # examples/MNIST/model.py:28
    model_case()
    flipout_case()
    mvn_case()
    checkpoint_fixtures()
    print("golden vectors written to", HERE)
