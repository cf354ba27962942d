"""GPU parity tests of the module API (the source-compatible mirror of pytorch_bayesian.nn / .prune)
against golden vectors produced by the reference itself (tests/golden/make_golden.py) and against
the CPU oracle.  eps recorded from the reference's torch.randn_like draws is injected into the CUDA
path; tolerances: 1e-5 relative in 'fp32' mode, 2e-3 in 'tf32' (BASELINE.json north_star).
The structure follows the reference's own tests (tests/test_nn/*.py, tests/test_prune.py), run on
the device because the hot path has no CPU implementation.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bayesianneuralnetworks_b200 as bnn
from bayesianneuralnetworks_b200.nn import (BayesianConvNd, BayesianLinear, BayesianNetworkModule, KLDivergence,
                                            NormalConv1d, NormalConv2d, NormalConv3d, NormalConvNd, NormalLinear,
                                            WeightNormal)
from bayesianneuralnetworks_b200.prune import PruneNormal
from oracle import variational_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"fp32": 1e-5, "tf32": 2e-3}


def rel_err(got, ref):
    ref = torch.as_tensor(ref).double()
    scale = ref.abs().max().clamp_min(1e-30)
    return float((got.detach().double().cpu() - ref).abs().max() / scale)


@pytest.fixture(autouse=True)
def _defaults():
    bnn.set_precision("fp32")
    bnn.set_mc_batching("auto")
    bnn.set_sample_partition(0, 1)
    bnn.manual_seed(0x5EED)
    yield
    bnn.set_precision("fp32")
    bnn.set_mc_batching("auto")


class Net(BayesianNetworkModule):
    def __init__(self, seq, samples=1):
        super().__init__(1, 1, samples)
        self.layers = seq

    def _forward(self, x):
        return self.layers(x)


def load_weight(w, mean, scale):
    with torch.no_grad():
        w.mean.copy_(torch.from_numpy(mean))
        w.scale.copy_(torch.from_numpy(scale))


# ------------------------------------------------------------------------------------------------ golden: layers
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_linear_matches_reference_golden(prec):
    z = np.load(os.path.join(GOLD, "linear_case.npz"))
    bnn.set_precision(prec)
    layer = NormalLinear(20, 7)
    load_weight(layer.weight, z["w_mean"], z["w_scale"])
    load_weight(layer.bias, z["b_mean"], z["b_scale"])
    layer.cuda()
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    dy = torch.from_numpy(z["dy"]).cuda()
    ys = []
    for s in range(3):          # the reference's loop: one pass per MC sample, W drawn before b
        with bnn.injected_eps({layer.weight: torch.from_numpy(z["eps_w"][s:s + 1]),
                               layer.bias: torch.from_numpy(z["eps_b"][s:s + 1])}):
            ys.append(layer(x))
    for s in range(3):
        assert rel_err(ys[s], z["y"][s]) < TOL[prec]
    kl = KLDivergence(number_of_batches=int(z["n_batches"]))(Net(torch.nn.Sequential(layer)))
    assert float(kl) == pytest.approx(float(z["kl"]), rel=1e-5)
    loss = sum((y * dy[s]).sum() for s, y in enumerate(ys)) + kl
    loss.backward()
    assert rel_err(layer.weight.mean.grad, z["g_w_mean"]) < TOL[prec]
    assert rel_err(layer.weight.scale.grad, z["g_w_scale"]) < TOL[prec]
    assert rel_err(layer.bias.mean.grad, z["g_b_mean"]) < TOL[prec]
    assert rel_err(layer.bias.scale.grad, z["g_b_scale"]) < TOL[prec]
    assert rel_err(x.grad, z["g_x"]) < TOL[prec]


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("name", ["ref_cfg_a", "ref_cfg_b", "strided_grouped", "c2_shape"])
def test_conv2d_matches_reference_golden(name, prec):
    z = np.load(os.path.join(GOLD, f"conv_case_{name}.npz"))
    cin, cout, k, stride, padding, dilation, groups, bias = [int(v) for v in z["cfg"]]
    bnn.set_precision(prec)
    layer = NormalConv2d(cin, cout, k, stride, padding, dilation, groups, bool(bias))
    load_weight(layer.weight, z["w_mean"], z["w_scale"])
    if bias:
        load_weight(layer.bias, z["b_mean"], z["b_scale"])
    layer.cuda()
    x = torch.from_numpy(z["x"]).cuda().requires_grad_(True)
    dy = torch.from_numpy(z["dy"]).cuda()
    ys = []
    for s in range(2):
        inj = {layer.weight: torch.from_numpy(z["eps_w"][s:s + 1])}
        if bias:
            inj[layer.bias] = torch.from_numpy(z["eps_b"][s:s + 1])
        with bnn.injected_eps(inj):
            ys.append(layer(x))
    for s in range(2):
        assert ys[s].shape == z["y"][s].shape
        assert rel_err(ys[s], z["y"][s]) < TOL[prec]
    kl = KLDivergence(number_of_batches=int(z["n_batches"]))(Net(torch.nn.Sequential(layer)))
    assert float(kl) == pytest.approx(float(z["kl"]), rel=1e-5)
    (sum((y * dy[s]).sum() for s, y in enumerate(ys)) + kl).backward()
    assert rel_err(layer.weight.mean.grad, z["g_w_mean"]) < TOL[prec]
    assert rel_err(layer.weight.scale.grad, z["g_w_scale"]) < TOL[prec]
    if bias:
        assert rel_err(layer.bias.mean.grad, z["g_b_mean"]) < TOL[prec]
        assert rel_err(layer.bias.scale.grad, z["g_b_scale"]) < TOL[prec]
    assert rel_err(x.grad, z["g_x"]) < TOL[prec]


def build_model_case(z, samples=3):
    seq = torch.nn.Sequential(
        torch.nn.Conv2d(1, 8, 3, padding=1, stride=2), torch.nn.ELU(),
        NormalConv2d(8, 8, 3, padding=1, stride=2), torch.nn.ELU(),
        torch.nn.Flatten(), NormalLinear(8 * 3 * 3, 10), torch.nn.Softmax(dim=-1))
    net = Net(seq, samples)
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    net.load_state_dict(sd)          # the reference's state_dict keys load unchanged
    return net.cuda()


@pytest.mark.parametrize("batching", ["auto", "never"])
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_network_elbo_step_matches_reference_golden(prec, batching):
    """examples/MNIST/train.py:57-63 on a small network: S=3 predictions, KL, loss and every gradient."""
    z = np.load(os.path.join(GOLD, "model_case.npz"))
    bnn.set_precision(prec)
    bnn.set_mc_batching(batching)
    net = build_model_case(z)
    conv, lin = net.layers[2], net.layers[5]
    x = torch.from_numpy(z["x"]).cuda()
    y = torch.from_numpy(z["y"]).cuda()
    eps = {conv.weight: z["eps.conv_w"], conv.bias: z["eps.conv_b"], lin.weight: z["eps.lin_w"],
           lin.bias: z["eps.lin_b"]}
    if batching == "auto":
        with bnn.injected_eps({k: torch.from_numpy(v) for k, v in eps.items()}):
            preds = net(x)
    else:
        preds = []
        for s in range(3):
            with bnn.injected_eps({k: torch.from_numpy(v[s:s + 1]) for k, v in eps.items()}):
                preds.append(net(x, samples=1))
    assert isinstance(preds, list) and len(preds) == 3
    for s in range(3):
        assert rel_err(preds[s], z["preds"][s]) < TOL[prec]
    kl = KLDivergence(number_of_batches=int(z["n_batches"]))(net)
    likelihood = torch.stack([F.cross_entropy(p, y) for p in preds]).mean()
    loss = likelihood + kl
    assert float(kl) == pytest.approx(float(z["kl"]), rel=1e-5)
    assert float(loss) == pytest.approx(float(z["loss"]), rel=TOL[prec])
    loss.backward()
    for name, p in net.named_parameters():
        assert rel_err(p.grad, z["grad." + name]) < TOL[prec] * 2, name


def test_checkpoint_kl_golden():
    """KLDivergence(1) on the Bayesian layers of examples/MNIST/mnist_pretrained.pth = 0.20435977."""
    z = np.load(os.path.join(GOLD, "mnist_ckpt_bayes_layers.npz"))
    gold = json.load(open(os.path.join(GOLD, "golden_values.json")))["mnist"]
    conv, lin = NormalConv2d(64, 64, 3, padding=1, stride=2), NormalLinear(576, 10)
    load_weight(conv.weight, z["conv_w_mean"], z["conv_w_scale"])
    load_weight(conv.bias, z["conv_b_mean"], z["conv_b_scale"])
    load_weight(lin.weight, z["lin_w_mean"], z["lin_w_scale"])
    load_weight(lin.bias, z["lin_b_mean"], z["lin_b_scale"])
    net = Net(torch.nn.Sequential(conv, torch.nn.Flatten(), lin)).cuda()
    assert float(KLDivergence(1)(net)) == pytest.approx(gold["kl_n_batches_1"], rel=5e-6)
    # and PruneNormal through the module API reproduces the reference's masks (bit-exact fingerprints)
    import hashlib
    PruneNormal()(net, torch.tensor(0.75))
    for w, h, c in zip((conv.weight, conv.bias, lin.weight, lin.bias), gold["prune"]["0.75"]["sha1_12"],
                       gold["prune"]["0.75"]["counts"]):
        mask = (w.scale == -30)
        assert int(mask.sum()) == c
        assert bool((w.mean[mask] == 0).all())
        assert hashlib.sha1(mask.cpu().numpy().tobytes()).hexdigest()[:12] == h


# ------------------------------------------------------------------------------------------------ API contract
def test_weight_normal_contract():
    """reference tests/test_nn/test_core.py:14-39."""
    for shape in [(3,), (3, 4), (2, 3, 4, 5)]:
        w = WeightNormal(*shape).cuda()
        assert isinstance(w.mean, torch.nn.Parameter) and isinstance(w.scale, torch.nn.Parameter)
        assert w.mean.shape == shape and w.scale.shape == shape
        assert w.shape == shape and w.size() == shape and w.size(0) == shape[0]
        assert w.device == w.mean.device and w.requires_grad
        assert isinstance(w.dist, torch.distributions.Normal)
        torch.nn.init.constant_(w.mean, 0)
        torch.nn.init.constant_(w.scale, -100)
        assert bool((w.stddev > 0).all())
        assert torch.equal(w.stddev ** 2, w.variance)
        w.sample()
        assert isinstance(w.sampled, torch.Tensor) and w.sampled.shape == shape
        assert torch.allclose(w.sampled, torch.zeros(shape, device="cuda"), atol=1e-5, rtol=1e-5)


def test_weight_normal_sample_statistics_and_determinism():
    w = WeightNormal(512, 256).cuda()
    with torch.no_grad():
        w.mean.uniform_(-1, 1)
        w.scale.normal_(-2.0, 0.15)
    w.sample()
    a, b = w.sampled, w.sampled
    assert torch.equal(a, b)                       # same draw until sample() is called again
    w.sample()
    c = w.sampled
    assert not torch.equal(a, c)
    draws = w.materialize(w._draw, 64)              # 64 further draws of the stream
    zscore = (draws - w.mean) / w.stddev
    assert abs(float(zscore.mean())) < 5 / np.sqrt(zscore.numel())
    assert abs(float(zscore.var()) - 1) < 5 * np.sqrt(2 / zscore.numel())
    per_elem_mean = zscore.mean(0)                 # each element over 64 draws: N(0, 1/64)
    assert abs(float(per_elem_mean.var()) * 64 - 1) < 0.05


def test_sampled_is_differentiable():
    w = WeightNormal(8, 5).cuda()
    with torch.no_grad():
        w.mean.normal_()
        w.scale.normal_(-1.0, 0.3)
    w.sample()
    s = w.sampled
    g = torch.randn_like(s)
    (s * g).sum().backward()
    eps = ((s - w.mean) / w.stddev).detach()
    assert torch.allclose(w.mean.grad, g, rtol=1e-6, atol=1e-6)
    assert torch.allclose(w.scale.grad, g * eps * torch.sigmoid(w.scale), rtol=1e-4, atol=1e-5)


def test_bayesian_linear_and_conv_shapes():
    """reference tests/test_nn/test_dense.py:23-35 and test_conv.py:22-43."""
    for i, o, bias in [(3, 4, True), (5, 2, False)]:
        m = BayesianLinear(i, o, bias, WeightNormal, torch.distributions.Normal(0, 1))
        assert m.weight.shape == (o, i)
        assert (m.bias.shape == (o,)) if bias else (m.bias is None)
    m = BayesianConvNd(4, 6, (3, 3), (1, 1), (0, 0), (1, 1), False, 2, True, WeightNormal,
                       torch.distributions.Normal(0, 1))
    assert m.weight.shape == (6, 2, 3, 3) and m.bias.shape == (6,)
    assert (m.kernel_size, m.stride, m.padding, m.dilation, m.transposed, m.groups) == \
        ((3, 3), (1, 1), (0, 0), (1, 1), False, 2)
    t = BayesianConvNd(4, 6, (3,), (1,), (0,), (1,), True, 2, False, WeightNormal, None)
    assert t.weight.shape == (4, 3, 3) and t.bias is None
    with pytest.raises(ValueError):
        BayesianConvNd(3, 4, (3,), (1,), (0,), (1,), False, 2, True, WeightNormal, None)
    with pytest.raises(ValueError):
        BayesianConvNd(4, 3, (3,), (1,), (0,), (1,), False, 2, True, WeightNormal, None)


@pytest.mark.parametrize("i,o,bias", [(3, 3, True), (5, 4, False), (576, 10, True)])
def test_normal_linear_known_answer(i, o, bias):
    """reference tests/test_nn/test_dense.py:38-70: W = 1, b = 3, sigma = 1e-10, x = ones -> i (+3)."""
    layer = NormalLinear(i, o, bias).cuda()
    assert torch.distributions.kl_divergence(layer.weight_prior, torch.distributions.Normal(0, .1)) < 1e-8
    assert isinstance(layer.sampled, tuple) and len(layer.sampled) == 2
    assert isinstance(layer.sampled[0], torch.Tensor)
    assert (isinstance(layer.sampled[1], torch.Tensor)) if bias else (layer.sampled[1] is None)
    with torch.no_grad():
        layer.weight.mean.fill_(1), layer.weight.scale.fill_(-100)
        if bias:
            layer.bias.mean.fill_(3), layer.bias.scale.fill_(-100)
    x = torch.ones(o, i, device="cuda")
    want = torch.full((o, o), float(i + (3 if bias else 0)), device="cuda")
    assert torch.allclose(layer(x), want, atol=1e-5, rtol=1e-5)
    assert torch.allclose(layer(x, sample=False), want, atol=1e-5, rtol=1e-5)


CONV_CFG = [(1, 1, 1, 1, 1, 1, 1), (3, 4, 3, 1, 1, 1, 1), (4, 6, 3, 2, 1, 2, 2)]


@pytest.mark.parametrize("cfg", CONV_CFG)
@pytest.mark.parametrize("bias", [True, False])
@pytest.mark.parametrize("nd", [1, 2, 3])
def test_normal_conv_known_answer(nd, cfg, bias):
    """reference tests/test_nn/test_conv.py:71-146: ones weights (+3 bias) against F.convNd."""
    i, o, k, s, p, d, g = cfg
    cls = {1: NormalConv1d, 2: NormalConv2d, 3: NormalConv3d}[nd]
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    layer = cls(i, o, k, s, p, d, g, bias).cuda()
    assert isinstance(layer.sampled, tuple) and len(layer.sampled) == 2
    assert (isinstance(layer.sampled[1], torch.Tensor)) if bias else (layer.sampled[1] is None)
    with torch.no_grad():
        layer.weight.mean.fill_(1), layer.weight.scale.fill_(-100)
        if bias:
            layer.bias.mean.fill_(3), layer.bias.scale.fill_(-100)
    x = torch.rand((2, i) + (10,) * nd, device="cuda")
    want = conv(x, torch.ones_like(layer.weight.mean), None, s, p, d, g) + (3 if bias else 0)
    got = layer(x)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("cfg", [(3, 5, (2, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 1)), (4, 8, 3, 2, 1, 1), (2, 4, (3, 1, 2), (1, 2, 1), (2, 0, 1), (2, 1, 1))])
def test_normal_conv3d_fused_path_matches_torch_conv3d_on_injected_eps(cfg):
    """NormalConv3d (conv.py:122-142) with groups == 1: torch lowers the input (pad + unfold, differentiable) and the
    sample-and-contract kernels do the rest for all S samples in one launch.  Output and the gradients of the input, the
    weight and the bias against torch's conv3d on mean + stddev * eps with the SAME eps, single pass and inside a batched
    Monte-Carlo forward (fp32 mode, 1e-5 class)."""
    import bayesianneuralnetworks_b200 as bnn
    i, o, k, s, p, d = cfg
    torch.manual_seed(5)
    layer = NormalConv3d(i, o, k, s, p, d).cuda()
    S, B = 3, 4
    x = torch.randn(B, i, 7, 8, 9, device="cuda", requires_grad=True)
    eps = {layer.weight: torch.randn((S,) + tuple(layer.weight.shape)), layer.bias: torch.randn(S, o)}
    net = Net(torch.nn.Sequential(layer), samples=S).cuda()
    with bnn.injected_eps(eps):
        preds = net(x)
    assert isinstance(preds, list) and len(preds) == S
    dy = [torch.randn_like(q) for q in preds]
    sum((q * g).sum() for q, g in zip(preds, dy)).backward()
    xr = x.detach().double().requires_grad_(True)
    mw, rw = (t.detach().double().requires_grad_(True) for t in (layer.weight.mean, layer.weight.scale))
    mb, rb = (t.detach().double().requires_grad_(True) for t in (layer.bias.mean, layer.bias.scale))
    total = 0
    for smp in range(S):
        w = mw + (1e-10 + F.softplus(rw)) * eps[layer.weight][smp].double().cuda()
        b = mb + (1e-10 + F.softplus(rb)) * eps[layer.bias][smp].double().cuda()
        ref = F.conv3d(xr, w, b, layer.stride, layer.padding, layer.dilation)
        assert preds[smp].shape == ref.shape
        assert float((preds[smp].double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
        total = total + (ref * dy[smp].double()).sum()
    total.backward()
    for got, want in ((x.grad, xr.grad), (layer.weight.mean.grad, mw.grad), (layer.weight.scale.grad, rw.grad),
                      (layer.bias.mean.grad, mb.grad), (layer.bias.scale.grad, rb.grad)):
        assert float((got.double() - want).abs().max()) <= 2e-5 * float(want.abs().max()) + 1e-9


def test_forward_sample_false_reuses_draw():
    layer = NormalLinear(32, 16).cuda()
    x = torch.randn(8, 32, device="cuda")
    a = layer(x)
    b = layer(x, sample=False)
    c = layer(x)
    assert torch.equal(a, b) and not torch.equal(a, c)
    w, bias = layer.sampled
    assert torch.allclose(c, F.linear(x, w, bias), atol=1e-5, rtol=1e-5)   # .sampled is the draw forward used


def test_kl_divergence_contract():
    """reference tests/test_nn/test_loss.py:23-34."""
    with pytest.raises(ValueError):
        KLDivergence()(Net(torch.nn.Sequential(torch.nn.Linear(3, 3))).cuda())
    for layer in (NormalLinear(3, 3), NormalLinear(3, 4, False), NormalConv2d(3, 4, 3)):
        out = KLDivergence(number_of_batches=2)(Net(torch.nn.Sequential(layer)).cuda())
        assert isinstance(out, torch.Tensor) and out.dim() == 0 and float(out) > 0
    # invisible to traverse: a Bayesian layer nested in a plain Module (utils.py:51-52)
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.inner = NormalLinear(3, 3)
    with pytest.raises(ValueError):
        KLDivergence()(Net(torch.nn.Sequential(Block())).cuda())


def test_kl_matches_torch_distributions_on_device():
    layers = [NormalLinear(40, 30, prior=torch.distributions.Normal(0.1, 0.5)), NormalConv2d(3, 8, 3)]
    net = Net(torch.nn.Sequential(*layers)).cuda()
    got = KLDivergence(number_of_batches=3)(net)
    per = []
    for l in layers:
        for w, prior in ((l.weight, l.weight_prior), (l.bias, l.bias_prior)):
            per.append(torch.distributions.kl_divergence(
                w.dist, torch.distributions.Normal(prior.loc.cuda(), prior.scale.cuda())).mean())
    want = torch.stack(per).mean() / 3
    assert float(got) == pytest.approx(float(want), rel=1e-5)
    got.backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    net.zero_grad()
    want.backward()
    for a, p in zip(g1, net.parameters()):
        assert rel_err(a, p.grad.cpu()) < 1e-5


@pytest.mark.parametrize("p", [0.0, 0.25, 0.5, torch.tensor(0.75), 1.0])
def test_prune_normal_fraction(p):
    """reference tests/test_prune.py:7-24 (fraction of changed means), plus scale == -30 on the selected."""
    torch.manual_seed(0)
    net = Net(torch.nn.Sequential(NormalConv2d(3, 8, 3), torch.nn.Flatten(), NormalLinear(8, 5))).cuda()
    before = [w.mean.detach().clone() for w in (net.layers[0].weight, net.layers[0].bias, net.layers[2].weight,
                                                net.layers[2].bias)]
    PruneNormal()(net, p)
    after = [net.layers[0].weight, net.layers[0].bias, net.layers[2].weight, net.layers[2].bias]
    for b, w in zip(before, after):
        k = int(p * b.numel())
        changed = (w.mean != b)
        assert int(changed.sum()) == k
        assert int((w.scale == -30).sum()) == k
        assert bool((w.mean[changed] == 0).all())


# ------------------------------------------------------------------------------------------------ MC batching
def make_bn_net(samples):
    torch.manual_seed(3)
    seq = torch.nn.Sequential(
        torch.nn.Conv2d(1, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ELU(),
        NormalConv2d(6, 6, 3, padding=1, stride=2), torch.nn.ELU(), torch.nn.Flatten(),
        NormalLinear(6 * 4 * 4, 5), torch.nn.Softmax(dim=-1))
    return Net(seq, samples).cuda()


def test_batched_forward_equals_loop_and_batchnorm_statistics():
    x = torch.rand(7, 1, 8, 8, device="cuda")
    import copy
    nets = []
    proto = make_bn_net(4)
    for mode in ("auto", "never"):
        bnn.set_mc_batching(mode)
        net = copy.deepcopy(proto)       # same parameters, same Philox stream ids and draw counters
        preds = net(x)
        assert isinstance(preds, list) and len(preds) == 4 and preds[0].shape == (7, 5)
        loss = torch.stack([p.square().sum() for p in preds]).mean()
        loss.backward()
        nets.append((net, preds))
    (a, pa), (b, pb) = nets
    for s in range(4):
        assert torch.allclose(pa[s], pb[s], rtol=1e-5, atol=1e-6)     # same Philox draws either way
    scale = max(float(p.grad.abs().max()) for p in b.parameters())
    for (n1, p1), (n2, p2) in zip(a.named_parameters(), b.named_parameters()):
        # (the conv bias in front of BatchNorm has an exactly-zero gradient: compare on the global scale)
        assert float((p1.grad - p2.grad).abs().max()) < 2e-5 * max(scale, float(p2.grad.abs().max())), n1
    bn_a, bn_b = a.layers[1], b.layers[1]
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 4
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-4, atol=1e-6)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-4, atol=1e-6)


def test_single_sample_returns_tensor_and_samples_kwarg():
    net = make_bn_net(1)
    x = torch.rand(3, 1, 8, 8, device="cuda")
    assert isinstance(net(x), torch.Tensor)                # utils.py:10-11
    out = net(x, samples=5)
    assert isinstance(out, list) and len(out) == 5
    assert not torch.equal(out[0], out[1])


def test_batchnorm_after_bayesian_layer_falls_back_to_loop():
    torch.manual_seed(4)
    seq = torch.nn.Sequential(NormalLinear(6, 6), torch.nn.BatchNorm1d(6), NormalLinear(6, 3))
    net = Net(seq, 3).cuda()
    out = net(torch.rand(5, 6, device="cuda"))
    assert isinstance(out, list) and len(out) == 3 and out[0].shape == (5, 3)
    assert int(net.layers[1].num_batches_tracked) == 3
    assert net._mc_plan()[0] is False


def test_sample_partition_reproduces_single_process_stream():
    """rank r of R evaluates global samples [r*S/R, (r+1)*S/R); the union is the 1-GPU result."""
    x = torch.rand(4, 1, 8, 8, device="cuda")
    net = make_bn_net(4).eval()
    state = [(w, w._draw) for w in net.modules() if isinstance(w, WeightNormal)]
    with torch.no_grad():
        full = net(x)
        parts = []
        for rank in range(2):
            for w, d in state:
                w._draw = d
            bnn.set_sample_partition(rank, 2)
            parts += net(x)
        bnn.set_sample_partition(0, 1)
    for s in range(4):
        assert torch.equal(full[s], parts[s])


def test_cpu_tensors_are_rejected_loudly():
    layer = NormalLinear(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        KLDivergence()(Net(torch.nn.Sequential(layer)))


# ------------------------------------------------------------------------------------------------ CUDA graphs
def test_graph_safe_rng_fresh_eps_per_replay_and_consistent_backward():
    """A captured step redraws eps on every replay (device-side Philox step counter); forward, backward and
    `.sampled` of one step use the same draw."""
    from bayesianneuralnetworks_b200 import runtime
    bnn.graph_safe_rng(True)
    try:
        layer = NormalLinear(64, 32).cuda()
        x = torch.randn(16, 64, device="cuda")
        counter = runtime.step_counter(x.device)
        counter.zero_()
        # eager semantics of the counter
        y0 = layer(x)
        y0_again = layer(x, sample=False)
        bnn.advance_rng_step()
        y1 = layer(x, sample=False)           # same host draw index, next device step -> different eps
        assert torch.equal(y0, y0_again) and not torch.equal(y0, y1)
        w, b = layer.sampled                  # materialised with the current counter value
        assert torch.allclose(y1, F.linear(x, w, b), atol=1e-5, rtol=1e-5)
        # captured forward + backward (static input, gradients of the parameters).  A fresh layer: autograd ties a
        # parameter's gradient accumulation to the stream of its first use, which must not be the legacy stream.
        layer = NormalLinear(64, 32).cuda()
        sx = x.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                layer.zero_grad(set_to_none=True)
                layer(sx).square().sum().backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        layer.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            bnn.advance_rng_step()
            out = layer(sx)
            out.square().sum().backward()
        outs = []
        for _ in range(3):
            g.replay()
            torch.cuda.synchronize()
            outs.append(out.detach().clone())
            w, b = layer.sampled
            # the weights this replay used: out = x W^T + b; dL/dmean = 2 out^T x; dL/dscale = dL/dmean * eps * sigmoid
            assert torch.allclose(out, F.linear(sx, w, b), atol=1e-4, rtol=1e-4)
            gw = 2 * out.detach().t() @ sx
            assert torch.allclose(layer.weight.mean.grad, gw, atol=1e-3, rtol=1e-3)
            eps = (w - layer.weight.mean) / layer.weight.stddev
            assert torch.allclose(layer.weight.scale.grad, gw * eps * torch.sigmoid(layer.weight.scale), atol=2e-3,
                                  rtol=2e-3)
        assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    finally:
        bnn.graph_safe_rng(False)


# ------------------------------------------------------------------------------------------------ example flow
class _ExampleFlatten(torch.nn.Module):
    """The user-defined Flatten of the reference's examples (examples/MNIST/model.py:6-12)."""

    def forward(self, x):
        return x.view(x.size(0), -1)


def _example_bcnn(samples):
    """examples/MNIST/model.py:15-36 at a reduced width (same layer types and order)."""
    from torch.nn import BatchNorm2d, Conv2d, ELU, Softmax
    seq = torch.nn.Sequential(
        Conv2d(1, 8, 5, padding=2, stride=2), BatchNorm2d(8), ELU(), Conv2d(8, 8, 3, padding=1), ELU(),
        Conv2d(8, 16, 3, padding=0, stride=2), ELU(), NormalConv2d(16, 16, 3, padding=1, stride=2), ELU(),
        _ExampleFlatten(), NormalLinear(16 * 3 * 3, 10), Softmax(dim=-1))
    return Net(seq, samples)


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 4e-3)])
def test_c1_mlp_elbo_step_matches_the_oracle(prec, tol):
    """BASELINE.json configs[0]: Bayesian MLP 784-400-400-10, batch 128, one MC sample,
    ELBO = NLL + KL, the whole step (forward, KL, cross-entropy, backward) against the CPU oracle's restatement of
    examples/MNIST/train.py:55-65 with the same injected eps; 1e-5 class in fp32 mode, 2e-3 class in TF32 mode."""
    from oracle import variational_oracle as orc
    bnn.set_precision(prec)
    torch.manual_seed(11)
    B, S = 128, 1

    class MLP(BayesianNetworkModule):
        def __init__(self):
            super().__init__(784, 10, S)
            self.layers = torch.nn.Sequential(torch.nn.Flatten(), NormalLinear(784, 400), torch.nn.ELU(),
                                              NormalLinear(400, 400), torch.nn.ELU(), NormalLinear(400, 10),
                                              torch.nn.Softmax(dim=-1))

        def _forward(self, x):
            return self.layers(x)

    net = MLP()
    lins = [net.layers[1], net.layers[3], net.layers[5]]
    x, y = torch.rand(B, 1, 28, 28), torch.randint(0, 10, (B,))
    tensors = [w for l in lins for w in (l.weight, l.bias)]
    eps = {w: torch.randn((S,) + tuple(w.shape)) for w in tensors}
    P = {w: (w.mean.detach().clone().requires_grad_(True), w.scale.detach().clone().requires_grad_(True)) for w in tensors}
    stages = [('torch', torch.nn.Flatten())]
    for i, l in enumerate(lins):
        stages.append(('linear', *P[l.weight], *P[l.bias], 0.0, 0.1))
        stages.append(('torch', torch.nn.ELU() if i < 2 else torch.nn.Softmax(dim=-1)))
    order = iter([eps[w][s] for s in range(S) for w in tensors])
    ref_loss, ref_pred = orc.ElboStepOracle(stages, S, 469).loss(x, y, eps_fn=lambda t: next(order))
    ref_loss.backward()
    net.cuda()
    with bnn.injected_eps(eps):
        pred = net(x.cuda())
    assert torch.is_tensor(pred) and pred.shape == (B, 10)              # S == 1: the bare tensor (utils.py:10-11)
    loss = F.cross_entropy(pred, y.cuda()) + KLDivergence(number_of_batches=469)(net)
    loss.backward()

    def rel(a, b):
        return float((a.detach().cpu().double() - b.detach().double()).abs().max() / b.detach().abs().max().clamp_min(1e-30))
    assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss))
    ref_pred = ref_pred[0] if isinstance(ref_pred, (list, tuple)) else ref_pred
    assert rel(pred, ref_pred) < tol
    for w in tensors:
        assert rel(w.mean.grad, P[w][0].grad) < 2 * tol and rel(w.scale.grad, P[w][1].grad) < 2 * tol


def test_elbo_adam_equals_adam_on_likelihood_plus_kl():
    """SURVEY §8f-3, optimizer half: ELBOAdam (likelihood back-propagated, KL gradient + Adam in one pass of
    bnn_adam_kl_step) walks the same trajectory as the reference's torch.optim.Adam on likelihood + KLDivergence
    (examples/MNIST/train.py:43,57-65), for the variational tensors and the deterministic trunk alike."""
    import copy
    bnn.set_precision("fp32")
    torch.manual_seed(5)
    ref_model = _example_bcnn(3).cuda()
    bnn.nn.register_rowwise_module(_ExampleFlatten)
    new_model = copy.deepcopy(ref_model)
    x = torch.rand(16, 1, 28, 28, device="cuda")
    y = torch.arange(16, device="cuda") % 10
    kld = KLDivergence(number_of_batches=7)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=2e-3)
    new_opt = bnn.optim.ELBOAdam(new_model, number_of_batches=7, lr=2e-3)

    def rewind(model, step):            # the same eps streams for both models: identical tensor ids, identical draws
        ids = iter(range(10_000, 20_000))
        for m in model.modules():
            if isinstance(m, bnn.nn.WeightNormal):
                m._tensor_id, m._draw = next(ids), 3 * step

    for step in range(6):
        rewind(ref_model, step), rewind(new_model, step)
        ref_opt.zero_grad()
        loss = bnn.nn.mc_mean_loss(F.cross_entropy, ref_model(x), y) + kld(ref_model)
        loss.backward()
        ref_opt.step()
        new_opt.zero_grad()
        bnn.nn.mc_mean_loss(F.cross_entropy, new_model(x), y).backward()
        new_opt.step()
    for (n, a), (_, b) in zip(ref_model.named_parameters(), new_model.named_parameters()):
        if n == "layers.0.bias":        # a conv bias in front of BatchNorm: its gradient is rounding noise (~1e-9), which
            continue                    # Adam normalises into lr-sized steps of random sign — not comparable
        assert torch.allclose(a, b, rtol=2e-4, atol=2e-6), (n, float((a - b).abs().max()))
    assert float(kld(new_model)) == pytest.approx(float(kld(ref_model)), rel=1e-5)


def test_mc_mean_loss_equals_the_reference_loop_body():
    """the batched likelihood term (one criterion call over the S*B rows of the batched Monte-Carlo
    forward) gives the loss and every gradient of torch.stack([CE(p, y) for p in preds]).mean() (train.py:59-61)."""
    bnn.set_precision("fp32")
    torch.manual_seed(0)
    model = _example_bcnn(6).cuda()
    bnn.nn.register_rowwise_module(_ExampleFlatten)
    x = torch.rand(32, 1, 28, 28, device="cuda")
    y = torch.arange(32, device="cuda") % 10
    grads = []
    for form in ("loop", "batched"):
        for m in model.modules():                              # identical eps streams for both forms: rewind the draws
            if isinstance(m, bnn.nn.WeightNormal):
                m._draw = 0
        model.zero_grad()
        preds = model(x)
        assert isinstance(preds, list) and len(preds) == 6 and preds.batched.shape[0] == 6 * 32
        if form == "loop":
            loss = torch.stack([F.cross_entropy(p, y) for p in preds]).mean()
        else:
            loss = bnn.nn.mc_mean_loss(F.cross_entropy, preds, y)
        loss.backward()
        grads.append((float(loss), [p.grad.clone() for p in model.parameters()]))
    assert grads[0][0] == pytest.approx(grads[1][0], rel=1e-6)
    for a, b in zip(grads[0][1], grads[1][1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_example_training_loop_learns_and_prunes(prec):
    """The body of examples/MNIST/train.py:53-65 and examples/MNIST/prune.py:47-50 on synthetic, separable data:
    the ELBO goes down, accuracy goes up, pruning keeps the fractions — first on the reference loop
    (set_mc_batching('never')), then on the batched Monte-Carlo forward, which the user-defined Flatten joins after
    being probed at run time."""
    from bayesianneuralnetworks_b200.nn import container
    if _ExampleFlatten in container._ROWWISE:               # registered by another test
        container._ROWWISE.remove(_ExampleFlatten)
    bnn.set_precision(prec)
    torch.manual_seed(0)
    model = _example_bcnn(4).cuda()
    bnn.set_mc_batching('never')
    protos = torch.rand(10, 1, 28, 28, device="cuda")
    y = torch.arange(64, device="cuda") % 10
    x = (protos[y] + 0.1 * torch.randn(64, 1, 28, 28, device="cuda")).clamp(0, 1)
    kld = KLDivergence(number_of_batches=100)
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)

    def step():
        opt.zero_grad()
        preds = model(x)
        loss = torch.stack([F.cross_entropy(p, y) for p in preds]).mean() + kld(model)
        loss.backward()
        opt.step()
        acc = (torch.stack(preds).mean(0).argmax(-1) == y).float().mean()
        return float(loss), float(acc)

    first = step()
    for _ in range(10):
        step()
    bnn.set_mc_batching('auto')
    for _ in range(60):
        last = step()
    plan = model._mc_plan()
    assert plan.ok and plan.verified and plan.probes and plan.probes[0].__dict__['_bnn_rowwise'] is True
    assert last[0] < first[0] - 0.1 and last[1] > 0.5, (first, last)
    sd = model.state_dict()
    assert "layers.7.weight.mean" in sd and "layers.10.bias.scale" in sd
    for p in torch.linspace(.75, 1, 3):                     # examples/MNIST/prune.py:49-50 (p is a 0-dim tensor)
        PruneNormal()(model, p)
        w = model.layers[7].weight
        assert int((w.scale == -30).sum()) == int(p * w.mean.numel())
    model.eval()
    with torch.no_grad():
        out = model(x)
    assert len(out) == 4 and torch.isfinite(out[0]).all()
