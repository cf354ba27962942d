"""Flipout layers against vectors produced by the reference (tests/golden/make_golden.py): torch
composites, so they run on the CPU as well; on the GPU their KL goes through the fused kernel."""
import os

import numpy as np
import pytest
import torch

from bayesianneuralnetworks_b200.nn import (BayesianNetworkModule, FlipoutNormalLinear, FlipOutNormalConv1d,
                                            FlipOutNormalConv2d, FlipOutNormalConv3d, KLDivergence, NormalLinear,
                                            WeightNormal)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "flipout_case.npz")


class Net(BayesianNetworkModule):
    def __init__(self, seq, samples=1):
        super().__init__(1, 1, samples)
        self.layers = seq

    def _forward(self, x):
        return self.layers(x)


def T(a):
    return torch.from_numpy(np.asarray(a))


def run_case(device, tol):
    z = np.load(GOLD)
    lin = FlipoutNormalLinear(12, 5)
    conv = FlipOutNormalConv2d(3, 4, 3, 2, 1)
    with torch.no_grad():
        lin.weight.mean.copy_(T(z["lin_w_mean"])), lin.weight.scale.copy_(T(z["lin_w_scale"]))
        conv.weight.mean.copy_(T(z["conv_w_mean"])), conv.weight.scale.copy_(T(z["conv_w_scale"]))
    lin.to(device), conv.to(device)
    lin.R, lin.S = T(z["lin_R"]).to(device), T(z["lin_S"]).to(device)
    conv.R, conv.S = T(z["conv_R"]).to(device), T(z["conv_S"]).to(device)
    y = lin(T(z["lin_x"]).to(device), sample=False)
    yc = conv(T(z["conv_x"]).to(device), sample=False)
    assert torch.allclose(y.detach().cpu(), T(z["lin_y"]), rtol=tol, atol=tol)
    assert torch.allclose(yc.detach().cpu(), T(z["conv_y"]), rtol=tol, atol=tol * 10)
    (y * T(z["lin_dy"]).to(device)).sum().backward()
    (yc * T(z["conv_dy"]).to(device)).sum().backward()
    return lin, conv, z


def test_flipout_matches_reference_vectors_cpu():
    lin, conv, z = run_case("cpu", 1e-6)
    # gradients of the data term (the golden gradient of the linear layer includes the KL term: checked on the GPU)
    assert torch.allclose(conv.weight.mean.grad, T(z["conv_g_mean"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(conv.weight.scale.grad, T(z["conv_g_scale"]), rtol=1e-5, atol=1e-5)


def test_flipout_contract():
    """reference tests/test_nn/test_dense.py:73-88, test_conv.py:149-175."""
    lin = FlipoutNormalLinear(3, 4, torch.distributions.Normal(0, 1))
    assert isinstance(lin, NormalLinear) and isinstance(lin.weight, WeightNormal) and lin.bias is None
    assert lin.weight.shape == (4, 3) and float(lin.weight_prior.scale) == 1.0
    assert isinstance(lin.sampled, tuple) and len(lin.sampled) == 2
    assert lin.sampled[0].shape == (4,) and lin.sampled[1].shape == (3,)
    assert set(lin.sampled[0].unique().tolist()) <= {-1.0, 1.0}
    for cls, nd in ((FlipOutNormalConv1d, 1), (FlipOutNormalConv2d, 2), (FlipOutNormalConv3d, 3)):
        conv = cls(3, 4, 3, 1, 1)
        out = conv(torch.rand((2, 3) + (6,) * nd))
        assert out.shape == (2, 4) + (6,) * nd and conv.bias is None
        assert conv.sampled[0].shape == (2, 4) + (1,) * nd and conv.sampled[1].shape == (2, 3) + (1,) * nd
    net = Net(torch.nn.Sequential(FlipOutNormalConv2d(1, 2, 3), torch.nn.Flatten(), FlipoutNormalLinear(8, 3)), samples=4)
    assert net._mc_plan()[0] is True                        # torch composites join the batched pass (on CUDA inputs)
    outs = net(torch.rand(2, 1, 4, 4))                      # a CPU input takes the reference loop
    assert isinstance(outs, list) and len(outs) == 4 and outs[0].shape == (2, 3)


@pytest.mark.gpu
def test_flipout_on_gpu_with_fused_kl():
    lin, conv, z = run_case("cuda", 2e-3)                   # cuDNN / cuBLAS may use TF32
    kl = KLDivergence(number_of_batches=3)(Net(torch.nn.Sequential(lin)))
    assert float(kl) == pytest.approx(float(z["lin_kl"]), rel=1e-5)
    kl.backward()
    assert torch.allclose(lin.weight.mean.grad.cpu(), T(z["lin_g_mean"]), rtol=2e-3, atol=2e-3)
    assert torch.allclose(lin.weight.scale.grad.cpu(), T(z["lin_g_scale"]), rtol=2e-3, atol=2e-3)


@pytest.mark.gpu
def test_composite_layers_join_the_batched_monte_carlo_forward():
    """Flipout and full-covariance layers are torch composites, but they take part in the batched Monte-Carlo forward
    (trunk once, S*B rows through every later layer, S noise draws per layer in one call).  Checked against the
    reference loop: identical in the noise-free limit, same first and second moments with noise."""
    import bayesianneuralnetworks_b200 as bnn
    from bayesianneuralnetworks_b200.nn import MultivariateNormalLinear
    torch.manual_seed(3)
    S, B = 6, 4
    net = Net(torch.nn.Sequential(torch.nn.Conv2d(1, 2, 3, padding=1), torch.nn.ELU(),
                                  FlipOutNormalConv2d(2, 2, 3, padding=1), torch.nn.ELU(), torch.nn.Flatten(),
                                  FlipoutNormalLinear(32, 5), torch.nn.ELU(), MultivariateNormalLinear(5, 3)),
              samples=S).cuda()
    assert net._mc_plan()[0] is True
    x = torch.rand(B, 1, 4, 4, device="cuda")
    saved = {n: p.detach().clone() for n, p in net.named_parameters()}
    try:
        with torch.no_grad():                               # noise-free limit: sigma ~ 1e-10
            for n, p in net.named_parameters():
                if n.endswith(".scale"):
                    p.fill_(-100.0)
        bnn.set_mc_batching("always")
        batched = net(x)
        bnn.set_mc_batching("never")
        loop = net(x)
        assert isinstance(batched, list) and len(batched) == S and batched[0].shape == (B, 3)
        assert batched.batched.shape == (S * B, 3)
        for a, b in zip(batched, loop):
            assert torch.allclose(a, b, rtol=5e-3, atol=2e-3)         # cuDNN / cuBLAS may use TF32 and pick other kernels
        with torch.no_grad():
            for n, p in net.named_parameters():
                p.copy_(saved[n])
        # with noise: moments over many Monte-Carlo samples agree between the two evaluation orders
        n_mc = 1500
        bnn.set_mc_batching("always")
        big = torch.stack(net(x, samples=n_mc))
        bnn.set_mc_batching("never")
        ref = torch.stack([torch.stack(net(x, samples=50)) for _ in range(n_mc // 50)]).flatten(0, 1)
        assert big.shape == ref.shape == (n_mc, B, 3)
        assert float(big.std(0).mean()) > 1e-3              # the samples do differ
        se = ref.std(0) / n_mc ** 0.5
        assert bool(((big.mean(0) - ref.mean(0)).abs() < 6 * se + 1e-4).all())
        assert torch.allclose(big.std(0), ref.std(0), rtol=0.15, atol=1e-4)
        # gradients flow to every variational parameter through the batched pass
        bnn.set_mc_batching("always")
        net.zero_grad()
        torch.stack(net(x)).square().mean().backward()
        for n, p in net.named_parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n
    finally:
        bnn.set_mc_batching("auto")


# ------------------------------------------------------------------------------------------------ full covariance (f-4)
def _load_mvn_case(device):
    from bayesianneuralnetworks_b200.nn import MultivariateNormalLinear
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mvn_case.npz"))
    mvn, lin = MultivariateNormalLinear(6, 4), NormalLinear(5, 6)
    with torch.no_grad():
        mvn.weight.mean.copy_(T(z["w_mean"])), mvn.weight.scale.copy_(T(z["w_scale"]))
        mvn.bias.mean.copy_(T(z["b_mean"])), mvn.bias.scale.copy_(T(z["b_scale"]))
        lin.weight.mean.copy_(T(z["lin_w_mean"])), lin.weight.scale.copy_(T(z["lin_w_scale"]))
        lin.bias.mean.copy_(T(z["lin_b_mean"])), lin.bias.scale.copy_(T(z["lin_b_scale"]))
    return z, mvn.to(device), lin.to(device)


def test_multivariate_normal_linear_matches_reference_vectors_cpu():
    """Reference quirks included: uniform noise, elementwise sqrt of the triangular matrix (core.py:68-69,91)."""
    from bayesianneuralnetworks_b200.nn import MultivariateNormalLinear, WeightMultivariateNormal
    z, mvn, _ = _load_mvn_case("cpu")
    assert isinstance(mvn.weight, WeightMultivariateNormal) and mvn.weight.shape == (4, 6) and mvn.weight.scale.shape == (4, 6, 6)
    assert isinstance(mvn.sampled, tuple) and len(mvn.sampled) == 2
    draws = iter([T(z["u_w"]), T(z["u_b"])])
    real = torch.rand_like
    torch.rand_like = lambda t, *a, **k: next(draws)
    try:
        y = mvn(T(z["x"]))
    finally:
        torch.rand_like = real
    assert torch.allclose(y, T(z["y"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(mvn(T(z["x"]), sample=False), T(z["y"]), rtol=1e-6, atol=1e-6)
    fresh = MultivariateNormalLinear(5, 3, False)
    assert fresh.bias is None and bool((fresh.weight.scale[:, 0, 1:] == -100).all())        # dense.py:106-109


@pytest.mark.gpu
def test_kl_of_mixed_model_with_full_covariance_layer_on_gpu():
    """loss.py:24-28,38: the MVN tensors go through torch.distributions, the factorised ones through the fused kernel;
    the mean is taken over all four listed tensors."""
    z, mvn, lin = _load_mvn_case("cuda")
    kl = KLDivergence(number_of_batches=2)(Net(torch.nn.Sequential(lin, mvn)))
    assert float(kl) == pytest.approx(float(z["kl"]), rel=1e-5)
    kl.backward()
    assert torch.allclose(mvn.weight.scale.grad.cpu(), T(z["g_w_scale"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(lin.weight.mean.grad.cpu(), T(z["g_lin_w_mean"]), rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("shape", [(64, 576, 10), (256, 128, 128), (33, 52, 7)])
def test_fused_flipout_linear_equals_the_two_contraction_composite(shape, prec, tol):
    """FlipoutNormalLinear on CUDA is ONE sampled contraction with the rank-one sign noise eps = R S^T (dense.py:63-83
    rewritten as x (mean + stddev o R S^T)^T): output and the gradients of x, mean and scale against the reference's
    two-contraction formula evaluated by torch in fp64 with the same signs — single pass and a batched pass of S = 5
    Monte-Carlo samples."""
    import bayesianneuralnetworks_b200 as bnn
    from bayesianneuralnetworks_b200.nn import flipout
    B, K, N = shape
    torch.manual_seed(4)
    bnn.set_precision(prec)
    try:
        lin = FlipoutNormalLinear(K, N).cuda()
        x = torch.randn(B, K, device="cuda", requires_grad=True)
        dy = torch.randn(B, N, device="cuda")
        y = lin(x)
        y.backward(dy)
        R, S = lin.R.double(), lin.S.double()
        xd = x.detach().double().requires_grad_(True)
        mu = lin.weight.mean.detach().double().requires_grad_(True)
        rho = lin.weight.scale.detach().double().requires_grad_(True)
        sigma = 1e-10 + torch.nn.functional.softplus(rho)
        ref = xd.matmul(mu.t()) + (xd * S).matmul(sigma.t()) * R
        ref.backward(dy.double())

        def close(a, b):
            return float((a.double() - b).abs().max()) <= tol * float(b.abs().max()) + 1e-12
        assert close(y, ref) and close(x.grad, xd.grad)
        assert close(lin.weight.mean.grad, mu.grad) and close(lin.weight.scale.grad, rho.grad)
        # sample=False reuses the draw
        assert torch.equal(lin(x, sample=False), y)
        # the composite (set_fused_flipout_linear(False)) gives the same numbers for the same signs
        flipout.set_fused_flipout_linear(False)
        assert close(lin(x, sample=False), ref)
        flipout.set_fused_flipout_linear(True)
        # batched Monte-Carlo pass: S sign pairs, one launch; every sample against the formula with ITS signs
        net = Net(torch.nn.Sequential(lin), samples=5).cuda()
        preds = net(x.detach())
        assert len(preds) == 5
        Rs, Ss = lin._signs
        for s in range(5):
            want = xd.detach().matmul(mu.detach().t()) + (xd.detach() * Ss[s].double()).matmul(sigma.detach().t()) * Rs[s].double()
            assert close(preds[s], want)
    finally:
        bnn.set_precision("fp32")
        flipout.set_fused_flipout_linear(True)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("tf32", 2e-3)])
@pytest.mark.parametrize("case", [
    # cls, in, out, kernel, stride, padding, dilation, spatial, batch
    (FlipOutNormalConv2d, 64, 64, 3, 2, 1, 1, (6, 6), 48),       # the FashionMNIST example layer's geometry
    (FlipOutNormalConv2d, 32, 40, 3, 1, 1, 2, (9, 7), 10),       # dilation, ragged channel count
    (FlipOutNormalConv1d, 32, 16, 5, 1, 2, 1, (33,), 12),        # 1-d as a height-1 2-d convolution
])
def test_flipout_conv_runs_on_the_contraction_kernels(case, prec, tol):
    """FlipOutNormalConv1d / 2d on CUDA (groups == 1, in_channels % 32 == 0): conv(x, mean) + conv(x * S, stddev) * R
    (conv.py:207-221) with both convolutions on the library's implicit-GEMM contraction kernels (injected eps = 0):
    output and the gradients of x, mean and scale against the same formula evaluated by torch in fp64 with the same
    per-example signs; sample=False reuses the signs; the torch composite gives the same numbers; the library was
    actually called (launch counter)."""
    import bayesianneuralnetworks_b200 as bnn
    from bayesianneuralnetworks_b200 import _C
    from bayesianneuralnetworks_b200.nn import flipout
    cls, cin, cout, k, stride, padding, dilation, spatial, B = case
    op = torch.nn.functional.conv2d if len(spatial) == 2 else torch.nn.functional.conv1d
    torch.manual_seed(8)
    bnn.set_precision(prec)
    try:
        layer = cls(cin, cout, k, stride, padding, dilation).cuda()
        x = torch.randn(B, cin, *spatial, device="cuda", requires_grad=True)
        before = _C.launch_count
        y = layer(x)
        assert _C.launch_count > before, "the contraction kernels were not launched"
        dy = torch.randn(y.shape, device="cuda")
        y.backward(dy)
        R, S = layer.R.double(), layer.S.double()
        assert R.shape[:2] == (B, cout) and S.shape[:2] == (B, cin)            # per example (conv.py:154-161)
        xd = x.detach().double().requires_grad_(True)
        mu = layer.weight.mean.detach().double().requires_grad_(True)
        rho = layer.weight.scale.detach().double().requires_grad_(True)
        sigma = 1e-10 + torch.nn.functional.softplus(rho)
        ref = op(xd, mu, None, stride, padding, dilation)
        ref = ref + op(xd * S, sigma, None, stride, padding, dilation) * R
        ref.backward(dy.double())

        def close(a, b):
            return float((a.double() - b).abs().max()) <= tol * float(b.abs().max()) + 1e-12
        assert y.shape == ref.shape
        assert close(y, ref) and close(x.grad, xd.grad)
        assert close(layer.weight.mean.grad, mu.grad) and close(layer.weight.scale.grad, rho.grad)
        assert torch.equal(layer(x, sample=False), y)
        flipout.set_flipout_conv_kernels(False)            # torch's composite (cuDNN, TF32 by default): same signs, same result
        comp = layer(x, sample=False)
        assert float((comp.double() - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    finally:
        bnn.set_precision("fp32")
        flipout.set_flipout_conv_kernels(True)
