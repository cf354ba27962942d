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
    assert net._mc_plan()[0] is False                       # torch composites: the reference loop, not the batched launch
    outs = net(torch.rand(2, 1, 4, 4))
    assert isinstance(outs, list) and len(outs) == 4 and outs[0].shape == (2, 3)


@pytest.mark.gpu
def test_flipout_on_gpu_with_fused_kl():
    lin, conv, z = run_case("cuda", 2e-3)                   # cuDNN / cuBLAS may use TF32
    kl = KLDivergence(number_of_batches=3)(Net(torch.nn.Sequential(lin)))
    assert float(kl) == pytest.approx(float(z["lin_kl"]), rel=1e-5)
    kl.backward()
    assert torch.allclose(lin.weight.mean.grad.cpu(), T(z["lin_g_mean"]), rtol=2e-3, atol=2e-3)
    assert torch.allclose(lin.weight.scale.grad.cpu(), T(z["lin_g_scale"]), rtol=2e-3, atol=2e-3)
