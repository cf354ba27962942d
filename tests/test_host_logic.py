"""CPU tests of the host-side mirror of the reference interface: traversal semantics (reference
tests/test_utils.py), containers (tests/test_nn/test_container.py), constructor contracts, state_dict
layout, the Monte-Carlo batching plan, draw bookkeeping — and that the C-ABI library loads and exports
every symbol include/bnn_b200.h declares.  No kernel is launched."""
import ctypes
import os
import re

import pytest
import torch

import bayesianneuralnetworks_b200 as bnn
from bayesianneuralnetworks_b200 import _C, runtime
from bayesianneuralnetworks_b200.nn import (BayesianModule, BayesianNetworkModule, NormalConv1d, NormalConv2d,
                                            NormalConv3d, NormalLinear, WeightNormal)
from bayesianneuralnetworks_b200.utils import _item_or_list, _pair, _single, _triple, apply_wb, traverse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Net(BayesianNetworkModule):
    def __init__(self, seq, samples=1):
        super().__init__(1, 1, samples)
        self.layers = seq

    def _forward(self, x):
        return self.layers(x)


# ------------------------------------------------------------------------------------------------ C ABI
def test_header_symbols_are_exported_by_the_library():
    header = open(os.path.join(ROOT, "include", "bnn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(bnn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert os.path.exists(_C.LIB_PATH), "libbnn_b200.so missing: run `python -m bayesianneuralnetworks_b200._build`"
    lib = ctypes.CDLL(_C.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/bnn_b200.h but not exported"
    assert declared == set(_C.EXPORTED_SYMBOLS), "ctypes binding and header disagree"
    assert _C.lib().bnn_abi_version() == 1
    assert isinstance(_C.lib().bnn_last_error_string(), bytes)


def test_binding_struct_sizes_match_the_header():
    assert ctypes.sizeof(_C.bnn_rng) == 64
    assert ctypes.sizeof(_C.bnn_view) == 24
    assert ctypes.sizeof(_C.bnn_conv2d_geom) == 64
    assert ctypes.sizeof(_C.bnn_kl_tensor) == 56
    assert ctypes.sizeof(_C.bnn_prune_tensor) == 56
    assert ctypes.sizeof(_C.bnn_prune_into_tensor) == 80
    assert ctypes.sizeof(_C.bnn_adam_tensor) == 88


def test_hot_path_rejects_cpu_tensors():
    layer = NormalLinear(3, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.zeros(1, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        layer.weight.sampled
    with pytest.raises(RuntimeError, match="CUDA"):
        bnn.prune.PruneNormal()(Net(torch.nn.Sequential(layer)), 0.5)


# ------------------------------------------------------------------------------------------------ utils
def test_item_or_list_and_ntuples():
    """reference tests/test_utils.py:6-24."""
    assert _item_or_list([1]) == 1 and _item_or_list([1, 2]) == [1, 2] and _item_or_list([]) == []
    assert _single(3) == (3,) and _pair(3) == (3, 3) and _triple(3) == (3, 3, 3)
    assert _pair((1, 2)) == (1, 2) and _pair([1, 2, 3]) == [1, 2, 3]      # iterables pass through unchecked


def test_apply_wb_semantics():
    """reference tests/test_utils.py:27-50."""
    m = NormalLinear(3, 3)
    assert apply_wb(m, lambda p: p.shape) == [(3, 3), (3,)]
    assert apply_wb(m, lambda p: None) is None
    assert apply_wb(m, lambda p, type: type, pass_type=True) == ['w', 'b']
    assert apply_wb(m, lambda p, module: module, pass_module=True) == [m, m]
    assert apply_wb(NormalLinear(3, 3, False), lambda p, type: type, pass_type=True) == ['w']
    assert apply_wb(m, lambda p, a, b=0: a + b, 1, b=2) == [3, 3]


def test_traverse_semantics():
    """reference tests/test_utils.py:53-65 and utils.py:50-67."""
    a, b = NormalLinear(3, 3), NormalConv2d(3, 4, 3)
    assert traverse(torch.nn.Linear(3, 3), lambda m: [1]) is None
    assert traverse(a, lambda m: [m]) == [a]
    assert traverse(a, lambda m: m) is None                      # non-list results are dropped
    seq = torch.nn.Sequential(a, torch.nn.ReLU(), torch.nn.Sequential(b), torch.nn.ModuleList([a]))
    assert traverse(seq, lambda m: [m]) == [a, b, a]             # registration order, nested containers
    assert traverse(torch.nn.Sequential(torch.nn.ReLU()), lambda m: [m]) is None

    class Block(torch.nn.Module):                                # not a traversed container type
        def __init__(self):
            super().__init__()
            self.inner = NormalLinear(3, 3)
    assert traverse(torch.nn.Sequential(Block()), lambda m: [m]) is None
    net = Net(torch.nn.Sequential(a, b))
    assert net.traverse(lambda m: [type(m).__name__]) == ['NormalLinear', 'NormalConv2d']


# ------------------------------------------------------------------------------------------------ containers / ctors
def test_containers():
    """reference tests/test_nn/test_container.py:16-33."""
    p, bp = torch.distributions.Normal(0, 1), torch.distributions.Normal(0, 2)
    m = BayesianModule(3, 4, p)
    assert (m.in_channels, m.out_channels, m.weight_prior, m.bias_prior) == (3, 4, p, p)
    assert BayesianModule(3, 4, p, bp).bias_prior is bp
    n = BayesianNetworkModule(3, 4, samples=7)
    assert (n.in_channels, n.out_channels, n.samples) == (3, 4, 7)
    with pytest.raises(NotImplementedError):
        n(torch.zeros(1, 3))


def test_constructor_signatures_and_state_dict_layout():
    lin = NormalLinear(3, 4, True, torch.distributions.Normal(0, 1))
    assert lin.weight.shape == (4, 3) and lin.bias.shape == (4,)
    assert float(lin.weight_prior.scale) == 1.0 and lin.bias_prior is lin.weight_prior
    quirk = NormalLinear(3, 3, torch.distributions.Normal(0, 1))      # reference tests/conftest.py:91 passes the
    assert quirk.bias is not None and float(quirk.weight_prior.scale) == pytest.approx(0.1)   # prior as `bias`
    c1, c2, c3 = NormalConv1d(4, 6, 3, groups=2), NormalConv2d(4, 6, 3, 2, 1, 1, 2, False), NormalConv3d(2, 2, (1, 2, 3))
    assert c1.weight.shape == (6, 2, 3) and c1.kernel_size == (3,) and c1.stride == (1,)
    assert c2.weight.shape == (6, 2, 3, 3) and c2.bias is None and c2.stride == (2, 2) and c2.groups == 2
    assert c3.weight.shape == (2, 2, 1, 2, 3) and c3.padding == (0, 0, 0) and not c3.transposed
    with pytest.raises(ValueError):
        NormalConv2d(3, 4, 3, groups=2)
    net = Net(torch.nn.Sequential(torch.nn.Conv2d(1, 2, 3), c2, lin))
    assert sorted(net.state_dict().keys()) == sorted([
        'layers.0.weight', 'layers.0.bias', 'layers.1.weight.mean', 'layers.1.weight.scale',
        'layers.2.weight.mean', 'layers.2.weight.scale', 'layers.2.bias.mean', 'layers.2.bias.scale'])
    # reference initialisation (dense.py:34-44): scale ~ N(-2, 0.15), |mean| <= 1/sqrt(fan_in)
    big = NormalLinear(400, 300)
    assert abs(float(big.weight.scale.mean()) + 2.0) < 0.01 and abs(float(big.weight.scale.std()) - 0.15) < 0.01
    assert float(big.weight.mean.abs().max()) <= 1 / 20 + 1e-6 and float(big.bias.mean.abs().max()) <= 1 / 20 + 1e-6


def test_draw_bookkeeping_and_partition_arithmetic():
    w = WeightNormal(3, 3)
    d0 = w._draw
    assert d0 == 1                               # the constructor's sample() (core.py:15)
    w.sample()
    assert w._last == (d0, 1) and w._draw == d0 + 1
    begin = w.advance(4, offset=8, total=16)     # rank 2 of 4, 4 samples each
    assert begin == d0 + 1 + 8 and w._last == (begin, 4) and w._draw == d0 + 17
    spec = w.draw_spec((1 << 32) + 5, 2)
    assert spec.sample_begin == 5 and spec.step == 1 and spec.tensor_id == w._tensor_id
    assert WeightNormal(2)._tensor_id != w._tensor_id
    with pytest.raises(ValueError):
        bnn.set_sample_partition(2, 2)
    with pytest.raises(ValueError):
        bnn.set_precision("bf16")


def test_mc_batching_plan():
    ok = Net(torch.nn.Sequential(torch.nn.Conv2d(1, 2, 3), torch.nn.BatchNorm2d(2), torch.nn.ELU(),
                                 NormalConv2d(2, 2, 3), torch.nn.Flatten(), NormalLinear(8, 3),
                                 torch.nn.Softmax(dim=-1)), samples=4)
    plan = ok._mc_plan()
    assert plan.ok and len(plan.bns) == 1 and not plan.probes
    ok.eval()
    assert ok._mc_plan().ok and ok._mc_plan().bns == []
    ok.layers[1].train()                          # a per-submodule toggle invalidates the cached plan
    assert len(ok._mc_plan().bns) == 1
    no_stats = Net(torch.nn.Sequential(torch.nn.BatchNorm1d(3, track_running_stats=False), NormalLinear(3, 3))).eval()
    assert len(no_stats._mc_plan().bns) == 1      # batch statistics even in eval mode: guarded
    assert not Net(torch.nn.Sequential(torch.nn.Linear(3, 3)))._mc_plan()[0]             # nothing Bayesian
    assert not Net(torch.nn.Sequential(NormalLinear(3, 3), torch.nn.Softmax(dim=0)))._mc_plan()[0]
    assert not Net(torch.nn.Sequential(torch.nn.Dropout(0.5), NormalLinear(3, 3)))._mc_plan()[0]

    class Custom(torch.nn.Module):
        def forward(self, x):
            return x.view(x.size(0), -1)
    probed = Net(torch.nn.Sequential(Custom(), NormalLinear(3, 3)))._mc_plan()
    assert probed.ok and len(probed.probes) == 1          # unknown stateless leaf: eligible, probed at run time
    bnn.nn.register_rowwise_module(Custom)
    registered = Net(torch.nn.Sequential(Custom(), NormalLinear(3, 3)))._mc_plan()
    assert registered.ok and not registered.probes
    bnn.nn.container._ROWWISE.remove(Custom)

    class Stateful(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(3))

        def forward(self, x):
            return x * self.w
    assert not Net(torch.nn.Sequential(Stateful(), NormalLinear(3, 3)))._mc_plan()[0]     # unknown code with state
    assert not Net(torch.nn.Sequential(NormalConv3d(2, 2, 1, groups=2)))._mc_plan()[0]    # grouped 3-d layers: no fused path
    assert Net(torch.nn.Sequential(NormalConv3d(1, 1, 1)))._mc_plan()[0]                  # groups == 1: lowering + sampled GEMM


# ------------------------------------------------------------------------------------------------ C ABI error behaviour
def test_c_abi_rejects_bad_arguments_without_a_gpu():
    """Argument validation happens before any device work: status codes and the thread-local message (no GPU needed)."""
    lib = _C.lib()
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)               # never dereferenced on these paths

    def msg():
        return lib.bnn_last_error_string().decode()

    assert lib.bnn_stddev(null, null, 8, null) == 1 and "NULL" in msg()
    assert lib.bnn_stddev(null, null, -1, null) == 1
    assert lib.bnn_stddev(null, null, 0, null) == 0                       # empty input is fine
    assert lib.bnn_kl(None, -1, null, null, null, null, 0, null) == 1
    assert lib.bnn_kl(None, 0, null, null, null, null, 0, null) == 0
    table = (_C.bnn_kl_tensor * 1)()
    table[0].mu, table[0].rho, table[0].numel, table[0].prior_scale = 16, 16, 4, 1.0
    assert lib.bnn_kl(table, 1, one, null, null, null, 0, null) == 5 and "workspace" in msg()      # BNN_ERR_WORKSPACE
    need = lib.bnn_kl_workspace_size(1)
    assert lib.bnn_kl(table, 1, one, null, null, ctypes.c_void_p(8), need, null) == 2              # BNN_ERR_MISALIGNED
    table[0].prior_scale = 0.0
    assert lib.bnn_kl(table, 1, one, null, null, ctypes.c_void_p(256), need, null) == 1 and "scale" in msg()
    pt = (_C.bnn_prune_tensor * 1)()
    pt[0].mu, pt[0].rho, pt[0].numel, pt[0].k = 16, 16, 10, 11
    ws = lib.bnn_prune_workspace_size(pt, 1)
    assert ws > 0
    assert lib.bnn_prune(pt, 1, ctypes.c_void_p(256), ws, null) == 1 and "k" in msg()
    rng = _C.make_rng(1, 0, 0)
    view = _C.make_view(16, 8, 1)
    # unknown precision, negative size, bias with only one of (mu_b, sigma_b)
    assert lib.bnn_sampled_gemm_fwd(one, 8, 0, one, one, null, null, null, null, view, 0, 4, 4, 8, 1, 0,
                                    ctypes.byref(rng), None, 7, null) == 1 and "precision" in msg()
    assert lib.bnn_sampled_gemm_fwd(one, 8, 0, one, one, null, null, null, null, view, 0, -4, 4, 8, 1, 0,
                                    ctypes.byref(rng), None, 0, null) == 1
    assert lib.bnn_sampled_gemm_fwd(one, 8, 0, one, one, one, null, null, null, view, 0, 4, 4, 8, 1, 0,
                                    ctypes.byref(rng), None, 0, null) == 1 and "bias" in msg()
    assert lib.bnn_sampled_gemm_fwd(one, 8, 0, one, one, null, null, null, null, view, 0, 4, 4, 8, 70000, 0,
                                    ctypes.byref(rng), None, 0, null) == 6                          # BNN_ERR_UNSUPPORTED
    geom = _C.bnn_conv2d_geom(1, 4, 8, 8, 2, 4, 3, 3, 6, 6, 1, 1, 0, 0, 1, 1)                       # c0 + Cg > C
    assert lib.bnn_im2col(one, one, ctypes.byref(geom), null) == 1 and "geometry" in msg()
    # likelihood tail: rows must be whole sample blocks, pitches must hold a row, the workspace is checked last
    ce_ws = lib.bnn_mc_cross_entropy_workspace_size()
    assert ce_ws >= 16 + 2 * 8 * 148
    assert lib.bnn_mc_cross_entropy_fwd(one, 4, one, 10, 3, 4, -100, one, one, one, one, ce_ws, null) == 1 and "sample blocks" in msg()
    assert lib.bnn_mc_cross_entropy_fwd(one, 3, one, 9, 3, 4, -100, one, one, one, one, ce_ws, null) == 1 and "pitch" in msg()
    assert lib.bnn_mc_cross_entropy_fwd(null, 4, one, 9, 3, 4, -100, one, one, one, one, ce_ws, null) == 1 and "NULL" in msg()
    assert lib.bnn_mc_cross_entropy_fwd(one, 4, one, 0, 3, 4, -100, one, one, one, one, ce_ws, null) == 1 and "empty" in msg()
    assert lib.bnn_mc_cross_entropy_bwd(one, 4, one, 9, 3, 4, -100, one, one, one, null, 4, null) == 1 and "NULL" in msg()
    assert lib.bnn_mc_cross_entropy_bwd(one, 4, one, 9, 3, 4, -100, one, one, one, one, 2, null) == 1 and "pitch" in msg()


def test_prune_workspace_plan_and_selftest_argument_checks():
    """bnn_prune_workspace_size is a pure host function: it must cover the general path's key workspace (4 B per
    element), the deferred list of small tensors (every element: 20 B each) and grow with every tensor added; the
    interval self test validates its arguments before touching the device."""
    lib = _C.lib()
    null = ctypes.c_void_p(None)

    def size(numels):
        t = (_C.bnn_prune_tensor * len(numels))()
        for i, n in enumerate(numels):
            t[i].mu, t[i].rho, t[i].numel, t[i].k = 16, 16, n, n // 2
        return lib.bnn_prune_workspace_size(t, len(numels))

    assert lib.bnn_prune_workspace_size(None, 0) == 256
    prev = 0
    for numels in ([10], [10, 4099], [10, 4099, 65536], [10, 4099, 65536, 65537], [10, 4099, 65536, 65537, 1 << 24]):
        s = size(numels)
        assert s % 256 == 0 and s > prev
        general = sum(4 * n for n in numels)
        small = sum(20 * n for n in numels if n <= 65536)
        assert s >= max(general, small) + len(numels) * (256 + 2 * 2049 * 4)
        prev = s
    assert size([1 << 24]) < 4.2 * (1 << 24) + (1 << 20)          # no more than the key workspace plus small change
    assert size([40] * 30) > size([40] * 24)                       # more than one group of descriptors
    at = (_C.bnn_adam_tensor * 1)()
    at[0].mu, at[0].rho, at[0].numel, at[0].kl_coeff, at[0].prior_scale = 16, 16, 4, 1.0, 0.1
    assert lib.bnn_adam_kl_step(at, 1, 1e-3, 0.9, 0.999, 1e-8, null, 1, null) == 1 and "moment" in lib.bnn_last_error_string().decode()
    assert lib.bnn_adam_kl_step(at, 1, 1e-3, 1.0, 0.999, 1e-8, null, 1, null) == 1                 # beta1 = 1
    assert lib.bnn_adam_kl_step(at, 1, 1e-3, 0.9, 0.999, 1e-8, null, 0, null) == 1 and "step" in lib.bnn_last_error_string().decode()
    assert lib.bnn_adam_kl_step(None, 0, 1e-3, 0.9, 0.999, 1e-8, null, 1, null) == 0
    assert lib.bnn_selftest_prune_interval(null, null, -1, null, null, 0, null) == 1
    assert lib.bnn_selftest_prune_interval(null, null, 4, null, null, 2, null) == 1
    assert lib.bnn_selftest_prune_interval(null, null, 4, null, null, 1, null) == 1 and "NULL" in lib.bnn_last_error_string().decode()
    assert lib.bnn_selftest_prune_interval(null, null, 0, null, null, 1, null) == 0


def test_mc_mean_loss_host_logic():
    """nn.mc_mean_loss == the reference loop body torch.stack([criterion(p, y) for p in preds]).mean()
    (examples/MNIST/train.py:59-61): one call over the batched rows when the list is an MCSamples of row blocks and
    the criterion is a row mean, the loop otherwise."""
    import torch.nn.functional as F
    from bayesianneuralnetworks_b200.nn import MCSamples, mc_mean_loss
    torch.manual_seed(0)
    S, B, C = 5, 7, 4
    base = torch.randn(S * B, C, requires_grad=True)
    y = torch.randint(0, C, (B,))
    preds = MCSamples(base.view(S, B, C).unbind(0))
    preds.batched = base
    assert isinstance(preds, list) and len(preds) == S
    w = torch.tensor([0.5, 1.0, 2.0, 1.5])
    for crit in (F.cross_entropy, torch.nn.CrossEntropyLoss(), torch.nn.CrossEntropyLoss(weight=w, label_smoothing=0.1),
                 torch.nn.CrossEntropyLoss(reduction='sum'), lambda p, t: F.cross_entropy(p, t) * 2):
        ref = torch.stack([crit(p, y) for p in preds]).mean()
        got = mc_mean_loss(crit, preds, y)
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
        g_ref, = torch.autograd.grad(ref, base, retain_graph=True)
        g_got, = torch.autograd.grad(got, base)
        assert torch.allclose(g_got, g_ref, rtol=1e-5, atol=1e-8)
    # regression criterion with a [B, C] target, a plain list (no batched tensor), and the bare tensor of S == 1
    t2 = torch.randn(B, C)
    assert torch.allclose(mc_mean_loss(F.mse_loss, preds, t2), torch.stack([F.mse_loss(p, t2) for p in preds]).mean())
    plain = [p.detach() for p in preds]
    assert torch.allclose(mc_mean_loss(F.cross_entropy, plain, y), torch.stack([F.cross_entropy(p, y) for p in plain]).mean())
    assert torch.equal(mc_mean_loss(F.cross_entropy, plain[0], y), F.cross_entropy(plain[0], y))


def test_fused_cross_entropy_eligibility_and_gradient_buffers():
    """Which criteria nn.mc_mean_loss may send to bnn_mc_cross_entropy (plain mean cross-entropy over class indices,
    the examples' criterion, train.py:40) and the single zeroed buffer a layer's backward accumulates into."""
    import torch.nn.functional as F
    from bayesianneuralnetworks_b200.functional import _zero_grads
    from bayesianneuralnetworks_b200.nn.elbo import _plain_cross_entropy
    assert _plain_cross_entropy(F.cross_entropy) == -100
    assert _plain_cross_entropy(torch.nn.CrossEntropyLoss()) == -100
    assert _plain_cross_entropy(torch.nn.CrossEntropyLoss(ignore_index=3)) == 3
    for other in (torch.nn.CrossEntropyLoss(weight=torch.ones(4)), torch.nn.CrossEntropyLoss(label_smoothing=0.1),
                  torch.nn.CrossEntropyLoss(reduction='sum'), torch.nn.NLLLoss(), F.nll_loss,
                  lambda p, t: F.cross_entropy(p, t)):
        assert _plain_cross_entropy(other) is None

    class Sub(torch.nn.CrossEntropyLoss):          # a subclass may override forward: not assumed to be plain
        pass
    assert _plain_cross_entropy(Sub()) is None
    grads, bias = _zero_grads((6, 4, 3, 3), 6, 'cpu')
    assert grads.shape == (2, 6, 4, 3, 3) and bias.shape == (2, 6) and not grads.any() and not bias.any()
    assert grads.is_contiguous() and bias.is_contiguous()
    assert bias.data_ptr() == grads.data_ptr() + 4 * grads.numel()         # one allocation, one fill
    assert _zero_grads(None, None, 'cpu') == (None, None)
    assert _zero_grads(None, 5, 'cpu')[0] is None and _zero_grads((2, 3), None, 'cpu')[1] is None
    with pytest.raises(RuntimeError, match="CUDA"):
        _C.mc_cross_entropy_fwd(torch.zeros(4, 3), torch.zeros(2, dtype=torch.int64))


@pytest.mark.skipif(not os.path.isdir("/root/reference/examples"), reason="the reference checkout is only mounted in the build container")
def test_reference_example_models_construct_unchanged_on_the_drop_in():
    """INTEGRATION.md option A: with `pytorch_bayesian` aliased to this package the reference's own example model files
    (examples/{MNIST,FashionMNIST,CIFAR10}/model.py, executed unmodified from the read-only checkout) import, build their
    networks, expose the reference's state_dict layout and load the shipped checkpoints.  (Forward needs the B200.)"""
    import importlib.util
    import sys
    saved = {k: sys.modules.get(k) for k in ("pytorch_bayesian", "pytorch_bayesian.nn", "pytorch_bayesian.prune",
                                             "pytorch_bayesian.utils")}
    sys.modules.update({"pytorch_bayesian": bnn, "pytorch_bayesian.nn": bnn.nn, "pytorch_bayesian.prune": bnn.prune,
                        "pytorch_bayesian.utils": bnn.utils})
    try:
        for example, cls, in_ch, ckpt in (("MNIST", "BCNN", 1, "mnist_pretrained.pth"),
                                          ("FashionMNIST", "BCNN", 1, "fmnist_pretrained.pth"),
                                          ("CIFAR10", "BCNN", 3, None)):
            path = f"/root/reference/examples/{example}/model.py"
            spec = importlib.util.spec_from_file_location(f"_ref_example_{example}", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            model = getattr(mod, cls)(in_ch, 10)
            assert isinstance(model, bnn.nn.BayesianNetworkModule)
            keys = model.state_dict().keys()
            assert any(k.endswith("weight.mean") for k in keys) and any(k.endswith("weight.scale") for k in keys)
            n_bayes = sum(isinstance(m, bnn.nn.BayesianModule) for m in model.modules())
            assert n_bayes == 2
            if ckpt and os.path.exists(f"/root/reference/examples/{example}/{ckpt}"):
                sd = torch.load(f"/root/reference/examples/{example}/{ckpt}", map_location="cpu")
                model.load_state_dict(sd)                      # same keys, same shapes
            # their own Flatten is user code without state: eligible for the batched forward, probed at run time.
            # MNIST (NormalConv2d + NormalLinear) runs on the fused kernels; the Flipout layers of FashionMNIST (f-1)
            # and the full-covariance head of CIFAR10 (f-4) are torch composites that join the batched pass
            plan = model._mc_plan()
            assert plan.ok and [type(m) for m in plan.probes] == [mod.Flatten]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_elbo_adam_plans_the_same_tensors_and_coefficients_as_kl_divergence():
    """bnn.optim.ELBOAdam folds the gradient of KLDivergence(n_batches)(model) into its step: it must see the tensors the
    reference's traverse sees (weights and biases as separate entries, loss.py:30-38), give each the coefficient
    1 / (numel * n_tensors * n_batches), and leave every other parameter to torch's Adam."""
    from torch.nn import Conv2d, ELU, Flatten, Sequential

    class Hidden(torch.nn.Module):                 # a Bayesian layer inside a plain module: invisible to traverse
        def __init__(self):
            super().__init__()
            self.inner = NormalLinear(10, 10)

        def forward(self, x):
            return self.inner(x)

    class Net(BayesianNetworkModule):
        def __init__(self):
            super().__init__(1, 10, 2)
            self.layers = Sequential(Conv2d(1, 4, 3), ELU(), NormalConv2d(4, 4, 3), Flatten(),
                                     NormalLinear(16, 10, bias=False), Hidden())

        def _forward(self, x):
            return self.layers(x)

    net = Net()
    opt = bnn.optim.ELBOAdam(net, number_of_batches=5, lr=1e-2)
    seen = [e["w"] for e in opt._var]
    assert seen == [net.layers[2].weight, net.layers[2].bias, net.layers[4].weight]
    for e in opt._var:
        assert e["coeff"] == pytest.approx(1.0 / (e["w"].mean.numel() * 3 * 5)) and (e["loc"], e["scale"]) == (0.0, pytest.approx(0.1))
    assert opt.param_groups[0]["variational"] and not opt.param_groups[1]["variational"]       # a torch Optimizer
    assert {id(p) for p in opt.param_groups[0]["params"]} == {id(p) for w in seen for p in (w.mean, w.scale)}
    others = {id(p) for p in opt.param_groups[1]["params"]}
    assert others == {id(p) for p in list(net.layers[0].parameters()) + list(net.layers[5].parameters())}
    net.layers[0].weight.grad = torch.ones_like(net.layers[0].weight)
    opt.zero_grad()
    assert net.layers[0].weight.grad is None


def test_pair_kernel_tile_plans_cover_every_row_block_exactly_once():
    """sampled_gemm_tma.cu: TilePlan.  For a sweep of (row blocks per sample, samples) the plan the launcher would use
    (bnn_debug_pair_tile_plan, host arithmetic) is decoded exactly as the kernel decodes blockIdx.x and must (1) give
    every sample a run of tiles that starts at row block 0, is contiguous and covers all of its row blocks, spilling at
    most two block pairs past the end (those rows are masked), and (2) beat the uniform grid's list schedule."""
    import ctypes
    lib = _C.lib()
    slots = 74
    n_plans = 0
    for m_blocks in list(range(5, 80)) + [128, 200]:
        for S in (3, 8, 16, 19, 30, 50, 60):
            out = (ctypes.c_int32 * 7)()
            assert lib.bnn_debug_pair_tile_plan(m_blocks, S, slots, out) == 0
            on, n_a, s1, a1, b1, a2, b2 = list(out)
            if not on:
                continue
            n_plans += 1
            n_b = s1 * b1 + (S - s1) * b2
            assert n_a == s1 * a1 + (S - s1) * a2
            tiles = {s: [] for s in range(S)}
            for t in range(n_a + n_b):                      # the kernel's decode of t = blockIdx.x >> 1
                if t < n_a:
                    first = s1 * a1
                    if t < first:
                        s, j = divmod(t, a1)
                    else:
                        q, j = divmod(t - first, a2)
                        s = s1 + q
                    tiles[s].append((j * 8, 8))
                else:
                    v, first = t - n_a, s1 * b1
                    if v < first:
                        s, j = divmod(v, b1)
                        a = a1
                    else:
                        q, j = divmod(v - first, b2)
                        s, a = s1 + q, a2
                    tiles[s].append((a * 8 + j * 6, 6))
            for s in range(S):
                run = sorted(tiles[s])
                pos = 0
                for start, size in run:
                    assert start == pos, (m_blocks, S, s, run)
                    pos += size
                assert m_blocks <= pos <= m_blocks + 5, (m_blocks, S, s, pos)      # <= 2 block pairs (+ odd block) to spare
            # list schedule: all 8-block tiles first, then the 6-block ones, against the uniform grid
            def makespan(costs):
                load = [0] * slots
                for c in costs:
                    load[load.index(min(load))] += c
                return max(load)
            uniform = S * ((((m_blocks + 1) // 2) + 3) // 4)
            assert makespan([4] * n_a + [3] * n_b) * 100 <= makespan([4] * uniform) * 92
    assert n_plans > 20
    out = (ctypes.c_int32 * 7)()
    lib.bnn_debug_pair_tile_plan(64, 16, 74, out)           # C3 conv forward: 74 + 72 tiles on 74 pair slots
    assert out[0] == 1 and out[1] + out[2] * out[4] + (16 - out[2]) * out[6] <= 148


def test_balanced_schedule_covers_every_work_item_exactly_once():
    """sampled_gemm_tma.cu: contract_pair_sk_kernel.  The segments every slot walks (bnn_debug_balanced_plan: the kernel's
    own sk_next evaluated on the host) must partition the work: forward / per-sample input gradient — every 256-row unit of
    every (sample, column tile) column in exactly one tile of 1..4 units over the whole reduction, tiles inside one column,
    never a lone unit after a full tile of the same run; summed input gradient — every k-block iteration of every
    1024-row tile in exactly one segment; the slots' loads differ by at most one item (tile cuts aside)."""
    n_cases = 0
    for m_blocks, S, gx, red_blocks, slots in [(64, 16, 1, 36, 74), (17, 5, 1, 3, 4), (11, 3, 2, 5, 7), (24, 6, 1, 9, 5),
                                               (9, 9, 1, 2, 2), (8, 32, 32, 128, 74), (5, 1, 1, 1, 3), (64, 16, 1, 36, 1)]:
        # ---- tiles of units (sum_samples = 0)
        m_units = (m_blocks + 1) // 2
        seen = {}
        loads = []
        for slot in range(slots):
            segs = _C.balanced_plan(m_blocks, S, gx, red_blocks, False, slots, slot)
            load = 0
            for lead_row0, mb_cap, y, z, i0, n in segs:
                assert 1 <= mb_cap <= 4 and i0 == 0 and n == red_blocks
                assert lead_row0 % 256 == 0 and 0 <= y < gx and 0 <= z < S
                u0 = lead_row0 // 256
                assert u0 + mb_cap <= m_units, "a tile stays inside its column"
                for u in range(u0, u0 + mb_cap):
                    key = (z, y, u)
                    assert key not in seen, f"unit {key} assigned twice"
                    seen[key] = slot
                load += mb_cap
            loads.append(load)
        assert len(seen) == S * gx * m_units
        assert max(loads) - min(loads) <= 1
        # ---- k-block ranges (sum_samples = 1): tiles of 1024 rows, all S samples in one tile
        m_tiles = (m_blocks + 7) // 8
        its = red_blocks * S
        cover = {}
        loads = []
        for slot in range(slots):
            segs = _C.balanced_plan(m_blocks, S, gx, red_blocks, True, slots, slot)
            load = 0
            for lead_row0, mb_cap, y, z, i0, n in segs:
                assert mb_cap == 4 and z == 0 and lead_row0 % 1024 == 0 and 0 <= y < gx and lead_row0 // 1024 < m_tiles
                assert 0 <= i0 and n >= 1 and i0 + n <= its
                for it in range(i0, i0 + n):
                    key = (y, lead_row0 // 1024, it)
                    assert key not in cover
                    cover[key] = slot
                load += n
            loads.append(load)
        assert len(cover) == gx * m_tiles * its
        assert max(loads) - min(loads) <= 1
        n_cases += 1
    assert n_cases == 8
    # the tile-size rule: runs of 5 / 6 / 7 units are cut 3 + 2 / 3 + 3 / 4 + 3, never 4 + 1
    segs = _C.balanced_plan(10, 1, 1, 4, False, 1, 0)               # one column of 5 units on one slot
    assert [s[1] for s in segs] == [3, 2]
    segs = _C.balanced_plan(14, 1, 1, 4, False, 1, 0)               # 7 units
    assert [s[1] for s in segs] == [4, 3]
    with pytest.raises(_C.BnnError):
        _C.balanced_plan(4, 1, 1, 1, False, 2, 2)                   # slot out of range
