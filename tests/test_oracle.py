"""CPU tests that PIN the oracle (oracle/variational_oracle.py): every function is checked against
vectors produced by the reference itself (tests/golden/make_golden.py imports /root/reference), against
the reference's own known-answer tests, and the Philox restatement against the Random123 known-answer
vectors.  No GPU needed."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import variational_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, tol=1e-6):
    b = T(b)
    return float((a.detach() - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_linear_oracle_matches_reference_vectors():
    z = np.load(os.path.join(GOLD, "linear_case.npz"))
    x = T(z["x"]).requires_grad_(True)
    leaves = [T(z[k]).clone().requires_grad_(True) for k in ("w_mean", "w_scale", "b_mean", "b_scale")]
    ys = [orc.linear_forward(x, leaves[0], leaves[1], T(z["eps_w"][s]), leaves[2], leaves[3], T(z["eps_b"][s]))
          for s in range(3)]
    for s in range(3):
        assert torch.equal(ys[s].detach(), T(z["y"][s]))            # same ATen ops: bit-identical
    kl = orc.kl_divergence([(leaves[0], leaves[1], 0.0, 0.1), (leaves[2], leaves[3], 0.0, 0.1)], int(z["n_batches"]))
    assert float(kl) == pytest.approx(float(z["kl"]), rel=1e-6)
    (sum((y * T(z["dy"][s])).sum() for s, y in enumerate(ys)) + kl).backward()
    for leaf, name in zip(leaves, ("g_w_mean", "g_w_scale", "g_b_mean", "g_b_scale")):
        assert close(leaf.grad, z[name])
    assert close(x.grad, z["g_x"])


@pytest.mark.parametrize("name", ["ref_cfg_a", "ref_cfg_b", "strided_grouped", "c2_shape"])
def test_conv_oracle_matches_reference_vectors(name):
    z = np.load(os.path.join(GOLD, f"conv_case_{name}.npz"))
    cin, cout, k, stride, padding, dilation, groups, bias = [int(v) for v in z["cfg"]]
    for s in range(2):
        y = orc.conv2d_forward(T(z["x"]), T(z["w_mean"]), T(z["w_scale"]), T(z["eps_w"][s]),
                               T(z["b_mean"]) if bias else None, T(z["b_scale"]) if bias else None,
                               T(z["eps_b"][s]) if bias else None, stride, padding, dilation, groups)
        assert close(y, z["y"][s])


def test_elbo_step_oracle_matches_reference_vectors():
    z = np.load(os.path.join(GOLD, "model_case.npz"))
    conv0 = torch.nn.Conv2d(1, 8, 3, padding=1, stride=2)
    with torch.no_grad():
        conv0.weight.copy_(T(z["param.layers.0.weight"])), conv0.bias.copy_(T(z["param.layers.0.bias"]))
    P = {k: T(z["param.layers." + k]).clone().requires_grad_(True) for k in
         ("2.weight.mean", "2.weight.scale", "2.bias.mean", "2.bias.scale", "5.weight.mean", "5.weight.scale",
          "5.bias.mean", "5.bias.scale")}
    stages = [('torch', torch.nn.Sequential(conv0, torch.nn.ELU())),
              ('conv2d', P["2.weight.mean"], P["2.weight.scale"], P["2.bias.mean"], P["2.bias.scale"], 0.0, 0.1,
               2, 1, 1, 1),
              ('torch', torch.nn.Sequential(torch.nn.ELU(), torch.nn.Flatten())),
              ('linear', P["5.weight.mean"], P["5.weight.scale"], P["5.bias.mean"], P["5.bias.scale"], 0.0, 0.1),
              ('torch', torch.nn.Softmax(dim=-1))]
    step = orc.ElboStepOracle(stages, 3, int(z["n_batches"]))
    draws = iter([T(z["eps." + n][s]) for s in range(3) for n in ("conv_w", "conv_b", "lin_w", "lin_b")])
    loss, preds = step.loss(T(z["x"]), T(z["y"]), eps_fn=lambda t: next(draws))
    for s in range(3):
        assert close(preds[s], z["preds"][s])
    assert float(loss) == pytest.approx(float(z["loss"]), rel=1e-6)
    loss.backward()
    for k, p in P.items():
        assert close(p.grad, z["grad.layers." + k]), k
    assert close(conv0.weight.grad, z["grad.layers.0.weight"])


def test_kl_and_prune_oracle_match_reference_checkpoint_values():
    z = np.load(os.path.join(GOLD, "mnist_ckpt_bayes_layers.npz"))
    gold = json.load(open(os.path.join(GOLD, "golden_values.json")))["mnist"]
    names = ["conv_w", "conv_b", "lin_w", "lin_b"]
    tensors = [(T(z[n + "_mean"]), T(z[n + "_scale"]), 0.0, 0.1) for n in names]
    assert float(orc.kl_divergence(tensors, 1)) == pytest.approx(gold["kl_n_batches_1"], rel=1e-6)
    sums = orc.kl_tensor_sums(tensors)      # SURVEY §8c per-tensor element sums
    for got, want in zip(sums, (9779.682, 1.4213, 3006.7746, 0.0793)):
        assert got == pytest.approx(want, rel=5e-4)
    for p, g in gold["prune"].items():
        for (mu, rho, _, _), h, c in zip(tensors, g["sha1_12"], g["counts"]):
            mask = orc.prune_mask(mu, rho, torch.tensor(float(p)))
            assert int(mask.sum()) == c
            assert hashlib.sha1(mask.numpy().tobytes()).hexdigest()[:12] == h
            k = orc.prune_count(torch.tensor(float(p)), mu.numel())
            assert torch.equal(orc.prune_mask_lowest_index(mu, rho, k), mask)   # k-th key unique here


def test_kl_and_prune_oracle_match_the_fashion_mnist_checkpoint_values():
    """examples/FashionMNIST/fmnist_pretrained.pth (Flipout layers: weights only) — KLDivergence(1) =
    0.289310575, per-tensor element sums 7308.5585 / 2190.8956, mask fingerprints 5ca0722d87fe / 61597c4bdbf7 (p = .75)
    and ccd4f08bf1d4 / 811c69ed4772 (p = .9)."""
    z = np.load(os.path.join(GOLD, "fmnist_ckpt_bayes_layers.npz"))
    gold = json.load(open(os.path.join(GOLD, "golden_values.json")))["fmnist"]
    assert gold["prune"]["0.75"]["sha1_12"] == ["5ca0722d87fe", "61597c4bdbf7"]
    assert gold["prune"]["0.9"]["sha1_12"] == ["ccd4f08bf1d4", "811c69ed4772"]
    tensors = [(T(z[n + "_mean"]), T(z[n + "_scale"]), 0.0, 0.1) for n in ("conv_w", "lin_w")]
    assert float(orc.kl_divergence(tensors, 1)) == pytest.approx(0.289310575, rel=1e-6)
    assert gold["kl_n_batches_1"] == pytest.approx(0.289310575, rel=1e-6)
    for got, want in zip(orc.kl_tensor_sums(tensors), (7308.5585, 2190.8956)):
        assert got == pytest.approx(want, rel=5e-4)
    for p, g in gold["prune"].items():
        for (mu, rho, _, _), h, c in zip(tensors, g["sha1_12"], g["counts"]):
            mask = orc.prune_mask(mu, rho, torch.tensor(float(p)))
            assert int(mask.sum()) == c
            assert hashlib.sha1(mask.numpy().tobytes()).hexdigest()[:12] == h


def test_reference_known_answers():
    """reference tests/test_nn/test_core.py:31-39, test_dense.py:57-70, test_conv.py:102-120."""
    rho = torch.full((4, 3), -100.0)
    sd = orc.stddev(rho)
    assert bool((sd > 0).all()) and torch.equal(sd ** 2, sd.pow(2))
    assert torch.allclose(orc.sample(torch.zeros(4, 3), rho, torch.randn(4, 3)), torch.zeros(4, 3), atol=1e-5)
    i, o = 5, 4
    y = orc.linear_forward(torch.ones(o, i), torch.ones(o, i), torch.full((o, i), -100.0), torch.randn(o, i),
                           torch.full((o,), 3.0), torch.full((o,), -100.0), torch.randn(o))
    assert torch.allclose(y, torch.full((o, o), float(i + 3)), atol=1e-5, rtol=1e-5)
    x = torch.rand(1, 3, 10, 10)
    y = orc.conv2d_forward(x, torch.ones(4, 3, 3, 3), torch.full((4, 3, 3, 3), -100.0), torch.randn(4, 3, 3, 3),
                           None, None, None, 1, 1, 1, 1)
    assert torch.allclose(y, F.conv2d(x, torch.ones(4, 3, 3, 3), None, 1, 1, 1, 1), atol=1e-5, rtol=1e-5)
    # SURVEY appendix A1
    assert float(orc.stddev(torch.tensor(-2.0))) == pytest.approx(0.126928, rel=1e-5)
    assert float(orc.stddev(torch.tensor(-30.0))) == pytest.approx(1.00094e-10, rel=1e-5)


def test_prune_count_float32_semantics():
    """int(p * numel) in float32 when p is a 0-dim tensor (examples/MNIST/prune.py:49)."""
    assert orc.prune_count(torch.tensor(0.8), 10 ** 9) == 800000000
    assert orc.prune_count(float(torch.tensor(0.8)), 10 ** 9) == 800000011
    assert orc.prune_count(0.5, 7) == 3


def test_mc_forward_list_or_tensor():
    assert isinstance(orc.mc_forward(lambda x, s: x + s, torch.zeros(2), 1), torch.Tensor)
    assert len(orc.mc_forward(lambda x, s: x + s, torch.zeros(2), 3)) == 3


def test_kl_raises_without_tensors():
    with pytest.raises(ValueError):
        orc.kl_divergence([], 1)


# ------------------------------------------------------------------------------------------------ Philox
def test_philox4x32_10_known_answer_vectors():
    """Random123 kat_vectors for philox4x32 with 10 rounds."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == want


def test_philox_eps_is_standard_normal_and_keyed():
    e = orc.philox_eps(1, 0, 2, 3, 200000)
    assert abs(e.mean()) < 0.012 and abs(e.var() - 1) < 0.02
    assert np.array_equal(e[:1000], orc.philox_eps(1, 0, 2, 3, 1000))
    assert np.array_equal(e[8:1000], orc.philox_eps(1, 0, 2, 3, 992, elem_offset=8))
    for other in (orc.philox_eps(2, 0, 2, 3, 1000), orc.philox_eps(1, 1, 2, 3, 1000), orc.philox_eps(1, 0, 3, 3, 1000),
                  orc.philox_eps(1, 0, 2, 4, 1000), orc.philox_eps(1, 1 << 32, 2, 3, 1000)):
        assert not np.array_equal(e[:1000], other)
