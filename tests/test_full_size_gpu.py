"""Parity at BASELINE.json's full sizes (too large for the CPU oracle): the C4 layer shape against plain torch fp32
on the same sampled weights, and size-independent properties of the KL / prune sweeps on a 4096x4096 tensor
(C5's tensor shape): additivity, threshold ordering, exact count, idempotence."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    from bayesianneuralnetworks_b200 import _C
    _C.lib()
    return _C


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def test_c4_layer_shape_against_torch_fp32(C):
    """4096x4096 NormalLinear, batch 1024, S = 2 (Philox eps): forward, data gradient and reparameterised weight
    gradients of the TF32 CTA-pair / TMA kernels against torch fp32 matmuls on the library-materialised W_s."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        M, N, K, S = 1024, 4096, 4096, 2
        g = torch.Generator(device="cuda").manual_seed(0)
        a = torch.randn(S, M, K, device="cuda", generator=g)
        mu = (torch.rand(N, K, device="cuda", generator=g) * 2 - 1) / 64
        rho = torch.randn(N, K, device="cuda", generator=g) * 0.15 - 2
        mub = (torch.rand(N, device="cuda", generator=g) * 2 - 1) / 64
        rhob = torch.randn(N, device="cuda", generator=g) * 0.15 - 2
        dy = torch.randn(S, M, N, device="cuda", generator=g)
        sig, sigb = C.stddev(rho), C.stddev(rhob)
        rw, rb = C.make_rng(77, 3, 5), C.make_rng(77, 3, 6)
        W, eps = C.materialize(mu, sig, S, 9, rw, want_eps=True)
        b = C.materialize(mub, sigb, S, 9, rb)
        y = torch.empty(S, M, N, device="cuda")
        C.sampled_gemm_fwd(a, K, M * K, mu, sig, mub, sigb, None, None, C.make_view(y.data_ptr(), N, 1), M * N, M, N, K,
                           S, 9, rw, rb, C.PREC_TF32)
        da = torch.empty(S, M, K, device="cuda")
        C.sampled_gemm_dgrad(C.make_view(dy.data_ptr(), N, 1), M * N, mu, sig, None, da, K, M * K, M, N, K, S, 9, rw,
                             C.PREC_TF32)
        dmu, drho = torch.zeros(N, K, device="cuda"), torch.zeros(N, K, device="cuda")
        C.sampled_gemm_wgrad(C.make_view(dy.data_ptr(), N, 1), M * N, a, K, M * K, rho, None, dmu, drho, M, N, K, S, 9,
                             rw, C.PREC_TF32)
        g_ref = torch.zeros(N, K, device="cuda")
        ge_ref = torch.zeros(N, K, device="cuda")
        for s in range(S):
            assert rel(y[s], a[s] @ W[s].t() + b[s]) < 2e-3
            assert rel(da[s], dy[s] @ W[s]) < 2e-3
            gs = dy[s].t() @ a[s]
            g_ref += gs
            ge_ref += gs * eps[s]
        assert rel(dmu, g_ref) < 2e-3
        assert rel(drho, ge_ref * torch.sigmoid(rho)) < 2e-3
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def test_kl_additivity_and_torch_value_at_c5_tensor_size(C):
    g = torch.Generator(device="cuda").manual_seed(1)
    mu = (torch.rand(4096, 4096, device="cuda", generator=g) * 2 - 1) / 64
    rho = torch.randn(4096, 4096, device="cuda", generator=g) * 0.15 - 2
    whole = C.kl([(mu, rho, None, None, 0.0, 0.1, 1.0)]).item()
    parts = C.kl([(mu[i * 512:(i + 1) * 512], rho[i * 512:(i + 1) * 512], None, None, 0.0, 0.1, 1.0) for i in range(8)])
    assert whole == pytest.approx(float(parts.sum()), rel=1e-9)          # fp64 accumulation: additive over row blocks
    sigma = 1e-10 + F.softplus(rho.double())
    ref = (0.5 * ((sigma / 0.1) ** 2 + (mu.double() / 0.1) ** 2 - 1 - 2 * torch.log(sigma / 0.1))).sum().item()
    assert whole == pytest.approx(ref, rel=1e-5)


@pytest.mark.parametrize("p", [0.75, 0.9])
def test_prune_properties_at_c5_tensor_size(C, prune_mode, p):
    torch.distributions.Distribution.set_default_validate_args(False)
    g = torch.Generator(device="cuda").manual_seed(2)
    mu = (torch.rand(4096, 4096, device="cuda", generator=g) * 2 - 1) / 64
    rho = torch.randn(4096, 4096, device="cuda", generator=g) * 0.15 - 2
    keys = torch.distributions.Normal(mu, 1e-10 + F.softplus(rho)).log_prob(0)
    k = int(torch.tensor(p) * mu.numel())                                # float32 arithmetic, like the reference
    m2, r2 = mu.clone(), rho.clone()
    mask = torch.empty(mu.shape, dtype=torch.uint8, device="cuda")
    C.prune([(m2, r2, k, mask, None)])
    sel = mask.bool()
    assert int(sel.sum()) == k                                           # exact count
    assert float(keys[sel].min()) >= float(keys[~sel].max())             # threshold ordering: a true top-k set
    assert bool((m2[sel] == 0).all()) and bool((r2[sel] == -30).all())
    assert torch.equal(m2[~sel], mu[~sel]) and torch.equal(r2[~sel], rho[~sel])
    # bit-exact against torch.topk on the device when the k-th key is unique
    kth = torch.topk(keys.flatten(), k).values[-1]
    if int((keys == kth).sum()) == 1:
        ref = torch.zeros(mu.numel(), dtype=torch.bool, device="cuda")
        ref[torch.topk(keys.flatten(), k).indices] = True
        assert torch.equal(sel.flatten(), ref)
    # idempotence: the pruned entries have the largest possible key and are re-selected first
    mask2 = torch.empty(mu.shape, dtype=torch.uint8, device="cuda")
    C.prune([(m2, r2, k, mask2, None)])
    assert torch.equal(mask2, mask)


def test_prune_into_equals_the_in_place_kernel_on_large_tensors_with_different_fractions(C):
    """Four tensors of 2^23 pairs with four different pruning fractions in one call: the one-sweep out-of-place kernel's
    outputs must equal the strictly in-place kernel's, inputs stay intact."""
    g = torch.Generator(device="cuda").manual_seed(8)
    n = 1 << 23
    mus = [(torch.rand(n, device="cuda", generator=g) * 2 - 1) / 64 for _ in range(4)]
    rhos = [torch.randn(n, device="cuda", generator=g) * 0.15 - 2 for _ in range(4)]
    ks = [int(p * n) for p in (0.75, 0.5, 0.9, 0.3)]
    keep = [(m.clone(), r.clone()) for m, r in zip(mus, rhos)]
    outs = C.prune_into([(m, r, k, None) for m, r, k in zip(mus, rhos, ks)])
    ref = [(m.clone(), r.clone()) for m, r in keep]
    C.prune([(m, r, k, None, None) for (m, r), k in zip(ref, ks)])
    torch.cuda.synchronize()
    for (mo, ro), (mr, rr), (m0, r0), (m, r), k in zip(outs, ref, keep, zip(mus, rhos), ks):
        assert torch.equal(m, m0) and torch.equal(r, r0)
        assert int((ro == -30).sum()) == k
        assert torch.equal(mo, mr) and torch.equal(ro, rr)
