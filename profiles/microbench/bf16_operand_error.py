"""Would bf16 operands (tcgen05 kind::f16, twice the TF32 MMA rate) keep the 2e-3 parity class at the C4 layer shape?
Numerical experiment, no kernel needed: round both operands of y = x W_s^T (M = 1024, N = K = 4096; x ~ N(0,1),
W_s = mu + sigma eps with the reference initialisation) to bf16 / TF32, contract in fp32 (cuBLAS with fp32 accumulation
of exactly representable products), compare with the fp64 result.  Metric of the parity tests: max|a - b| / max|b|."""
import torch

torch.manual_seed(0)
M, N, K = 1024, 4096, 4096
x = torch.randn(M, K, device="cuda")
mu = (torch.rand(N, K, device="cuda") * 2 - 1) / 64
sigma = torch.nn.functional.softplus(torch.randn(N, K, device="cuda") * 0.15 - 2)
w = mu + sigma * torch.randn(N, K, device="cuda")
ref = x.double() @ w.double().t()
torch.backends.cuda.matmul.allow_tf32 = False


def tf32(t):      # round to nearest, ties away (cvt.rna.tf32): 10 explicit mantissa bits
    i = t.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


for name, f in (("fp32", lambda t: t), ("tf32 (cvt.rna)", tf32), ("bf16", lambda t: t.bfloat16().float())):
    y = (f(x).double() @ f(w).double().t())          # exact products of the rounded operands
    err = (y - ref).abs()
    print(f"{name:16s} max-norm relative error {float(err.max() / ref.abs().max()):.2e}   rms relative {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.2e}")
# the weight gradient contracts over M = 32768 rows (S = 32 samples x 1024): dW = dY^T X
dy = torch.randn(4096, 4096, device="cuda")          # a 4096-row slice is enough to see the trend
xs = torch.randn(4096, K, device="cuda")
refg = dy.double().t() @ xs.double()
for name, f in (("tf32 (cvt.rna)", tf32), ("bf16", lambda t: t.bfloat16().float())):
    g = f(dy).double().t() @ f(xs).double()
    print(f"wgrad {name:16s} max-norm relative error {float((g - refg).abs().max() / refg.abs().max()):.2e}")
