"""bnn_nchw_to_nhwc_bias_grad (dY NCHW -> NHWC rows + bias gradient in one pass) at the C3 and C2 conv-layer shapes:
CUDA-event time, effective bandwidth (one read + one write of dY), result checked against torch's permute."""
import sys

import torch

sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C as C  # noqa: E402

for (R, B, N, H) in ((8192, 512, 128, 4), (2048, 256, 64, 3)):
    dy = torch.randn(R, N, H, H, device='cuda')
    rho = torch.full((N,), -2.0, device='cuda')
    dm, dr = torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda')
    rng = C.make_rng(1, 0, 2)

    def run():
        return C.nchw_to_nhwc_bias_grad(dy, B, rho, None, dm, dr, 0, rng)
    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = run()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    ok = torch.equal(out, dy) and out.is_contiguous(memory_format=torch.channels_last)
    dm.zero_()
    run()
    ok = ok and torch.allclose(dm, dy.sum((0, 2, 3)), rtol=1e-4, atol=1e-3)
    print((R, B, N, H), f"{best * 1e3:.1f} us  {2 * dy.numel() * 4 / best / 1e6:.0f} GB/s", "ok" if ok else "MISMATCH")
