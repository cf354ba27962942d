import sys, torch
sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C as C
M, N, K, S = 1024, 4096, 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = torch.Generator(device='cuda').manual_seed(0)
a = torch.randn(S, M, K, device='cuda', generator=g)
mu = (torch.rand(N, K, device='cuda', generator=g) * 2 - 1) / 64
rho = torch.randn(N, K, device='cuda', generator=g) * 0.15 - 2
mub = torch.zeros(N, device='cuda'); rhob = torch.full((N,), -2.0, device='cuda')
sig = C.stddev(rho); sigb = C.stddev(rhob)
y = torch.empty(S, M, N, device='cuda')
dy = torch.randn(S, M, N, device='cuda', generator=g)
da = torch.empty(S, M, K, device='cuda')
dmu = torch.zeros(N, K, device='cuda'); drho = torch.zeros(N, K, device='cuda')
rw, rb = C.make_rng(1, 0, 1), C.make_rng(1, 0, 2)
def run():
    C.sampled_gemm_fwd(a, K, M * K, mu, sig, mub, sigb, None, None, C.make_view(y.data_ptr(), N, 1), M * N, M, N, K, S, 0, rw, rb, 0)
    C.sampled_gemm_dgrad(C.make_view(dy.data_ptr(), N, 1), M * N, mu, sig, None, da, K, M * K, M, N, K, S, 0, rw, 0)
    C.sampled_gemm_wgrad(C.make_view(dy.data_ptr(), N, 1), M * N, a, K, M * K, rho, None, dmu, drho, M, N, K, S, 0, rw, 0)
for _ in range(2):
    run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
C.sampled_gemm_fwd(a, K, M * K, mu, sig, mub, sigb, None, None, C.make_view(y.data_ptr(), N, 1), M * N, M, N, K, S, 0, rw, rb, 0)
ev[1].record()
C.sampled_gemm_dgrad(C.make_view(dy.data_ptr(), N, 1), M * N, mu, sig, None, da, K, M * K, M, N, K, S, 0, rw, 0)
ev[2].record()
C.sampled_gemm_wgrad(C.make_view(dy.data_ptr(), N, 1), M * N, a, K, M * K, rho, None, dmu, drho, M, N, K, S, 0, rw, 0)
ev[3].record()
torch.cuda.synchronize()
fl = 2.0 * M * N * K * S
for i, nm in enumerate(["fwd", "dgrad", "wgrad"]):
    t = ev[i].elapsed_time(ev[i + 1]) * 1e-3
    print(f"{nm}: {t*1e3:.3f} ms  {fl/t/1e12:.1f} TFLOP/s")
import ctypes
lib = C.lib()
try:
    f = lib.bnn_debug_wait_counters
    f.restype = ctypes.c_int; f.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    buf = (ctypes.c_ulonglong * 8)()
    f(buf, 1)
    for nm, fn in (("fwd", lambda: C.sampled_gemm_fwd(a, K, M * K, mu, sig, mub, sigb, None, None, C.make_view(y.data_ptr(), N, 1), M * N, M, N, K, S, 0, rw, rb, 0)),
                   ("dgrad", lambda: C.sampled_gemm_dgrad(C.make_view(dy.data_ptr(), N, 1), M * N, mu, sig, None, da, K, M * K, M, N, K, S, 0, rw, 0))):
        fn(); torch.cuda.synchronize()
        if f(buf, 1) == 0 and buf[5]:
            n = buf[5]
            print(f"{nm}: CTAs {n} kernel {buf[0]/n:.0f} cyc/CTA | MMA waits full_w {buf[1]/n*(2 if 'pair' in sys.argv else 1):.0f} full_a {buf[2]/n:.0f} | gen waits empty_w {buf[3]/n:.0f} | TMA waits empty_a {buf[4]/n:.0f}")
except AttributeError:
    pass
