import sys, time, subprocess, threading, torch
sys.path.insert(0, '.')
from bayesianneuralnetworks_b200 import _C as C
M, N, K, S = 1024, 4096, 4096, 32
g = torch.Generator(device='cuda').manual_seed(0)
a = torch.randn(S, M, K, device='cuda', generator=g)
mu = (torch.rand(N, K, device='cuda', generator=g) * 2 - 1) / 64
rho = torch.randn(N, K, device='cuda', generator=g) * 0.15 - 2
mub = torch.zeros(N, device='cuda'); rhob = torch.full((N,), -2.0, device='cuda')
sig = C.stddev(rho); sigb = C.stddev(rhob)
y = torch.empty(S, M, N, device='cuda')
dy = torch.randn(S, M, N, device='cuda', generator=g)
dmu = torch.zeros(N, K, device='cuda'); drho = torch.zeros(N, K, device='cuda')
rw, rb = C.make_rng(1, 0, 1), C.make_rng(1, 0, 2)
which = sys.argv[1]
def run():
    if which == "fwd":
        C.sampled_gemm_fwd(a, K, M * K, mu, sig, mub, sigb, None, None, C.make_view(y.data_ptr(), N, 1), M * N, M, N, K, S, 0, rw, rb, 0)
    else:
        C.sampled_gemm_wgrad(C.make_view(dy.data_ptr(), N, 1), M * N, a, K, M * K, rho, None, dmu, drho, M, N, K, S, 0, rw, 0)
for _ in range(3): run()
torch.cuda.synchronize()
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(20): run()
    n += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
p.terminate()
ms = e0.elapsed_time(e1) / n
print(which, f"{ms:.3f} ms/launch sustained, {2.0*M*N*K*S/ms/1e9:.1f} TFLOP/s")
print("clock/power samples:", rows[2:30:3])
