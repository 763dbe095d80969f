"""The VAE mu GEMM (M=256, N=12544, K=25088, reparameterisation epilogue) in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.float16
M, N, K = 256, 12544, 25088
a = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
b = torch.randn(N, device=dev); eps = torch.randn(M, N, device=dev); z = torch.empty(M * 49, 256, device=dev, dtype=dt)
flush = torch.empty(256 * 2 ** 20, device=dev, dtype=torch.uint8)
def run(): L.gemm(a, w, z, M, N, K, bias=b, eps=eps, eps_c=256, eps_hw=49, ldd=N)
for _ in range(2): run()
ts = []
for _ in range(5):
    flush.zero_(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ms = sorted(ts)[len(ts) // 2]
print(f"mu GEMM: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TF/s  weight stream {N * K * 2 / ms / 1e9:.2f} TB/s")
