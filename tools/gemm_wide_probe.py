"""A/B timing of the wide (256 x 384) pair tiles: run with GCV_GEMM_WIDE=0 / 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.float16
for (M, N, K) in ((100352, 384, 1536), (62720, 384, 1536), (100352, 384, 768), (25088, 768, 3072)):
    a = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    bias, gamma = torch.randn(N, device=dev), torch.rand(N, device=dev)
    x = torch.randn(M, N, device=dev).to(dt)
    f = lambda: L.gemm(a, w, x, M, N, K, bias=bias, gamma=gamma, residual=x, ldr=N)
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"M{M} N{N} K{K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TF/s")
