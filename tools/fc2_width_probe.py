"""fc2-type GEMMs (bias + layer-scale + residual) at forced tile widths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.float16
def timed(name, fn, flops, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"{name:50s} {ms:8.4f} ms {flops / ms / 1e9:8.1f} TF/s", flush=True)
for (M, N, K) in ((100352, 384, 1536), (62720, 384, 1536), (25088, 768, 3072), (14848, 768, 3072), (100352, 384, 768), (25088, 768, 1536), (401408, 192, 384)):
    a = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev); g = torch.rand(N, device=dev); x = torch.randn(M, N, device=dev).to(dt)
    for bn in (0, 256, 192, 128):
        timed(f"M{M} N{N} K{K} res bn={bn or 'auto'}", lambda: L.gemm(a, w, x, M, N, K, bias=b, gamma=g, residual=x, ldr=N, backend=(1000 + bn) if bn else 0), 2.0 * M * N * K)
