"""Stage 2/3 MLP GEMM shapes: our kernel (GCV_DEBUG=0/1/2 via env) next to cuBLAS (torch.matmul) on the same shape."""
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
    print(f"{name:56s} {ms:8.4f} ms {flops / ms / 1e9:8.1f} TF/s", flush=True)
print("GCV_DEBUG =", os.environ.get("GCV_DEBUG", "0"))
for (M, C) in ((100352, 384), (62720, 384), (25088, 768), (14848, 768)):
    a = torch.randn(M, C, device=dev).to(dt); w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
    b1 = torch.randn(4 * C, device=dev); hid = torch.empty(M, 4 * C, device=dev, dtype=dt)
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt); b2 = torch.randn(C, device=dev)
    g = torch.rand(C, device=dev); x = torch.randn(M, C, device=dev).to(dt)
    fl = 2.0 * M * 4 * C * C
    timed(f"fc1 M{M} N{4*C} K{C} bias+gelu", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU), fl)
    timed(f"fc1 M{M} N{4*C} K{C} bias", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1), fl)
    if os.environ.get("GCV_DEBUG", "0") == "0":
        timed(f"  cuBLAS same shape", lambda: torch.matmul(a, w1.t(), out=hid), fl)
    timed(f"fc2 M{M} N{C} K{4*C} bias+gamma+res", lambda: L.gemm(hid, w2, x, M, C, 4 * C, bias=b2, gamma=g, residual=x, ldr=C), fl)
    timed(f"fc2 M{M} N{C} K{4*C} plain", lambda: L.gemm(hid, w2, x, M, C, 4 * C), fl)
    if os.environ.get("GCV_DEBUG", "0") == "0":
        timed(f"  cuBLAS same shape", lambda: torch.matmul(hid, w2.t(), out=x), fl)
