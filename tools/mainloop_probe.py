"""Mainloop experiments (GCV_DEBUG / GCV_GEMM_STAGES via env) on the stage 2/3 shapes and 8192^3."""
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
    print(f"{name:44s} {ms:8.4f} ms {flops / ms / 1e9:8.1f} TF/s", flush=True)
print("GCV_DEBUG =", os.environ.get("GCV_DEBUG", "0"), "GCV_GEMM_STAGES =", os.environ.get("GCV_GEMM_STAGES", "-"))
for (M, N, K) in ((100352, 1536, 384), (100352, 384, 1536), (25088, 3072, 768), (25088, 768, 3072), (8192, 8192, 8192)):
    a = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    d = torch.empty(M, N, device=dev, dtype=dt)
    timed(f"M{M} N{N} K{K} plain", lambda: L.gemm(a, w, d, M, N, K), 2.0 * M * N * K)
