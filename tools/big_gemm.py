import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.bfloat16
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/reps
for (M, N, K) in ((8192, 8192, 8192), (4096, 4096, 4096), (50176, 1536, 384), (50176, 384, 1536), (12544, 3072, 768), (12544, 768, 3072)):
    a = torch.randn(M, K, device=dev).to(dt); w = torch.randn(N, K, device=dev).to(dt)
    d = torch.empty(M, N, device=dev, dtype=dt)
    ms = timed(lambda: L.gemm(a, w, d, M, N, K))
    ms_t = timed(lambda: torch.matmul(a, w.t(), out=d))
    print(f"M={M} N={N} K={K}: ours {ms:7.3f} ms {2*M*N*K/ms/1e9:7.1f} TF/s | cuBLAS {ms_t:7.3f} ms {2*M*N*K/ms_t/1e9:7.1f} TF/s")
