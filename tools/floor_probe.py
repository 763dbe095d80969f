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
M = 802816
for N in (32, 128):
    for K in (64, 96, 128, 256, 384, 768, 1536):
        a = torch.randn(M, K, device=dev).to(dt); w = torch.randn(N, K, device=dev).to(dt)
        d = torch.empty(M, N, device=dev, dtype=dt)
        ms = timed(lambda: L.gemm(a, w, d, M, N, K, backend=1000 + N))
        tiles = (M // 128) / 148
        print(f"N={N:4d} K={K:5d}: {ms:7.3f} ms  {ms*1e-3*1.9e9/tiles:8.0f} cyc/tile  A-read {M*K*2/ms/1e9:7.1f} GB/s  {2*M*N*K/ms/1e9:7.1f} TF/s")
        del a, w, d
