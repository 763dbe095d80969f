"""Epilogue cost experiments on the stage-0 fc1 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.bfloat16
def timed(name, fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    print(f"{name:50s} {s.elapsed_time(e)/reps:8.3f} ms")
M, C = 802816, 96
a = torch.randn(M, C, device=dev).to(dt)
w = (torch.randn(4*C, C, device=dev)/C**0.5).to(dt)
bias = torch.randn(4*C, device=dev)
d = torch.empty(M, 4*C, device=dev, dtype=dt)
timed("fc1 K96 N384 bias+gelu", lambda: L.gemm(a, w, d, M, 4*C, C, bias=bias, act=L.ACT_GELU))
timed("fc1 K96 N384 bias+relu", lambda: L.gemm(a, w, d, M, 4*C, C, bias=bias, act=L.ACT_RELU))
timed("fc1 K96 N384 bias", lambda: L.gemm(a, w, d, M, 4*C, C, bias=bias))
timed("fc1 K96 N384 plain", lambda: L.gemm(a, w, d, M, 4*C, C))
for bn in (64, 128, 192, 256):
    timed(f"fc1 K96 N384 plain block_n={bn}", lambda: L.gemm(a, w, d, M, 4*C, C, backend=1000+bn))
# tiny N: epilogue nearly free -> mainloop/launch floor
d2 = torch.empty(M, 32, device=dev, dtype=dt)
timed("K96 N32 plain (mainloop floor)", lambda: L.gemm(a, w, d2, M, 32, C))
a2 = torch.randn(M, 384, device=dev).to(dt); w2 = (torch.randn(96, 384, device=dev)/20).to(dt)
x = torch.randn(M, 96, device=dev).to(dt); g = torch.rand(96, device=dev); b2 = torch.randn(96, device=dev)
timed("fc2 K384 N96 bias+gamma+res", lambda: L.gemm(a2, w2, x, M, 96, 384, bias=b2, gamma=g, residual=x, ldr=96))
timed("fc2 K384 N96 plain", lambda: L.gemm(a2, w2, x, M, 96, 384))
# pure copy reference for HBM
src = torch.empty(M*384, device=dev, dtype=dt); dst = torch.empty_like(src)
timed("torch copy 616MB (r+w)", lambda: dst.copy_(src))
