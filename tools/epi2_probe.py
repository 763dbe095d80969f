"""Epilogue cost split on the stage 2 fc1 / fc2 shapes: GCV_DEBUG=0 full, 2 no global stores, 6 TMEM drain only, 1 none."""
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
    print(f"dbg{os.environ.get('GCV_DEBUG','0')} {name:44s} {ms:8.4f} ms {flops / ms / 1e9:8.1f} TF/s", flush=True)
M, C = 100352, 384
a = torch.randn(M, C, device=dev).to(dt); w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
b1 = torch.randn(4 * C, device=dev); hid = torch.empty(M, 4 * C, device=dev, dtype=dt)
w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt); b2 = torch.randn(C, device=dev)
g = torch.rand(C, device=dev); x = torch.randn(M, C, device=dev).to(dt)
stats = torch.rand(M, C // 32, 2, device=dev); cs = torch.randn(4 * C, device=dev)
fl = 2.0 * M * 4 * C * C
timed("fc1 bias+gelu", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU), fl)
timed("fc1 bias+gelu+ln", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU, ln_stats=stats, ln_colsum=cs, ln_eps=1e-6), fl)
timed("fc1 plain", lambda: L.gemm(a, w1, hid, M, 4 * C, C), fl)
timed("fc2 bias+gamma+res", lambda: L.gemm(hid, w2, x, M, C, 4 * C, bias=b2, gamma=g, residual=x, ldr=C), fl)
timed("fc2 plain", lambda: L.gemm(hid, w2, x, M, C, 4 * C), fl)
timed("fc2 plain bn256", lambda: L.gemm(hid, w2, x, M, C, 4 * C, backend=1256), fl)
timed("fc2 plain bn128", lambda: L.gemm(hid, w2, x, M, C, 4 * C, backend=1128), fl)
timed("fc2 bias+gamma+res bn192", lambda: L.gemm(hid, w2, x, M, C, 4 * C, bias=b2, gamma=g, residual=x, ldr=C, backend=1192), fl)
timed("fc2 bias+gamma+res bn128", lambda: L.gemm(hid, w2, x, M, C, 4 * C, bias=b2, gamma=g, residual=x, ldr=C, backend=1128), fl)
timed("fc2 bias+gamma+res bn256", lambda: L.gemm(hid, w2, x, M, C, 4 * C, bias=b2, gamma=g, residual=x, ldr=C, backend=1256), fl)
timed("fc1 bias+gelu+ln bn192", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU, ln_stats=stats, ln_colsum=cs, ln_eps=1e-6, backend=1192), fl)
timed("fc1 bias+gelu+ln bn128", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU, ln_stats=stats, ln_colsum=cs, ln_eps=1e-6, backend=1128), fl)
stats1 = torch.rand(M, 1, 2, device=dev)
timed("fc1 bias+gelu+ln (1 stats chunk)", lambda: L.gemm(a, w1, hid, M, 4 * C, C, bias=b1, act=L.ACT_GELU, ln_stats=stats1, ln_colsum=cs, ln_eps=1e-6), fl)
