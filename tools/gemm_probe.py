"""Time the stage-2 / stage-3 MLP GEMM shapes of the bs256 step in isolation (fc1 with the folded LayerNorm + GELU, fc2 with
layer scale + residual) -- the shapes DESIGN.md section 4 quotes.  With a -DGCV_GEMM_WHATIF build, GCV_DEBUG=1/2/6 gives the
mainloop-only / no-store / TMEM-drain-only what-if timings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from genconvit_b200 import lib as L  # noqa: E402

dev, dt = torch.device("cuda:0"), torch.float16


def timed(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for (M, N, K, kind) in ((100352, 1536, 384, "ln"), (100352, 384, 1536, "res"), (25088, 3072, 768, "ln"), (25088, 768, 3072, "res"),
                        (401408, 192, 384, "plain")):
    a = torch.randn(M, K, device=dev).to(dt)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=dev)
    d = torch.randn(M, N, device=dev).to(dt)
    if kind == "ln":
        st = torch.rand(M, K // 32, 2, device=dev) + 1.0
        f = lambda: L.gemm(a, w, d, M, N, K, bias=bias, act=L.ACT_GELU, ln_stats=st, ln_colsum=bias, ln_eps=1e-6)  # noqa: E731
    elif kind == "res":
        g = torch.rand(N, device=dev)
        f = lambda: L.gemm(a, w, d, M, N, K, bias=bias, gamma=g, residual=d, ldr=N)  # noqa: E731
    else:
        f = lambda: L.gemm(a, w, d, M, N, K, bias=bias)  # noqa: E731
    ms = timed(f)
    print(f"M{M} N{N} K{K} {kind:5s} {ms:.4f} ms {2.0 * M * N * K / ms / 1e9:7.1f} TF/s")
