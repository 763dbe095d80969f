"""Launch a few representative kernels in isolation (for ncu --set full captures and quick timing).
Usage: python tools/kernel_probe.py [dwconv|fc1|fc2|all] [--batch 64]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from genconvit_b200 import lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("which", nargs="?", default="all")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
dev, dt = "cuda", torch.bfloat16
B = args.batch
torch.manual_seed(0)


def timed(name, fn, work, unit):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.reps
    print(f"{name:40s} {ms:8.3f} ms  {work / ms / 1e9:9.1f} {unit}")


if args.which in ("dwconv", "all"):
    for (H, C) in ((56, 96), (28, 192), (14, 384), (7, 768)):
        x = torch.randn(B, H, H, C, device=dev).to(dt)
        y = torch.empty_like(x)
        taps, bias = torch.randn(49, C, device=dev) / 7, torch.randn(C, device=dev)
        lw, lb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        timed(f"dwconv7_ln B{B} H{H} C{C}", lambda: L.dwconv7_ln(x, y, taps, bias, lw, lb, 1e-6, B, H, H, C),
              2.0 * x.numel() * 2 / 1e3, "TB/s")
        timed(f"  layernorm_rows only B{B} H{H} C{C}", lambda: L.layernorm_rows(y, y, lw, lb, 1e-6, B * H * H, C),
              2.0 * x.numel() * 2 / 1e3, "TB/s")
if args.which in ("dwstats", "all"):
    # the tensor-core depthwise kernel as the 16-bit forward calls it (un-normalised output + LayerNorm partial sums)
    for (H, C, mult) in ((56, 96, 2), (28, 192, 2), (14, 384, 2), (7, 768, 2), (28, 96, 1), (14, 192, 1)):
        Bn = B * mult
        x = torch.randn(Bn, H, H, C, device=dev).to(dt)
        y = torch.empty_like(x)
        stats = torch.empty(Bn * H * H, C // 32, 2, device=dev)
        taps, bias = torch.randn(49, C, device=dev) / 7, torch.randn(C, device=dev)
        timed(f"dwconv7_stats B{Bn} H{H} C{C}", lambda: L.dwconv7_stats(x, y, stats, taps, bias, Bn, H, H, C),
              2.0 * x.numel() * 2 / 1e3, "TB/s")
if args.which in ("fc1", "all"):
    for (T, C) in ((3136, 96), (784, 192), (196, 384), (49, 768)):
        M = B * T
        a = torch.randn(M, C, device=dev).to(dt)
        w = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
        bias = torch.randn(4 * C, device=dev)
        d = torch.empty(M, 4 * C, device=dev, dtype=dt)
        timed(f"fc1+gelu M{M} N{4 * C} K{C}", lambda: L.gemm(a, w, d, M, 4 * C, C, bias=bias, act=L.ACT_GELU),
              2.0 * M * 4 * C * C / 1e3, "TF/s")
if args.which in ("fc2", "all"):
    for (T, C) in ((3136, 96), (784, 192), (196, 384), (49, 768)):
        M = B * T
        a = torch.randn(M, 4 * C, device=dev).to(dt)
        w = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt)
        bias, gamma = torch.randn(C, device=dev), torch.rand(C, device=dev)
        x = torch.randn(M, C, device=dev).to(dt)
        timed(f"fc2+res M{M} N{C} K{4 * C}",
              lambda: L.gemm(a, w, x, M, C, 4 * C, bias=bias, gamma=gamma, residual=x, ldr=C), 2.0 * M * 4 * C * C / 1e3, "TF/s")
if args.which in ("fused", "all"):
    for (T, C) in ((3136, 96), (784, 192)):
        M = B * T
        y = torch.randn(M, C, device=dev).to(dt)
        x = torch.randn(M, C, device=dev).to(dt)
        w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
        w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt)
        b1, b2, g = torch.randn(4 * C, device=dev), torch.randn(C, device=dev), torch.rand(C, device=dev) * 0.1
        timed(f"mlp_fused M{M} C{C}", lambda: L.mlp_fused(y, w1, b1, w2, b2, g, x, M, C), 16.0 * M * C * C / 1e3, "TF/s")
