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
if args.which in ("fusedln",):
    dtl = torch.float16
    for (T, C) in ((3136, 96), (784, 192)):
        M = B * T
        y = torch.randn(M, C, device=dev).to(dtl)
        x = torch.randn(M, C, device=dev).to(dtl)
        w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dtl)
        w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dtl)
        st = torch.rand(M, C // 32, 2, device=dev) + 1.0
        b1, cs, b2, g = torch.randn(4 * C, device=dev), torch.randn(4 * C, device=dev), torch.randn(C, device=dev), torch.rand(C, device=dev) * 0.1
        timed(f"mlp_fused_ln M{M} C{C}", lambda: L.mlp_fused_ln(y, st, 1e-6, w1, b1, cs, w2, b2, g, x, M, C), 16.0 * M * C * C / 1e3, "TF/s")
if args.which in ("fused", "all"):
    for (T, C) in ((3136, 96), (784, 192)):
        M = B * T
        y = torch.randn(M, C, device=dev).to(dt)
        x = torch.randn(M, C, device=dev).to(dt)
        w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
        w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt)
        b1, b2, g = torch.randn(4 * C, device=dev), torch.randn(C, device=dev), torch.rand(C, device=dev) * 0.1
        timed(f"mlp_fused M{M} C{C}", lambda: L.mlp_fused(y, w1, b1, w2, b2, g, x, M, C), 16.0 * M * C * C / 1e3, "TF/s")

if args.which == "top":
    # One launch of every shape that matters in the bs256 step (fp16), each on its own small set of buffers: the input of
    # the per-shape `ncu --set full` capture (profiles/r2_ncu_full_top_kernels.json).  Prints the launch order with each
    # launch's algorithmic bytes so that the capture can be labelled.
    import json
    dt = torch.float16
    order = []

    def note(name, shape, algo_bytes, flops=0.0):
        order.append({"kernel": name, "shape": shape, "algorithmic_bytes": algo_bytes, "flops": flops})

    def gemm(M, N, K, act=L.ACT_NONE, res=False, ln=False):
        a = torch.randn(M, K, device=dev).to(dt)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
        bias = torch.randn(N, device=dev)
        d = torch.randn(M, N, device=dev).to(dt)
        kw = {}
        if res:
            kw = dict(gamma=torch.rand(N, device=dev), residual=d, ldr=N)
        if ln:
            st = torch.rand(M, K // 32, 2, device=dev) + 1.0       # per-chunk partial sums: the statistics warps' path
            kw = dict(ln_stats=st, ln_colsum=torch.randn(N, device=dev), ln_eps=1e-6)
        L.gemm(a, w, d, M, N, K, bias=bias, act=act, **kw)
        note("gemm_tcgen05", f"M{M} N{N} K{K}" + ("+res" if res else "") + ("+ln" if ln else ""),
             2.0 * (M * K + N * K + M * N) + (2.0 * M * N if res else 0.0), 2.0 * M * N * K)

    def dw(Bn, H, C):
        x = torch.randn(Bn, H, H, C, device=dev).to(dt)
        y = torch.empty_like(x)
        st = torch.empty(Bn * H * H, C // 32, 2, device=dev)
        L.dwconv7_stats(x, y, st, torch.randn(49, C, device=dev) / 7, torch.randn(C, device=dev), Bn, H, H, C)
        note("dwconv7_mma", f"B{Bn} H{H} C{C}", 2.0 * x.numel() * 2 + st.numel() * 4)

    def fused(M, C):
        y, x = torch.randn(M, C, device=dev).to(dt), torch.randn(M, C, device=dev).to(dt)
        w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt)
        w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt)
        st = torch.rand(M, C // 32, 2, device=dev) + 1.0
        L.mlp_fused_ln(y, st, 1e-6, w1, torch.randn(4 * C, device=dev), torch.randn(4 * C, device=dev), w2,
                       torch.randn(C, device=dev), torch.rand(C, device=dev) * 0.1, x, M, C)
        note("mlp_fused", f"M{M} C{C} +ln", 3 * 2.0 * M * C + st.numel() * 4, 16.0 * M * C * C)

    dw(512, 56, 96); dw(512, 28, 192); dw(512, 14, 384); dw(512, 7, 768)
    fused(512 * 3136, 96); fused(512 * 784, 192)
    gemm(100352, 1536, 384, act=L.ACT_GELU, ln=True); gemm(100352, 384, 1536, res=True)
    gemm(25088, 3072, 768, act=L.ACT_GELU, ln=True); gemm(25088, 768, 3072, res=True)
    gemm(256, 12544, 25088)
    gemm(401408, 192, 384); gemm(100352, 384, 768)
    # helper kernels
    x4 = torch.randn(512, 56, 56, 96, device=dev).to(dt)
    a4 = torch.empty(512 * 28 * 28, 384, device=dev, dtype=dt)
    L.ln_patchify2(x4, a4, torch.ones(96, device=dev), torch.zeros(96, device=dev), 1e-6, 512, 56, 56, 96)
    note("ln_patchify2", "B512 H56 C96", 2.0 * x4.numel() * 2)
    fr = torch.randn(256, 3, 224, 224, device=dev)
    tok = torch.empty(256 * 56 * 56, 96, device=dev, dtype=dt)
    L.stem_fused(fr, tok, (torch.randn(96, 48, device=dev) / 7).to(dt), torch.zeros(96, device=dev), torch.ones(96, device=dev),
                 torch.zeros(96, device=dev), 1e-6, 256, 224, 224, True)
    note("stem_fused", "B256 H224 nchw", fr.numel() * 4.0 + tok.numel() * 2.0)
    e1 = torch.empty(256 * 112 * 112, 16, device=dev, dtype=dt)
    L.conv3x3_first(fr, e1, torch.randn(16, 3, 3, 3, device=dev) / 5, torch.zeros(16, device=dev), 1, L.ACT_RELU, True, 256, 224, 224)
    note("conv3x3_first", "B256 H224 s1 pool", fr.numel() * 4.0 + e1.numel() * 2.0)
    e2 = torch.empty(256 * 56 * 56, 32, device=dev, dtype=dt)
    L.conv3x3_c16(e1, e2, (torch.randn(32, 144, device=dev) / 12).to(dt), torch.zeros(32, device=dev), 1, L.ACT_RELU, True, 256, 112, 112)
    note("conv3x3_c16", "B256 H112 s1 pool", e1.numel() * 2.0 + e2.numel() * 2.0)
    e3 = torch.empty(256 * 28 * 28, 64, device=dev, dtype=dt)
    L.conv3x3_c32(e2, e3, (torch.randn(64, 288, device=dev) / 17).to(dt), torch.zeros(64, device=dev), 1, L.ACT_RELU, True, 256, 56, 56)
    note("conv3x3_c32", "B256 H56 s1 pool", e2.numel() * 2.0 + e3.numel() * 2.0)
    e4 = torch.empty(256 * 14 * 14, 128, device=dev, dtype=dt)
    L.conv3x3_tc(e3, e4, (torch.randn(128, 576, device=dev) / 24).to(dt), torch.zeros(128, device=dev), 1, L.ACT_RELU, True,
                 256, 28, 28, 64, 128)
    note("conv3x3_tc", "B256 H28 C64 s1 pool", e3.numel() * 2.0 + e4.numel() * 2.0 + 128 * 576 * 2.0, 2.0 * 256 * 784 * 128 * 576)
    e5 = torch.empty(256 * 7 * 7, 256, device=dev, dtype=dt)
    L.conv3x3_tc(e4, e5, (torch.randn(256, 1152, device=dev) / 34).to(dt), torch.zeros(256, device=dev), 1, L.ACT_RELU, True,
                 256, 14, 14, 128, 256)
    note("conv3x3_tc", "B256 H14 C128 s1 pool", e4.numel() * 2.0 + e5.numel() * 2.0 + 256 * 1152 * 2.0, 2.0 * 256 * 196 * 256 * 1152)
    d32 = torch.randn(256 * 56 * 56, 32, device=dev).to(dt)
    img = torch.empty(256 * 224 * 224, 3, device=dev, dtype=dt)
    L.convt2x2_mma(d32, img, (torch.randn(64, 32, device=dev) / 6).to(dt), torch.zeros(64, device=dev), L.ACT_RELU, 256, 56, 56,
                   32, w2=(torch.randn(12, 16, device=dev) / 4).to(dt), b2=torch.zeros(12, device=dev))
    note("convt2x2_mma", "B256 H56 C32 +tail", d32.numel() * 2.0 + img.numel() * 2.0)
    u8 = torch.randint(0, 256, (256, 224, 224, 3), device=dev, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    L.stem_fused_u8(u8, tok, (torch.randn(96, 48, device=dev) / 7).to(dt), torch.zeros(96, device=dev), torch.ones(96, device=dev),
                    torch.zeros(96, device=dev), 1e-6, 256, 224, 224, mean, std)
    note("stem_fused", "B256 H224 u8", u8.numel() * 1.0 + tok.numel() * 2.0)
    L.conv3x3_first_u8(u8, e1, torch.randn(16, 3, 3, 3, device=dev) / 5, torch.zeros(16, device=dev), 1, L.ACT_RELU, True, 256,
                       224, 224, mean, std)
    note("conv3x3_first", "B256 H224 s1 pool u8", u8.numel() * 1.0 + e1.numel() * 2.0)
    torch.cuda.synchronize()
    json.dump(order, open("gpurun_out/r2_top_order.json", "w"))
    print(len(order), "launches")
