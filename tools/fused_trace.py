import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.bfloat16
for (T, C) in ((3136, 96), (784, 192)):
    M = 256 * T
    y = torch.randn(M, C, device=dev).to(dt); x = torch.randn(M, C, device=dev).to(dt)
    w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(dt); w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(dt)
    b1, b2, g = torch.randn(4 * C, device=dev), torch.randn(C, device=dev), torch.rand(C, device=dev) * 0.1
    for _ in range(3):
        L.mlp_fused(y, w1, b1, w2, b2, g, x, M, C)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    n = L.load().gcv_debug_fused_trace(buf)
    tiles = (M // 128 + 147) // 148
    print(f"C={C} tiles/CTA~{tiles} trace n={n}")
    names = {0: ("producer", ["x_empty", "ring_empty"]), 8: ("mma", ["x_full", "s_empty", "ring_full(fc1)", "h_full", "o_empty", "ring_full(fc2)"]),
             16: ("epi warp2 (grp0)", ["s_full", "h_empty", "o_full"]), 24: ("epi warp10 (grp1)", ["s_full", "h_empty", "o_full"])}
    for base, (role, labels) in names.items():
        tot = buf[base + 7]
        print(f"  {role:18s} total {tot:9d} cyc ({tot/tiles:7.0f}/tile): " + ", ".join(f"{l}={buf[base+i]/tiles:6.0f}" for i, l in enumerate(labels)))
