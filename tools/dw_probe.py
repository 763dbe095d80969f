"""dwconv7_stats (tensor-core depthwise 7x7) on the bench shapes, in isolation (timing or an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.float16
reps = int(os.environ.get("REPS", "10"))
for (B, H, C) in ((512, 56, 96), (512, 28, 192), (512, 14, 384), (512, 7, 768), (256, 28, 96), (256, 7, 384)):
    x = torch.randn(B, H, H, C, device=dev).to(dt); y = torch.empty_like(x)
    stats = torch.empty(B * H * H, C // 32, 2, device=dev)
    taps, bias = torch.randn(49, C, device=dev) / 7, torch.randn(C, device=dev)
    fn = lambda: L.dwconv7_stats(x, y, stats, taps, bias, B, H, H, C)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"dwconv7_stats B{B} H{H} C{C}: {ms:.4f} ms  {2 * x.numel() * 2 / ms / 1e9:.2f} TB/s (in+out)", flush=True)
