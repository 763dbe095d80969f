"""Small-K / small-N GEMMs of the stem and the transposed-conv decoders under the GCV_DEBUG knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genconvit_b200 import lib as L
dev, dt = "cuda", torch.float16
def timed(name, fn, byts, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / reps
    print(f"dbg{os.environ.get('GCV_DEBUG','0')} {name:40s} {ms:8.4f} ms {byts / ms / 1e9:8.2f} TB/s", flush=True)
for (M, N, K) in ((1605632, 96, 48), (802816, 64, 32), (802816, 64, 288), (3211264, 32, 144)):
    a = torch.randn(M, K, device=dev).to(dt); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
    b = torch.randn(N, device=dev); d = torch.empty(M, N, device=dev, dtype=dt)
    timed(f"M{M} N{N} K{K} bias", lambda: L.gemm(a, w, d, M, N, K, bias=b), 2.0 * M * (N + K))
