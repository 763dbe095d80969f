"""Per-launch CUDA-event timing of one eager GenConViT step (bs256): which shapes are slow.
Usage (GPU box): python tools/profile_step.py [--batch 256] [--dtype bf16] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from genconvit_b200 import lib  # noqa: E402
from genconvit_b200.runtime import VideoScorer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--tags-out", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
import model.genconvit as _gm  # noqa: E402
_gm.set_two_streams(False)      # per-launch durations: no overlap between the two networks
model = bench.build_model({"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype], dev)
sc = VideoScorer(model, args.batch, 16, use_graph=False)
sc.x_static.normal_().clamp_(-2.1, 2.6)
with torch.no_grad():
    for _ in range(args.steps):
        lib.profile = []
        sc._step()
        torch.cuda.synchronize()
        prof, lib.profile = lib.profile, None
if args.tags_out:
    # launch order of one step with each launch's shape tag and algorithmic work (FLOPs for the GEMM kernels, bytes for
    # the others): zipped with an `ncu -k regex:gcv` capture of the same program by tools/ncu_step_traffic.py
    import json
    json.dump([[name, tag, work] for name, work, _s, _e, tag in prof], open(args.tags_out, "w"))
rows = {}
for name, work, s, e, tag in prof:
    r = rows.setdefault((name, tag), [0, 0.0, 0.0])
    r[0] += 1
    r[1] += s.elapsed_time(e)
    r[2] += work
tot = sum(r[1] for r in rows.values())
print(f"total {tot:.2f} ms over {len(prof)} launches")
for (name, tag), (cnt, ms, work) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    tensor = name.startswith("gemm") or name in ("mlp_fused", "conv3x3_tc")     # these report FLOPs as their work
    rate = work / (ms / 1e3) / (1e12 if tensor else 1e9)
    unit = "TF/s" if tensor else "GB/s"
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt:<3d} {ms / cnt:7.3f} ms/launch {rate:8.1f} {unit}  {name:18s} {tag}")
