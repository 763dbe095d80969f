"""BASELINE.json config 4 on one GPU: a DFDC-shaped synthetic sweep -- 1000 videos x 15 frames of 224x224, scored end to end
(pinned host frames -> per-video class / score on the host) through genconvit_b200.runtime.VideoScorer at several batch
sizes, bf16.  Prints one JSON line per batch size.  Usage: python tools/sweep_dfdc.py [--videos 1000] [--dtype bf16]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from genconvit_b200.runtime import VideoScorer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=1000)
ap.add_argument("--fpv", type=int, default=15)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batches", default="60,240,960")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build_model({"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype], dev)
for bs in [int(b) for b in args.batches.split(",")]:
    vids = bs // args.fpv
    sc = VideoScorer(model, vids * args.fpv, args.fpv)
    hosts = [torch.randn(vids * args.fpv, 3, 224, 224).clamp_(-2.1179, 2.64).pin_memory() for _ in range(2)]
    outs = [torch.empty((2, vids), dtype=torch.float32).pin_memory() for _ in range(2)]
    steps = -(-args.videos // vids)
    for i in range(3):
        sc.submit(hosts[i & 1], outs[i & 1])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        sc.submit(hosts[i & 1], outs[i & 1])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = steps * vids * args.fpv
    print(json.dumps({"workload": f"{steps * vids} videos x {args.fpv} frames, batch {vids * args.fpv} frames, {args.dtype}, end to end",
                      "frames_per_s": n / dt, "videos_per_s": steps * vids / dt, "ms_per_batch": 1e3 * dt / steps,
                      "fake_fraction": float(outs[(steps - 1) & 1][0].mean())}), flush=True)
    del sc
    torch.cuda.empty_cache()
