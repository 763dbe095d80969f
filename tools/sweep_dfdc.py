"""BASELINE.json config 4: a DFDC-shaped synthetic sweep -- 1000 videos x 15 frames of 224x224, scored end to end (pinned
host uint8 face crops -> per-video class / score on the host) through genconvit_b200.runtime.VideoScorer at several batch
sizes, bf16, on 1..8 GPUs: videos are sharded over the ranks (runtime.shard_videos), weights are replicated, every step
ends in the all-gather of the per-video results (runtime.gather_scores) and rank 0 holds all scores.  One JSON line per
batch size (rank 0).

    python tools/sweep_dfdc.py [--videos 1000] [--dtype bf16] [--batches 60,240,960]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_dfdc.py
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from genconvit_b200.runtime import VideoScorer, gather_scores, shard_videos  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=1000)
ap.add_argument("--fpv", type=int, default=15)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batches", default="60,240,960", help="frames per GPU per step")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
model = bench.build_model({"bf16": torch.bfloat16, "fp16": torch.float16}[args.dtype], dev)
lo, hi = shard_videos(args.videos, rank, world)
mine = hi - lo
for bs in [int(b) for b in args.batches.split(",")]:
    vids = bs // args.fpv
    sc = VideoScorer(model, vids * args.fpv, args.fpv)
    g = torch.Generator().manual_seed(7 + rank)
    hosts = [torch.randint(0, 256, (vids * args.fpv, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(2)]
    outs = [torch.empty((2, vids), dtype=torch.float32).pin_memory() for _ in range(2)]
    steps = -(-max(shard_videos(args.videos, r, world)[1] - shard_videos(args.videos, r, world)[0] for r in range(world)) // vids)

    def step(i):
        sc.submit(hosts[i & 1], outs[i & 1])
        if world > 1:
            gather_scores(sc.out)            # the one collective of the path: every rank ends up with all scores
    for i in range(3):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # device time, max over ranks
    dt = float(ms.item()) / 1e3
    n_vid = min(args.videos, steps * vids * world)
    if rank == 0:
        print(json.dumps({"workload": f"{args.videos} videos x {args.fpv} frames over {world} GPU(s), {vids * args.fpv} frames per GPU per step "
                                      f"({steps} steps), {args.dtype}, end to end from pinned host uint8 crops, all-gather of per-video scores",
                          "n_gpus": world, "frames_per_s": n_vid * args.fpv / dt, "videos_per_s": n_vid / dt,
                          "ms_per_step": 1e3 * dt / steps, "videos_scored": n_vid,
                          "fake_fraction_rank0": float(outs[(steps - 1) & 1][0].mean())}), flush=True)
    del sc
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
