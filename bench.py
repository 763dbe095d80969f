#!/usr/bin/env python
"""Benchmark of the GenConViT batched frame-inference forward (ED + VAE -> pred_vid scores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic frames per GPU (256 frames of
224x224, i.e. 16 videos x 16 frames): ED + VAE forward, fused per-video scoring and, for N > 1,
the all-gather of per-video scores.  Prints ONE JSON line (rank 0).

  value     frames/s, whole job, inputs already resident in HBM (CUDA-graph replay per step)
  e2e       the same metric through genconvit_b200.runtime.VideoScorer with pinned HOST frames:
            H2D of every batch and D2H of its scores are inside the timed region
  roofline  the dominant kernel (the tcgen05 GEMM): sum of algorithmic FLOPs / sum of CUDA-event
            durations over every launch of one step, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a torch-fp32 port of the reference forward) on the host cores

--impl reference times the reference's CPU implementation of the path (the oracle port: the
reference's own files need timm==0.6.5, which is not installable offline) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "GenConViT ED+VAE frames/sec @224^2, bs256 per GPU"
GFLOP_PER_FRAME_CONTRACTION = 29.49      # BASELINE.md section 2 (2xMAC, mu once, Swin excluded)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


_RESULT_FD = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is not None:
        os.write(_RESULT_FD, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.thread = [], None, None
        try:
            uuid = str(torch.cuda.get_device_properties(device).uuid)
            self.sel = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        except Exception:
            self.sel = str(device)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.sel, f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, power = [], 0.0, set(), 0.0
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                power = max(power, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": power or None}


def build_model(dtype, device):
    """Random-init GenConViT (ED + VAE) with the reference architecture; layer-scale, norms and
    biases are randomised so no branch of the network degenerates to identity."""
    from model.config import load_config
    from model.genconvit import GenConViT
    from model.genconvit_ed import GenConViTED
    from model.genconvit_vae import GenConViTVAE
    cfg = load_config()
    torch.manual_seed(0)
    with torch.device(device):
        ed, vae = GenConViTED(cfg), GenConViTVAE(cfg)
    with torch.no_grad():
        for m in (ed, vae):
            for name, p in m.named_parameters():
                if name.endswith("gamma"):
                    p.uniform_(0.0, 0.5)
                elif name.endswith("bias"):
                    p.normal_(0.0, 0.02)
    model = GenConViT.from_modules(ed, vae).to(device).eval()
    model.set_compute_dtype(dtype)
    return model


def cpu_oracle_fps(n_frames, threads):
    """Time the CPU oracle (full GenConViT forward + pred_vid) on ``n_frames`` frames."""
    from oracle import nets
    from oracle.weights import make_state_dict, synthetic_eps, synthetic_frames
    torch.set_num_threads(threads)
    sd_ed, sd_vae = make_state_dict("ed", 0), make_state_dict("vae", 0, skip_var=True)
    x, eps = synthetic_frames(n_frames, 0), synthetic_eps(n_frames, 7)

    def run(k):
        with torch.no_grad():
            return nets.pred_vid(nets.genconvit_forward(sd_ed, sd_vae, x[:k], eps[:k]))
    run(min(16, n_frames))                                  # warm-up (thread pool, oneDNN primitives, allocator)
    best = float("inf")
    for _ in range(2):                                      # best of two: the first full-size pass still warms caches
        t0 = time.perf_counter()
        run(n_frames)
        best = min(best, time.perf_counter() - t0)
    return n_frames / best, (sd_ed, sd_vae, x, eps, run)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_reference_model(device, half):
    """The UNMODIFIED reference files (staged by __graft_entry__.build() into baseline/_ref/ -- git-ignored, shipped to
    the GPU box) through their own classes: model.genconvit.GenConViT wrapping GenConViTED / GenConViTVAE, scored by the
    reference's own pred_func.pred_vid.  timm==0.6.5 is not installable offline: oracle/timm_standin.py supplies
    ``timm.create_model`` (module shells with timm's parameter names running the same torch operators).  Random-init
    weights (the modules' own constructors).  Returns (model, pred_vid) or None when nothing is staged."""
    if not os.path.exists(os.path.join(REF_DIR, "model", "genconvit.py")):
        return None
    from oracle import timm_standin
    sys.path.insert(0, REF_DIR)                 # the reference's ``model`` / ``dataset`` packages shadow the repo's
    for name in [m for m in sys.modules if m == "model" or m.startswith("model.")]:
        del sys.modules[name]
    timm_standin.install()
    cwd = os.getcwd()
    os.chdir(REF_DIR)                           # model/config.py reads model/config.yaml relative to the package
    try:
        import model.pred_func as ref_pred
        from model.config import load_config
        from model.genconvit import GenConViT
        from model.genconvit_ed import GenConViTED
        from model.genconvit_vae import GenConViTVAE
        assert os.path.abspath(ref_pred.__file__).startswith(REF_DIR), ref_pred.__file__
        cfg = load_config()
        torch.manual_seed(0)
        ref = GenConViT.__new__(GenConViT)      # skip only the weight/*.pth loading of __init__ (no checkpoints offline)
        torch.nn.Module.__init__(ref)
        ref.net, ref.fp16 = "genconvit", bool(half)
        ref.model_ed = GenConViTED(cfg, pretrained=False).eval()
        ref.model_vae = GenConViTVAE(cfg, pretrained=False).eval()
    finally:
        os.chdir(cwd)
    ref.eval().to(device)
    if half:
        ref.half()
    return ref, ref_pred.pred_vid


def run_reference(args, rank, world):
    """--impl reference: the reference's own implementation of the path.  Default: CPU, fp32, all host cores (the
    reference arm of the contract).  --ref-device cuda --ref-half: the same modules run eagerly on the GPU through
    cuDNN / cuBLAS in fp16, channels-last -- the 'existing Blackwell path' our arm reports as gpu_eager_baseline."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample, fpv = args.ref_frames, args.fpv
    on_gpu = args.ref_device == "cuda"
    dev = torch.device("cuda", 0) if on_gpu else torch.device("cpu")
    loaded = load_reference_model(dev, args.ref_half) if not args.ref_port else None
    g = torch.Generator().manual_seed(0)
    x = torch.randn(sample, 3, 224, 224, generator=g).clamp_(-2.1179, 2.64)
    if loaded is not None:
        ref, ref_pred_vid = loaded
        kind = "reference+timm-standin"
        x = x.to(dev)
        if args.ref_half:
            x = x.half()
        if on_gpu:
            x = x.contiguous(memory_format=torch.channels_last)

        def run():
            out = None
            with torch.no_grad():
                for v in range(0, sample, fpv):             # the reference scores one video per call (prediction.py)
                    out = ref_pred_vid(x[v:v + fpv], ref)
            return out
        if args.ref_batched:                                # one forward over the whole sample, then per-video scoring
            from oracle import nets

            def run():                                      # noqa: F811
                with torch.no_grad():
                    rows = ref(x).float()
                    n = x.shape[0]
                    return [nets.pred_vid(torch.cat((rows[v:v + fpv], rows[n + v:n + v + fpv]))) for v in range(0, n, fpv)][-1]
    else:
        if on_gpu:
            emit({"impl": "reference", "unavailable": "baseline/_ref is not staged: the GPU-eager baseline needs the reference files"})
            return
        kind = "port"
        from oracle import nets
        from oracle.weights import make_state_dict, synthetic_eps
        sd_ed, sd_vae = make_state_dict("ed", 0), make_state_dict("vae", 0, skip_var=True)
        eps = synthetic_eps(sample, 7)

        def run():
            with torch.no_grad():
                return nets.pred_vid(nets.genconvit_forward(sd_ed, sd_vae, x, eps))

    def sync():
        if on_gpu:
            torch.cuda.synchronize()
    for _ in range(args.warmup):
        run()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    sync()
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    where = (f"GPU eager ({torch.cuda.get_device_name(0)}), {'fp16' if args.ref_half else 'fp32'}, channels-last, cuDNN/cuBLAS"
             if on_gpu else f"torch CPU fp32, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": ("f16" if args.ref_half else "f32"), "data": "synthetic",
        "config": {"workload": "full GenConViT (ED+VAE) forward + pred_vid scoring, 224x224, random-init",
                   "frames_per_step": sample, "frames_per_video": fpv,
                   "note": "bounded sample of the bs256 workload" + ("" if on_gpu else " on host cores")
                           + ("; one forward per video like prediction.py" if loaded is not None and not args.ref_batched else "")},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": (0 if on_gpu else threads), "kind": kind,
                         "sample": f"{sample} frames per step x {args.steps} steps after {args.warmup} warm-up steps, {where}; "
                                   + ("the unmodified reference files (baseline/_ref) with oracle/timm_standin.py as timm "
                                      "(timm==0.6.5 is not installable offline); as executed: mu three times + var"
                                      if kind != "port" else "oracle port of the reference forward (mu once)")},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def gpu_eager_baseline(args):
    """Our arm's side measurement: the reference modules run eagerly on this GPU (fp16, channels-last) in a child
    process (the reference's ``model`` package cannot share a process with the drop-in ``model`` package)."""
    if not os.path.exists(os.path.join(REF_DIR, "model", "genconvit.py")):
        return {"unavailable": "baseline/_ref not staged"}
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--ref-device", "cuda", "--ref-half",
           "--ref-frames", str(args.batch), "--ref-batched", "--steps", "5", "--warmup", "3", "--fpv", str(args.fpv)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, WORLD_SIZE="1", RANK="0"))
        line = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as exc:                                  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    if "value" not in line:
        return line
    return {"value": line["value"], "unit": "frames/s", "ms_per_step": line["ms_per_step"], "dtype": "f16",
            "frames_per_step": args.batch,
            "what": "reference modules (baseline/_ref + timm stand-in) eager on this GPU: .half(), channels-last, "
                    "cuDNN / cuBLAS, one batched forward + per-video scoring; mu three times + var as the reference executes"}


def check_against_oracle(args, model, device, rank, world):
    """--check K: every rank scores K frames of its own shard (fixed eps) through the product path and through the CPU
    oracle on the same weights, and the verdicts are gathered: max |dlogit|, per-frame decisions, per-video tuples."""
    import torch.distributed as dist
    from oracle import nets
    from model import pred_func
    k, fpv = min(args.check, args.batch), args.fpv
    k -= k % fpv
    g = torch.Generator().manual_seed(4242 + rank)
    x = torch.randn(k, 3, 224, 224, generator=g).clamp_(-2.1179, 2.64)
    eps = torch.randn(k, 12544, generator=g)
    sd_ed = {n_: t.detach().float().cpu() for n_, t in model.model_ed.state_dict().items()}
    sd_vae = {n_: t.detach().float().cpu() for n_, t in model.model_vae.state_dict().items()}
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)))
    with torch.no_grad():
        want = torch.cat([nets.genconvit_forward(sd_ed, sd_vae, x[i:i + 16], eps[i:i + 16]).view(2, -1, 2)
                          for i in range(0, k, 16)], dim=1).reshape(-1, 2)
        model.model_vae.set_epsilon(eps.to(device))
        try:
            got = model(x.to(device)).float().cpu()
            cls, val = pred_func.pred_videos(x.to(device), model, fpv)
        finally:
            model.model_vae.set_epsilon(None)
    err = (got - want).abs().max().item()
    margin = (want[:, 0] - want[:, 1]).abs()
    flips = [(int(i), float(margin[i])) for i in (got.argmax(1) != want.argmax(1)).nonzero().flatten()]
    vid_err = 0.0
    for v in range(k // fpv):
        c_want, v_want = nets.pred_vid(torch.cat((want[v * fpv:(v + 1) * fpv], want[k + v * fpv:k + (v + 1) * fpv])))
        vid_err = max(vid_err, abs(float(val[v]) - v_want))
    ok = err <= (2e-2 if args.dtype != "fp32" else 1e-4) and all(m < 2 * err for _, m in flips) and vid_err <= 5e-3
    mine = torch.tensor([err, vid_err, float(len(flips)), float(ok)], device=device)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    allv = torch.stack(allv).cpu()
    return {"frames_per_rank": k, "max_abs_dlogit": float(allv[:, 0].max()), "max_abs_dscore": float(allv[:, 1].max()),
            "near_tie_flips": int(allv[:, 2].sum()), "ranks_ok": int(allv[:, 3].sum()), "ranks": world,
            "what": "product path vs CPU oracle on the same weights / frames / eps, every rank on its own shard"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from genconvit_b200 import lib
    from genconvit_b200.runtime import VideoScorer, gather_scores

    lib.load()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    n, fpv = args.batch, args.fpv
    model = build_model(dtype, device)
    scorer = VideoScorer(model, n, fpv, eps=None, use_graph=not args.no_graph)
    g = torch.Generator(device=device).manual_seed(100 + rank)
    scorer.x_static.normal_(generator=g).clamp_(-2.1179, 2.64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step_resident():
        scorer.run_resident()
        if world > 1:
            gather_scores(scorer.out)

    clocks = ClockSampler(device)
    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    barrier()
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms / 1e3)

    # ---------------- end to end: pinned host frames -> scores on the host ----------------
    # raw uint8 NHWC face crops, the input of the reference's preprocess_frame (model/pred_func.py:95-108): 1 byte per
    # value over PCIe into one of two graph input buffers, normalised inside the first kernels of both networks
    # (--e2e-fp32: pre-processed fp32 NCHW frames)
    if args.e2e_fp32:
        hosts = [torch.randn(n, 3, 224, 224).clamp_(-2.1179, 2.64).pin_memory() for _ in range(2)]
    else:
        hosts = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    outs = [torch.empty((2, scorer.n_videos), dtype=torch.float32).pin_memory() for _ in range(2)]

    def step_e2e(i):
        scorer.submit(hosts[i & 1], outs[i & 1])
        if world > 1:
            gather_scores(scorer.out)

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    torch.cuda.synchronize(device)
    ms_e2e = e0.elapsed_time(e1)
    barrier()
    clocks.stop()
    if world > 1:
        t = torch.tensor([ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * n * args.steps / (ms_e2e / 1e3)

    # ---------------- optional parity check of THIS run's scores against the CPU oracle (every rank, its own shard) ----
    check = None
    if args.check > 0:
        check = check_against_oracle(args, model, device, rank, world)

    # ---------------- per-kernel breakdown of one step (eager, CUDA events per launch) ----------------
    kernels = {}
    if rank == 0:
        import model.genconvit as _gm
        _prev_ts = _gm.set_two_streams(False)      # serialised: a kernel's duration is its own, not a time-sliced one
        with torch.no_grad():
            for _ in range(2):
                lib.profile = []
                scorer._step()
                torch.cuda.synchronize(device)
                prof, lib.profile = lib.profile, None
        _gm.set_two_streams(_prev_ts)
        for name, work, s, e, _tag in prof:
            k = kernels.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
            k["launches"] += 1
            k["ms"] += s.elapsed_time(e)
            k["work"] += work
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_ms = sum(k["ms"] for k in kernels.values()) or 1.0
    # Kernels timed one by one in a short eager step run at boost clocks: their denominator is the BURST cuBLAS figure;
    # the whole step (graph replays back to back for the whole timed region) is set against the SUSTAINED one.
    peak_tf = float(peaks["bf16_tflops"])
    peak_tf_sustained = float(peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"])
    peak_gbs = float(peaks["hbm_gbs"])
    tensor_kernels = ("gemm_tcgen05", "mlp_fused", "conv3x3_tc")

    def roof(name, k):
        tensor = name in tensor_kernels or name.startswith("gemm")
        rate = k["work"] / (k["ms"] / 1e3) / (1e12 if tensor else 1e9) if k["ms"] > 0 else 0.0
        peak = peak_tf if tensor else peak_gbs
        return {"kernel": name, "bound": "tensor" if tensor else "hbm", "achieved": rate, "peak": peak,
                "unit": "TFLOP/s" if tensor else "GB/s", "frac": rate / peak if peak else None,
                "launches_per_step": k["launches"], "ms_per_step": k["ms"], "share_of_step": k["ms"] / total_ms}

    ranked = sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])
    rooflines = [roof(n, k) for n, k in ranked if k["ms"] / total_ms >= 0.02]
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh)
    top_name, top_k = ranked[0] if ranked else ("", {"ms": 0.0, "work": 0.0, "launches": 0})
    top = roof(top_name, top_k) if ranked else {}
    # all tensor-core work of the step (stand-alone GEMMs + fused MLP) against the tensor roof
    t_ms = sum(kernels[n]["ms"] for n in tensor_kernels if n in kernels)
    t_work = sum(kernels[n]["work"] for n in tensor_kernels if n in kernels)
    tensor_tflops = t_work / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0
    step_tflops = (value / world) * GFLOP_PER_FRAME_CONTRACTION / 1e3
    breakdown = {
        name: {"launches": k["launches"], "ms": round(k["ms"], 3), "share": round(k["ms"] / total_ms, 4),
               ("tflops" if (name in tensor_kernels or name.startswith("gemm")) else "gbs"):
                   round(k["work"] / (k["ms"] / 1e3) / (1e12 if (name in tensor_kernels or name.startswith("gemm")) else 1e9), 1)
                   if k["ms"] > 0 else 0}
        for name, k in ranked}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps, _ = cpu_oracle_fps(args.cpu_frames, threads)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"{args.cpu_frames} frames of the same workload (full GenConViT forward + pred_vid), "
                         f"torch CPU fp32, {threads} threads, warm-up of 16 frames, best of 2 passes; oracle port (mu once)"}

    eager = None
    if world == 1 and not args.no_gpu_eager:
        del hosts, outs
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args)
        hosts = [torch.empty(n * 224 * 224 * 3 * (4 if args.e2e_fp32 else 1), dtype=torch.uint8)]

    in_bytes = hosts[0].numel() * hosts[0].element_size()
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "full GenConViT (ED+VAE, pred_vid scoring) bs=256 224x224 per GPU, random-init weights",
                   "frames_per_gpu_per_step": n, "frames_per_video": fpv, "global_batch": n * world,
                   "parallelism": f"dp{world} (frames sharded, weights replicated, all-gather of per-video scores)",
                   "cuda_graph": not args.no_graph, "vae_eps": "fresh randn per step (reference behaviour)",
                   "streams": "ED and VAE networks overlap on two CUDA streams inside the step (per-kernel breakdown "
                              "and rooflines are measured with the step serialised on one stream)",
                   "l2": "per-step inputs (154 MB) and activations (GBs) exceed the 126 MB L2; no explicit flush",
                   "inputs": "`value`: pre-processed fp32 NCHW frames resident in HBM (the reference modules' input); "
                             "`e2e`: raw uint8 NHWC crops from pinned host memory, read by the first kernels of both "
                             "networks and normalised in registers (bit-identical), so e2e moves 4x fewer input bytes "
                             "through HBM as well as PCIe and can exceed `value`"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 2 * scorer.n_videos * 4,
                "api": "genconvit_b200.runtime.VideoScorer.submit (pinned host "
                       + ("fp32 NCHW pre-processed frames" if args.e2e_fp32
                          else "uint8 NHWC face crops, normalised inside the first conv / stem kernels")
                       + " -> host per-video class / score)"},
        "gpu_launches": scorer.launches_per_step * args.steps,
        "launches_per_step": scorer.launches_per_step,
        "roofline": dict(top, traffic=traffic.get(top_name),
                         peak_source=peaks["_source"] + (" bf16_tflops (cuBLAS bf16 burst: kernels are timed one by one)"
                                                         if top.get("bound") == "tensor" else " hbm_gbs (device copy)"),
                         how="dominant kernel of one eager step: sum of algorithmic work / sum of CUDA-event durations "
                             "over its launches",
                         tensor_kernels={"achieved": tensor_tflops, "frac": tensor_tflops / peak_tf, "peak": peak_tf,
                                         "frac_of_sustained": tensor_tflops / peak_tf_sustained,
                                         "ms_per_step": t_ms, "kernels": list(tensor_kernels),
                                         "def": "all launches of the tcgen05 kernels (GEMM, fused MLP, implicit-GEMM 3x3 conv) of one "
                                                "step, algorithmic FLOPs / CUDA-event time, against the burst peak"},
                         whole_step={"achieved": step_tflops, "frac": step_tflops / peak_tf_sustained,
                                     "peak": peak_tf_sustained, "frac_of_burst": step_tflops / peak_tf,
                                     "def": "frames/s/GPU x 29.49 GFLOP contraction per frame, against the sustained peak"}),
        "rooflines": rooflines,
        "traffic_per_shape": traffic.get("per_shape"),
        "kernels": breakdown,
        "clocks": clocks.summary(),
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "check": check,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="fp16", choices=["bf16", "fp16", "fp32"],
                    help="kernel precision; fp16 is BASELINE config 2 (the reference's own --fp16 mode); bf16 runs at the same speed")
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--fpv", type=int, default=16, help="frames per video")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=64, help="frames timed for cpu_baseline")
    ap.add_argument("--ref-frames", type=int, default=64, help="frames per step of --impl reference")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: where the reference runs")
    ap.add_argument("--ref-half", action="store_true", help="--impl reference on cuda: .half() like the reference's --fp16")
    ap.add_argument("--ref-batched", action="store_true", help="--impl reference: one forward per step instead of one per video")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port even if baseline/_ref is staged")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the gpu_eager_baseline side measurement")
    ap.add_argument("--e2e-fp32", action="store_true", help="e2e leg ships pre-processed fp32 frames instead of uint8 crops")
    ap.add_argument("--check", type=int, default=0, help="verify this many frames per rank against the CPU oracle")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: route everything libraries print (e.g. NCCL's version banner)
    # to stderr and keep the real stdout for the result.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
