"""Bulk scoring runtime: pinned-host frames in, per-video (class, score) out.

``VideoScorer`` is the batched, pipelined form of the reference's per-video loop
(prediction.py:231-266 -> pred_func.pred_vid): every step copies one batch of frames from
pinned host memory on a copy stream while the previous batch is still computing, replays the
whole GenConViT forward + fused scoring kernel as ONE CUDA graph, and reads back only the
per-video results.  Across GPUs the work is sharded by whole batches (frames are independent
until the per-video mean), weights are replicated, and the only collective is an all-gather of
the per-video scores (``gather_scores``).
"""
from __future__ import annotations

import torch

from . import engine
from . import lib as L


_MEAN = (0.485, 0.456, 0.406)     # reference dataset/loader.py:64-65 (normalize_data()["vid"]), as in model.pred_func
_STD = (0.229, 0.224, 0.225)


class VideoScorer:
    def __init__(self, model, batch_frames, frames_per_video, eps=None, use_graph=True, img=224):
        """model: model.genconvit.GenConViT on a CUDA device; batch_frames: frames per step (whole
        videos: a multiple of frames_per_video).  eps: None -> fresh N(0,1) noise every step like
        the reference (genconvit_vae.py:46); or a fixed [batch_frames,12544] tensor.

        ``submit`` takes either the reference's pre-processed frames (fp32 NCHW, model/pred_func.py:95-108) or the raw
        uint8 NHWC face crops that function starts from.  The latter cross PCIe at a quarter of the bytes and are never
        expanded: the first kernels of both networks (encoder conv 1, ConvNeXt stem) read the bytes and normalise in
        registers, bit-identically to the host arithmetic.  The graph input is double-buffered -- one captured graph per
        uint8 buffer, sharing one memory pool -- so batch k+1's copy lands while batch k computes, with no staging copy.
        (fp32 compute mode: the uint8 frames go through ``gcv_preprocess_frames`` into the fp32 input buffer.)"""
        if batch_frames % frames_per_video:
            raise ValueError("batch_frames must hold whole videos")
        self.model = model
        self.dev = next(model.parameters()).device
        self.n, self.fpv = batch_frames, frames_per_video
        self.n_videos = batch_frames // frames_per_video
        self.n_nets = 1 if model.net in ("ed", "vae") else 2
        dev = self.dev
        self.img = img
        self.x_static = torch.empty((batch_frames, 3, img, img), dtype=torch.float32, device=dev)
        self.x_stage = None                  # fp32 staging buffer, allocated on the first fp32 submit
        self.u8_stage = [None, None]         # uint8 NHWC staging buffers (double-buffered), on the first uint8 submit
        self.u8_turn = 0
        self.has_vae = model.net != "ed"
        self.fixed_eps = eps is not None
        self.eps = (eps.to(dev, torch.float32).contiguous() if eps is not None
                    else torch.empty((batch_frames, 12544), dtype=torch.float32, device=dev)) if self.has_vae else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ev_h2d = torch.cuda.Event()
        self.ev_stage_free = torch.cuda.Event()
        self.ev_u8_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.graph = None
        self.u8_graph = [None, None]         # one captured step per uint8 input buffer (16-bit modes)
        self.use_graph = use_graph
        from .modules import compute_dtype_of
        first = model.model_ed if model.net != "vae" else model.model_vae
        self.u8_fused = compute_dtype_of(first, getattr(first, "compute_dtype", None)) != torch.float32
        self.launches_per_step = 0
        self.out = None                      # [2, V] fp32 on device: row 0 = class, row 1 = score
        self._prepare(use_graph)

    # -- one step worth of kernels (eager or under capture) --
    def _step(self, frames=None):
        if self.has_vae and not self.fixed_eps:
            self.eps.normal_()
        x1, x2 = self.model.forward_parts(self.x_static if frames is None else frames, self.eps)
        if self.out is None:
            self.out = torch.empty((2, self.n_videos), dtype=torch.float32, device=self.dev)   # row 0 class, row 1 score
        # fused scoring straight from the two logit buffers: no torch.cat, no per-row copies -- the step launches only
        # this library's kernels (plus the eps generator)
        with torch.cuda.device(self.dev):
            L.score_videos_pair(x1, x2, self.n, self.fpv, self.out)

    def _prepare(self, use_graph):
        with torch.no_grad():
            self.x_static.zero_()
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                before = L.launches
                self._step()                 # warm-up: packs weights, sets kernel attributes
                self.launches_per_step = L.launches - before
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            if use_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step()
                self.graph = g

    def run_resident(self):
        """One step on whatever is in ``x_static`` (inputs already in HBM)."""
        with torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step()

    def _run_u8(self, i):
        """One step straight from uint8 buffer ``i`` (captured on first use; the second graph shares the first one's pool:
        the two never run at the same time)."""
        with torch.no_grad():
            frames = engine.U8Frames(self.u8_stage[i], _MEAN, _STD)
            if not self.use_graph:
                self._step(frames)
                return
            if self.u8_graph[i] is None:
                cur = torch.cuda.current_stream(self.dev)
                s = torch.cuda.Stream(device=self.dev)
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    self._step(frames)       # eager once: the results of this very batch, and a warm allocator
                cur.wait_stream(s)
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                other = self.u8_graph[i ^ 1] or self.graph
                with torch.cuda.graph(g, pool=other.pool() if other is not None else None):
                    self._step(frames)
                self.u8_graph[i] = g
                return                       # the eager run above already produced this batch's scores
            self.u8_graph[i].replay()

    def submit(self, frames_host, out_host):
        """Enqueue one batch: pinned-host frames -> device, forward + scoring, results -> pinned host.
        Asynchronous; overlaps this batch's H2D with the previous batch's compute.

        frames_host: uint8 [N,H,W,3] face crops (preferred: 1 byte per value over PCIe, normalised on the GPU into the
        graph's input buffer) or fp32 [N,3,H,W] already pre-processed like model.pred_func.preprocess_frame."""
        cur = torch.cuda.current_stream(self.dev)
        if frames_host.dtype == torch.uint8:
            if tuple(frames_host.shape) != (self.n, self.img, self.img, 3):
                raise ValueError(f"uint8 frames must be [{self.n},{self.img},{self.img},3], got {tuple(frames_host.shape)}")
            i = self.u8_turn
            self.u8_turn ^= 1
            if self.u8_stage[i] is None:
                self.u8_stage[i] = torch.empty((self.n, self.img, self.img, 3), dtype=torch.uint8, device=self.dev)
                self.ev_u8_free[i].record(cur)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.ev_u8_free[i])          # the step that last read this buffer is done
                self.u8_stage[i].copy_(frames_host, non_blocking=True)
                self.ev_h2d.record(self.copy_stream)
            cur.wait_event(self.ev_h2d)
            if self.u8_fused:
                self._run_u8(i)
                self.ev_u8_free[i].record(cur)
                out_host.copy_(self.out, non_blocking=True)
                return
            with torch.cuda.device(self.dev):
                L.preprocess_frames(self.u8_stage[i], self.x_static, self.n, self.img, self.img, _MEAN, _STD)
            self.ev_u8_free[i].record(cur)
        else:
            if self.x_stage is None:
                self.x_stage = torch.empty_like(self.x_static)
                self.ev_stage_free.record(cur)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.ev_stage_free)
                self.x_stage.copy_(frames_host, non_blocking=True)
                self.ev_h2d.record(self.copy_stream)
            cur.wait_event(self.ev_h2d)
            self.x_static.copy_(self.x_stage, non_blocking=True)
            self.ev_stage_free.record(cur)
        self.run_resident()
        out_host.copy_(self.out, non_blocking=True)

    def score(self, frames_host):
        """Synchronous convenience: -> (classes [V] int64, scores [V] float32) on the host."""
        out = torch.empty((2, self.n_videos), dtype=torch.float32).pin_memory()
        self.submit(frames_host, out)
        torch.cuda.current_stream(self.dev).synchronize()
        return out[0].to(torch.int64), out[1].clone()


def shard_videos(n_videos: int, rank: int, world: int):
    """Whole videos per rank, contiguous, remainder spread over the first ranks."""
    base, rem = divmod(n_videos, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_scores(local: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the path: all-gather of per-video results ([2,V_local] -> [world,2,V_local]).
    NCCL over NVLink on GPUs (<= 32 KB: latency-bound); gloo in the CPU tests."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    local = local.contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out.view((world,) + tuple(local.shape))
