"""Host-side forward of GenConViT on the sm_100a kernel library.

Everything here is orchestration: weights are re-packed once into kernel layouts
(NHWC / K-major, BatchNorm folded, the VAE ``mu`` matrix permuted from the
reference's NCHW-flatten order to NHWC), and a forward is a fixed sequence of
C-ABI kernel launches on the current torch stream (capturable in a CUDA graph).
torch supplies device memory and the stream only; no torch operator touches an
activation.

Reference call sites are cited per function; the arithmetic contract is the
CPU oracle (``oracle/``), which is never imported from here.
"""
from __future__ import annotations

import os

import torch

from . import lib as L

DEPTHS = (3, 3, 9, 3)
DIMS = (96, 192, 384, 768)
# the fused fc1->GELU->fc2 kernel (stages 0-1); GCV_NO_FUSED_MLP=1 falls back to two GEMM launches (A/B timing)
FUSED_MLP = os.environ.get("GCV_NO_FUSED_MLP", "0") != "1"
# 16-bit modes: depthwise conv on the tensor cores writing the un-normalised output + LayerNorm partial sums, with the
# LayerNorm folded into fc1 (weights pre-scaled by the LN weight, row statistics applied in the fc1 epilogue), so the
# block makes no separate normalisation pass.  GCV_NO_LNFOLD=1 keeps the explicit dwconv7_ln kernel (A/B timing).
LN_FOLD = os.environ.get("GCV_NO_LNFOLD", "0") != "1"
# 16-bit modes: the stem (4x4 patchify + conv + LayerNorm2d) as ONE tensor-core kernel reading the frames directly.
# GCV_NO_FUSED_STEM=1 keeps im2col + GEMM + row LayerNorm (A/B timing).
FUSED_STEM = os.environ.get("GCV_NO_FUSED_STEM", "0") != "1"
# GCV_LN_FINALIZE=1: reduce the LayerNorm partial sums of stages 2-3 with a separate gcv_ln_finalize launch per block
# instead of inside the fc1 GEMM (its statistics warps); A/B timing only
LN_FINALIZE = os.environ.get("GCV_LN_FINALIZE", "0") == "1"


def _on_device(fn):
    """Run a forward with the tensor's device as the current CUDA device: the C ABI launches on the current device's
    stream and keeps per-device kernel attributes, so a model on cuda:1 must not launch from cuda:0's context."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, x, *a, **k):
        L.require_cuda_tensor(x, type(self).__name__)
        with torch.cuda.device(x.device):
            return fn(self, x, *a, **k)
    return wrapped


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _cd(t, dev, dt):
    return t.detach().to(device=dev, dtype=torch.float32).to(dt).contiguous()


def _empty(shape, dt, dev):
    return torch.empty(shape, dtype=dt, device=dev)


class PackedConvNeXt:
    """Kernel-layout copy of a timm-style ConvNeXt-T ``state_dict`` (prefix-free keys)."""

    def __init__(self, sd, dev, dt):
        self.dev, self.dt = dev, dt
        # variant (convnext_tiny / convnext_large, reference prediction.py:314-318) read off the state_dict itself
        DIMS = tuple(sd[f"stages.{s}.blocks.0.gamma"].shape[0] for s in range(4))
        DEPTHS = tuple(sum(1 for k in sd if k.startswith(f"stages.{s}.blocks.") and k.endswith(".gamma")) for s in range(4))
        self.dims, self.depths = DIMS, DEPTHS
        w = sd["stem.0.weight"]                                   # [96,3,4,4] -> [96, (kh,kw,c)]
        self.stem_w = _cd(w.permute(0, 2, 3, 1).reshape(w.shape[0], 48), dev, dt)
        self.stem_w_oihw = _cd(w.reshape(w.shape[0], 48), dev, dt)  # [96, (c,kh,kw)]: the fused stem reads NCHW frames
        self.stem_b = _f32(sd["stem.0.bias"], dev)
        self.stem_ln = (_f32(sd["stem.1.weight"], dev), _f32(sd["stem.1.bias"], dev))
        self.stages = []
        for s, depth in enumerate(DEPTHS):
            st = {"blocks": []}
            p = f"stages.{s}."
            if s > 0:
                st["ds_ln"] = (_f32(sd[p + "downsample.0.weight"], dev), _f32(sd[p + "downsample.0.bias"], dev))
                w = sd[p + "downsample.1.weight"]                 # [Co,Ci,2,2] -> [Co, (kh,kw,ci)]
                st["ds_w"] = _cd(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), dev, dt)
                st["ds_b"] = _f32(sd[p + "downsample.1.bias"], dev)
            for k in range(depth):
                q = f"{p}blocks.{k}."
                c = DIMS[s]
                fold = {}
                if dt != torch.float32:
                    # LayerNorm folded into fc1: W1' = W1 diag(ln_w) (rounded to the kernel dtype), its column sums
                    # taken from the ROUNDED matrix (so a constant row still cancels exactly), b1' = b1 + W1 ln_b
                    w1 = sd[q + "mlp.fc1.weight"].detach().to(device=dev, dtype=torch.float32)
                    lw, lb = _f32(sd[q + "norm.weight"], dev), _f32(sd[q + "norm.bias"], dev)
                    w1f = (w1 * lw[None, :]).to(dt).contiguous()
                    fold = dict(fc1_wf=w1f, fc1_cs=w1f.float().sum(dim=1).contiguous(),
                                fc1_bf=(_f32(sd[q + "mlp.fc1.bias"], dev) + w1 @ lb).contiguous())
                st["blocks"].append(dict(fold, 
                    taps=_f32(sd[q + "conv_dw.weight"].reshape(c, 49).t(), dev),     # [49, C]
                    dw_b=_f32(sd[q + "conv_dw.bias"], dev),
                    ln_w=_f32(sd[q + "norm.weight"], dev), ln_b=_f32(sd[q + "norm.bias"], dev),
                    fc1_w=_cd(sd[q + "mlp.fc1.weight"], dev, dt), fc1_b=_f32(sd[q + "mlp.fc1.bias"], dev),
                    fc2_w=_cd(sd[q + "mlp.fc2.weight"], dev, dt), fc2_b=_f32(sd[q + "mlp.fc2.bias"], dev),
                    gamma=_f32(sd[q + "gamma"], dev)))
            self.stages.append(st)
        self.head_ln = (_f32(sd["head.norm.weight"], dev), _f32(sd["head.norm.bias"], dev))
        self.head_w = _cd(sd["head.fc.weight"], dev, dt)
        self.head_b = _f32(sd["head.fc.bias"], dev)

    def forward_tokens(self, a0, segments, outs, act, backend=L.GEMM_AUTO):
        """Run stem-GEMM .. head on pre-patchified tokens.

        a0       [sum_i B_i*H_i*W_i, 48] stem im2col rows, segment after segment
        segments list of (B, H, W) *after* the 4x4 stem (e.g. (N, 56, 56)); segments share the
                 weights, so all GEMMs run once over the concatenated tokens
        outs     list of (first_image, n_images, view, ldd): head logits of images
                 [first, first+n) are written to ``view`` (a [n, >=1000]-column window with
                 leading dimension ldd) with ``act`` already applied -- the reference applies
                 its GELU/ReLU to the concatenated backbone logits before ``fc``
        timm ConvNeXt.forward = head(stages(stem(x))) as reached from reference
        genconvit_ed.py:82-83 / genconvit_vae.py:111-112.
        """
        dev, dt = self.dev, self.dt
        DIMS = self.dims
        if isinstance(a0, list):
            # 16-bit modes: a0 is a list of frame sources (tensor, nchw?, n_images, H, W), one per run of tokens; the
            # fused stem kernel turns each straight into normalised tokens (no im2col matrix, no separate LayerNorm)
            m = sum(n * (hh // 4) * (ww // 4) for _, _, n, hh, ww in a0)
            x = _empty((m, DIMS[0]), dt, dev)
            r = 0
            for src, nchw, n, hh, ww in a0:
                if isinstance(src, U8Frames):
                    L.stem_fused_u8(src.t, x[r:], self.stem_w_oihw, self.stem_b, self.stem_ln[0], self.stem_ln[1], 1e-6,
                                    n, hh, ww, src.mean, src.std)
                else:
                    L.stem_fused(src, x[r:], self.stem_w_oihw if nchw else self.stem_w, self.stem_b, self.stem_ln[0],
                                 self.stem_ln[1], 1e-6, n, hh, ww, nchw)
                r += n * (hh // 4) * (ww // 4)
        else:
            m = a0.shape[0]
            x = _empty((m, DIMS[0]), dt, dev)
            L.gemm(a0, self.stem_w, x, m, DIMS[0], 48, bias=self.stem_b, backend=backend)
            L.layernorm_rows(x, x, self.stem_ln[0], self.stem_ln[1], 1e-6, m, DIMS[0])
        segs = list(segments)
        for s, st in enumerate(self.stages):
            c = DIMS[s]
            if s > 0:
                cin = DIMS[s - 1]
                new = [(b, h // 2, w // 2) for b, h, w in segs]
                m2 = sum(b * h * w for b, h, w in new)
                a = _empty((m2, 4 * cin), dt, dev)
                ro = ri = 0
                for (b, h, w), (_, h2, w2) in zip(segs, new):
                    L.ln_patchify2(x[ri:], a[ro:], st["ds_ln"][0], st["ds_ln"][1], 1e-6, b, h, w, cin)
                    ri += b * h * w
                    ro += b * h2 * w2
                x = _empty((m2, c), dt, dev)
                L.gemm(a, st["ds_w"], x, m2, c, 4 * cin, bias=st["ds_b"], backend=backend)
                segs, m = new, m2
            y = _empty((m, c), dt, dev)
            fused = FUSED_MLP and backend == L.GEMM_AUTO and L.mlp_fused_supported(dt, c)
            # the folded-LayerNorm GEMM epilogue stages its per-column vectors in shared memory (N <= 3072: C <= 768)
            fold = LN_FOLD and backend == L.GEMM_AUTO and dt != torch.float32 and (4 * c <= 3072 or
                                                                                    L.mlp_fused_supported(dt, c))
            stats = _empty((m, c // 32, 2), torch.float32, dev) if fold else None
            hid = None if fused else _empty((m, 4 * c), dt, dev)
            for blk in st["blocks"]:
                r = 0
                for b, h, w in segs:
                    if fold:
                        L.dwconv7_stats(x[r:], y[r:], stats[r:], blk["taps"], blk["dw_b"], b, h, w, c)
                    else:
                        L.dwconv7_ln(x[r:], y[r:], blk["taps"], blk["dw_b"], blk["ln_w"], blk["ln_b"], 1e-6, b, h, w, c)
                    r += b * h * w
                if fused and fold:
                    L.mlp_fused_ln(y, stats, 1e-6, blk["fc1_wf"], blk["fc1_bf"], blk["fc1_cs"], blk["fc2_w"],
                                   blk["fc2_b"], blk["gamma"], x, m, c)
                elif fused:
                    L.mlp_fused(y, blk["fc1_w"], blk["fc1_b"], blk["fc2_w"], blk["fc2_b"], blk["gamma"], x, m, c)
                else:
                    if fold and LN_FINALIZE:
                        # partial sums -> one (rstd, -mean*rstd) pair per row by a separate launch (A/B timing)
                        rowstat = _empty((m, 2), torch.float32, dev)
                        L.ln_finalize(stats, rowstat, m, c, 1e-6)
                        L.gemm(y, blk["fc1_wf"], hid, m, 4 * c, c, bias=blk["fc1_bf"], act=L.ACT_GELU, ln_stats=rowstat,
                               ln_colsum=blk["fc1_cs"], ln_eps=1e-6, backend=backend)
                    elif fold:
                        # the GEMM's statistics warps reduce the partial sums to (rstd, -mean*rstd) per row, a tile ahead
                        L.gemm(y, blk["fc1_wf"], hid, m, 4 * c, c, bias=blk["fc1_bf"], act=L.ACT_GELU, ln_stats=stats,
                               ln_colsum=blk["fc1_cs"], ln_eps=1e-6, backend=backend)
                    else:
                        L.gemm(y, blk["fc1_w"], hid, m, 4 * c, c, bias=blk["fc1_b"], act=L.ACT_GELU, backend=backend)
                    L.gemm(hid, blk["fc2_w"], x, m, c, 4 * c, bias=blk["fc2_b"], gamma=blk["gamma"], residual=x,
                           ldr=c, backend=backend)
        c = DIMS[3]
        n_img = sum(b for b, _, _ in segs)
        pooled = _empty((n_img, c), dt, dev)
        r = i = 0
        for b, h, w in segs:
            L.pool_ln(x[r:], pooled[i:], self.head_ln[0], self.head_ln[1], 1e-6, b, h * w, c)
            r += b * h * w
            i += b
        for first, cnt, view, ldd in outs:
            L.gemm(pooled[first:], self.head_w, view, cnt, 1000, c, bias=self.head_b, act=act, ldd=ldd,
                   out_f32=view.dtype == torch.float32 and dt != torch.float32, backend=backend)

    @_on_device
    def forward_images(self, x, act=L.ACT_NONE, backend=L.GEMM_AUTO):
        """``backbone(x)`` for fp32 NCHW frames -> fp32 [N,1000] ImageNet logits."""
        n, _, hh, ww = x.shape
        if FUSED_STEM and self.dt != torch.float32 and backend == L.GEMM_AUTO and self.dims[0] == 96:
            a0 = [(x, True, n, hh, ww)]
        else:
            a0 = _empty((n * (hh // 4) * (ww // 4), 48), self.dt, self.dev)
            L.stem_patchify_nchw(x, a0, n, hh, ww)
        out = _empty((n, 1000), torch.float32, self.dev)
        self.forward_tokens(a0, [(n, hh // 4, ww // 4)], [(0, n, out, 1000)], act, backend)
        return out


SWIN_DEPTHS = (2, 2, 6, 2)           # swin_tiny_patch4_window7_224; the large variant is read off the state_dict
SWIN_HEADS = (3, 6, 12, 24)


class PackedSwin:
    """Kernel-layout copy of a timm-style swin_tiny_patch4_window7_224 ``state_dict`` (prefix-free keys) and its
    forward on the CUDA kernels.  The reference builds this network as ``self.embedder``
    (model/genconvit_ed.py:69, model/genconvit_vae.py:96) and hands it to HybridEmbed, but no GenConViT logit depends
    on it; it is provided as the standalone callable ``model.embedder(x)``.
    Arithmetic contract: oracle/backbones.py swin_forward (pinned against torchvision)."""

    def __init__(self, sd, dev, dt):
        self.dev, self.dt = dev, dt
        self.embed = sd["patch_embed.proj.weight"].shape[0]
        SWIN_DEPTHS = tuple(sum(1 for k in sd if k.startswith(f"layers.{l}.blocks.") and k.endswith(".norm1.weight"))
                            for l in range(4))
        self.heads = tuple(sd[f"layers.{l}.blocks.0.attn.relative_position_bias_table"].shape[1] for l in range(4))
        w = sd["patch_embed.proj.weight"]                         # [96,3,4,4] -> [96, (kh,kw,c)]
        self.pe_w = _cd(w.permute(0, 2, 3, 1).reshape(w.shape[0], 48), dev, dt)
        self.pe_b = _f32(sd["patch_embed.proj.bias"], dev)
        self.pe_ln = (_f32(sd["patch_embed.norm.weight"], dev), _f32(sd["patch_embed.norm.bias"], dev))
        self.layers = []
        for l, depth in enumerate(SWIN_DEPTHS):
            blocks = []
            for k in range(depth):
                q = f"layers.{l}.blocks.{k}."
                blocks.append(dict(
                    n1=(_f32(sd[q + "norm1.weight"], dev), _f32(sd[q + "norm1.bias"], dev)),
                    n2=(_f32(sd[q + "norm2.weight"], dev), _f32(sd[q + "norm2.bias"], dev)),
                    qkv_w=_cd(sd[q + "attn.qkv.weight"], dev, dt), qkv_b=_f32(sd[q + "attn.qkv.bias"], dev),
                    proj_w=_cd(sd[q + "attn.proj.weight"], dev, dt), proj_b=_f32(sd[q + "attn.proj.bias"], dev),
                    table=_f32(sd[q + "attn.relative_position_bias_table"], dev),
                    fc1_w=_cd(sd[q + "mlp.fc1.weight"], dev, dt), fc1_b=_f32(sd[q + "mlp.fc1.bias"], dev),
                    fc2_w=_cd(sd[q + "mlp.fc2.weight"], dev, dt), fc2_b=_f32(sd[q + "mlp.fc2.bias"], dev)))
            layer = {"blocks": blocks}
            if l < 3:
                q = f"layers.{l}.downsample."
                layer["merge_ln"] = (_f32(sd[q + "norm.weight"], dev), _f32(sd[q + "norm.bias"], dev))
                layer["merge_w"] = _cd(sd[q + "reduction.weight"], dev, dt)
            self.layers.append(layer)
        self.norm = (_f32(sd["norm.weight"], dev), _f32(sd["norm.bias"], dev))
        self.head_w, self.head_b = _cd(sd["head.weight"], dev, dt), _f32(sd["head.bias"], dev)

    @_on_device
    def forward_images(self, x, backend=L.GEMM_AUTO):
        """fp32 NCHW 224x224 frames -> fp32 [N,1000] logits (timm SwinTransformer.forward)."""
        dev, dt = self.dev, self.dt
        n, _, hh, ww = x.shape
        if (hh, ww) != (224, 224):
            raise L.GcvError(f"swin_*_patch4_window7_224 takes 224x224 frames, got {hh}x{ww}")
        res, c = 56, self.embed
        m = n * res * res
        a0 = _empty((m, 48), dt, dev)
        L.stem_patchify_nchw(x, a0, n, hh, ww)
        t = _empty((m, c), dt, dev)
        L.gemm(a0, self.pe_w, t, m, c, 48, bias=self.pe_b, backend=backend)
        L.layernorm_rows(t, t, self.pe_ln[0], self.pe_ln[1], 1e-5, m, c)
        for l, layer in enumerate(self.layers):
            heads = self.heads[l]
            h = _empty((m, c), dt, dev)
            qkv = _empty((m, 3 * c), dt, dev)
            att = _empty((m, c), dt, dev)
            hid = _empty((m, 4 * c), dt, dev)
            for k, blk in enumerate(layer["blocks"]):
                shift = 0 if (k % 2 == 0 or res <= 7) else 3
                # x = x + proj(W-MSA(norm1(x)))
                L.layernorm_rows(t, h, blk["n1"][0], blk["n1"][1], 1e-5, m, c)
                L.gemm(h, blk["qkv_w"], qkv, m, 3 * c, c, bias=blk["qkv_b"], backend=backend)
                L.swin_window_attention(qkv, att, blk["table"], n, res, c, heads, shift)
                L.gemm(att, blk["proj_w"], t, m, c, c, bias=blk["proj_b"], residual=t, ldr=c, backend=backend)
                # x = x + fc2(GELU(fc1(norm2(x))))
                L.layernorm_rows(t, h, blk["n2"][0], blk["n2"][1], 1e-5, m, c)
                L.gemm(h, blk["fc1_w"], hid, m, 4 * c, c, bias=blk["fc1_b"], act=L.ACT_GELU, backend=backend)
                L.gemm(hid, blk["fc2_w"], t, m, c, 4 * c, bias=blk["fc2_b"], residual=t, ldr=c, backend=backend)
            if l < 3:
                m2 = m // 4
                mg = _empty((m2, 4 * c), dt, dev)
                L.swin_patch_merge(t, mg, n, res, c)
                L.layernorm_rows(mg, mg, layer["merge_ln"][0], layer["merge_ln"][1], 1e-5, m2, 4 * c)
                t = _empty((m2, 2 * c), dt, dev)
                L.gemm(mg, layer["merge_w"], t, m2, 2 * c, 4 * c, backend=backend)
                m, res, c = m2, res // 2, 2 * c
        L.layernorm_rows(t, t, self.norm[0], self.norm[1], 1e-5, m, c)
        pooled = _empty((n, c), dt, dev)
        L.mean_tokens(t, pooled, n, res * res, c)
        out = _empty((n, 1000), torch.float32, dev)
        L.gemm(pooled, self.head_w, out, n, 1000, c, bias=self.head_b, out_f32=dt != torch.float32, backend=backend)
        return out


class U8Frames:
    """Raw uint8 NHWC face crops [N,H,W,3] on the GPU plus the Normalize(mean, std) the reference applies on the host
    (model/pred_func.py:95-108, dataset/loader.py:63-77).  Accepted wherever the engine takes the pre-processed fp32 NCHW
    frames (16-bit modes): the first-touch kernels (encoder conv 1, ConvNeXt stem) read the bytes and normalise in
    registers -- bit-identical to preprocessing first -- so the fp32 frames never exist."""

    def __init__(self, t, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3 or not t.is_contiguous():
            raise ValueError(f"U8Frames takes a contiguous uint8 [N,H,W,3] tensor, got {t.dtype} {tuple(t.shape)}")
        self.t, self.mean, self.std = t, tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.shape = (t.shape[0], 3, t.shape[1], t.shape[2])       # the NCHW shape of the frames it stands for
        self.device, self.is_cuda = t.device, t.is_cuda


def _first_conv(x, y, w, b, stride, act, pool, n, hh, ww, dt):
    if isinstance(x, U8Frames):
        if dt == torch.float32:
            raise ValueError("uint8 frames are consumed by the 16-bit kernels; pre-process them for the fp32 mode")
        L.conv3x3_first_u8(x.t, y, w, b, stride, act, pool, n, hh, ww, x.mean, x.std)
    else:
        L.conv3x3_first(x, y, w, b, stride, act, pool, n, hh, ww)


def _pack_conv3x3(w, dev, dt):
    """[Co,Ci,3,3] -> [Co, (kh,kw,ci)] matching gcv_im2col3x3's column order."""
    return _cd(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), dev, dt)


def _pack_convt(w, b, dev, dt):
    """ConvTranspose2d k2 s2 weight [Ci,Co,2,2] -> GEMM B [(i,j,co), Ci] and the bias repeated per tap.
    The 16 -> 3 output layer is also kept in fp32 for the streaming kernel (third entry; None for the others)."""
    co = w.shape[1]
    wm = w.permute(2, 3, 1, 0).reshape(4 * co, w.shape[0])
    return _cd(wm, dev, dt), _f32(b.repeat(4), dev), (_f32(wm, dev) if (w.shape[0], co) == (16, 3) else None)


def _run_convt_stack(x, layers, b, h, w, act, dt, dev, backend):
    """x: [B*h*w, Ci] tokens -> NHWC image after the k2s2 transposed-conv stack.  Wide layers: tcgen05 GEMM with the
    pixel-shuffle epilogue; 64 -> 32 and the fused 32 -> 16 -> 3 tail (16-bit modes): the HMMA kernel of convt_mma.cu."""
    k = 0
    while k < len(layers):
        wt, bias, w32 = layers[k]
        co, ci = wt.shape[0] // 4, wt.shape[1]
        small = dt != torch.float32 and (b * h * w) % 16 == 0 and backend == L.GEMM_AUTO
        if small and (ci, co) == (32, 16) and k + 2 == len(layers) and layers[k + 1][2] is not None:
            out = _empty((b * 16 * h * w, 3), dt, dev)
            L.convt2x2_mma(x, out, wt, bias, act, b, h, w, ci, w2=layers[k + 1][0], b2=layers[k + 1][1])
            return out, 4 * h, 4 * w
        out = _empty((b * 4 * h * w, co), dt, dev)
        if w32 is not None:
            L.convt2x2_small(x, out, w32, bias, act, b, h, w, ci, co)
        elif small and (ci, co) in ((64, 32), (32, 16)):
            L.convt2x2_mma(x, out, wt, bias, act, b, h, w, ci)
        else:
            L.gemm(x, wt, out, b * h * w, 4 * co, ci, bias=bias, act=act, store=L.STORE_PIXEL_SHUFFLE2,
                   ps=(h, w, co), backend=backend)
        x, h, w = out, 2 * h, 2 * w
        k += 1
    return x, h, w


def _heads(cat, n, fc_w, fc_b, fc2_w, fc2_b, act, dt, dev, backend):
    """fc2(act(fc(cat))) -- ``cat`` already holds act(concatenated backbone logits).  -> fp32 [n,2]."""
    hid = _empty((n, 512), dt, dev)          # 500 columns used; leading dimension padded for 16-byte rows
    L.gemm(cat, fc_w, hid, n, 500, 2000, bias=fc_b, act=act, ldd=512, backend=backend)
    logits = _empty((n, 2), torch.float32, dev)
    L.gemm(hid, fc2_w, logits, n, 2, 500, lda=512, bias=fc2_b, out_f32=dt != torch.float32, backend=backend)
    return logits


class PackedED:
    """Network A (reference model/genconvit_ed.py:63-89)."""

    def __init__(self, sd, dev, dt):
        self.dev, self.dt = dev, dt
        self.enc0_w = _f32(sd["encoder.features.0.weight"], dev)          # direct fp32 conv, OIHW
        self.enc0_b = _f32(sd["encoder.features.0.bias"], dev)
        self.enc = [(_pack_conv3x3(sd[f"encoder.features.{i}.weight"], dev, dt), _f32(sd[f"encoder.features.{i}.bias"], dev))
                    for i in (3, 6, 9, 12)]
        self.dec = [_pack_convt(sd[f"decoder.features.{i}.weight"], sd[f"decoder.features.{i}.bias"], dev, dt)
                    for i in (0, 2, 4, 6, 8)]
        self.backbone = PackedConvNeXt({k[len("backbone."):]: v for k, v in sd.items()
                                        if k.startswith("backbone.") and not k.startswith("backbone.patch_embed.")}, dev, dt)
        self.fc_w, self.fc_b = _cd(sd["fc.weight"], dev, dt), _f32(sd["fc.bias"], dev)
        self.fc2_w, self.fc2_b = _cd(sd["fc2.weight"], dev, dt), _f32(sd["fc2.bias"], dev)

    def encode(self, x, backend=L.GEMM_AUTO):
        """Encoder (genconvit_ed.py:13-36): -> ([N*49, 256] tokens, 7, 7)."""
        dev, dt = self.dev, self.dt
        n, _, hh, ww = x.shape
        h, w, c = hh // 2, ww // 2, 16
        e = _empty((n * h * w, c), dt, dev)
        _first_conv(x, e, self.enc0_w, self.enc0_b, 1, L.ACT_RELU, True, n, hh, ww, dt)
        for wt, bias in self.enc:
            co = wt.shape[0]
            if (c, co) == (16, 32) and dt != torch.float32 and h % 2 == 0 and w % 2 == 0:
                # the widest layer (112 x 112 pixels per frame): direct tensor-core conv + ReLU + pool, no im2col matrix
                nxt = _empty((n * (h // 2) * (w // 2), co), dt, dev)
                L.conv3x3_c16(e, nxt, wt, bias, 1, L.ACT_RELU, True, n, h, w)
                e, h, w, c = nxt, h // 2, w // 2, co
                continue
            if (c, co) == (32, 64) and dt != torch.float32 and h % 2 == 0 and w % 2 == 0:
                nxt = _empty((n * (h // 2) * (w // 2), co), dt, dev)
                L.conv3x3_c32(e, nxt, wt, bias, 1, L.ACT_RELU, True, n, h, w)
                e, h, w, c = nxt, h // 2, w // 2, co
                continue
            if backend == L.GEMM_AUTO and L.conv3x3_tc_supported(dt, c, co) and h % 2 == 0 and w % 2 == 0:
                # 64 -> 128, 128 -> 256: tcgen05 implicit GEMM (TMA-gathered taps), ReLU + pool in the epilogue
                nxt = _empty((n * (h // 2) * (w // 2), co), dt, dev)
                L.conv3x3_tc(e, nxt, wt, bias, 1, L.ACT_RELU, True, n, h, w, c, co)
                e, h, w, c = nxt, h // 2, w // 2, co
                continue
            a = _empty((n * h * w, 9 * c), dt, dev)
            L.im2col3x3(e, a, n, h, w, c, 1)
            full = _empty((n * h * w, co), dt, dev)
            L.gemm(a, wt, full, n * h * w, co, 9 * c, bias=bias, act=L.ACT_RELU, backend=backend)
            e = _empty((n * (h // 2) * (w // 2), co), dt, dev)
            L.maxpool2(full, e, n, h, w, co)
            h, w, c = h // 2, w // 2, co
        return e, h, w

    def decode(self, e, n, h, w, backend=L.GEMM_AUTO):
        """Decoder (genconvit_ed.py:43-61): -> NHWC [N, 32h, 32w, 3]."""
        return _run_convt_stack(e, self.dec, n, h, w, L.ACT_RELU, self.dt, self.dev, backend)

    @_on_device
    def forward(self, x, backend=L.GEMM_AUTO):
        """GenConViTED.forward (genconvit_ed.py:77-89) -> fp32 logits [N,2].

        The two backbone passes share weights, so decoded and original frames run as
        one 2N-image batch; rows [0,N) are the decoded images (cat order x1=decoded, x2=images).
        """
        dev, dt = self.dev, self.dt
        n, _, hh, ww = x.shape
        if x.shape[1] != 3 or hh % 32 or ww % 32 or hh < 32 or ww < 32:
            # five 2x2 max-pools, then five k2s2 transposed convs must give back H x W (genconvit_ed.py:13-61,81-83)
            raise ValueError(f"GenConViTED takes [N,3,H,W] frames with H, W multiples of 32, got {tuple(x.shape)}")
        e, h, w = self.encode(x, backend)
        dec, dh, dw = self.decode(e, n, h, w, backend)
        assert (dh, dw) == (hh, ww)
        th, tw = hh // 4, ww // 4
        if FUSED_STEM and dt != torch.float32 and backend == L.GEMM_AUTO and self.backbone.dims[0] == 96:
            a0 = [(dec, False, n, hh, ww), (x, True, n, hh, ww)]
        elif isinstance(x, U8Frames):
            raise ValueError("uint8 frames need the 16-bit fused stem (convnext_tiny, bf16 / fp16)")
        else:
            a0 = _empty((2 * n * th * tw, 48), dt, dev)
            L.stem_patchify_nhwc(dec, a0, n, hh, ww)
            L.stem_patchify_nchw(x, a0[n * th * tw:], n, hh, ww)
        cat = _empty((n, 2000), dt, dev)
        self.backbone.forward_tokens(a0, [(2 * n, th, tw)], [(0, n, cat, 2000), (n, n, cat[:, 1000:], 2000)],
                                     L.ACT_GELU, backend)
        return _heads(cat, n, self.fc_w, self.fc_b, self.fc2_w, self.fc2_b, L.ACT_GELU, dt, dev, backend)


class PackedVAE:
    """Network B (reference model/genconvit_vae.py:90-116)."""

    def __init__(self, sd, dev, dt, with_var=False):
        self.dev, self.dt = dev, dt
        self.enc = []
        for i in (0, 3, 6, 9):
            # BatchNorm2d in eval mode (running stats, eps 1e-5) folded into the conv (genconvit_vae.py:16-29)
            b = f"encoder.features.{i + 1}."
            scale = sd[b + "weight"].float() / torch.sqrt(sd[b + "running_var"].float() + 1e-5)
            w = sd[f"encoder.features.{i}.weight"].float() * scale.view(-1, 1, 1, 1)
            bias = (sd[f"encoder.features.{i}.bias"].float() - sd[b + "running_mean"].float()) * scale + sd[b + "bias"].float()
            if i == 0:
                self.enc0_w, self.enc0_b = _f32(w, dev), _f32(bias, dev)
            else:
                self.enc.append((_pack_conv3x3(w, dev, dt), _f32(bias, dev)))
        self.mu_w, self.mu_b = self._pack_latent(sd["encoder.mu.weight"], sd["encoder.mu.bias"])
        self.var_w = self.var_b = None
        self._var_src = (sd["encoder.var.weight"], sd["encoder.var.bias"]) if with_var else None
        self.dec = [_pack_convt(sd[f"decoder.features.{i}.weight"], sd[f"decoder.features.{i}.bias"], dev, dt)
                    for i in (0, 2, 4, 6)]
        pre = "convnext_backbone."
        self.backbone = PackedConvNeXt({k[len(pre):]: v for k, v in sd.items()
                                        if k.startswith(pre) and not k.startswith(pre + "patch_embed.")}, dev, dt)
        self.fc_w, self.fc_b = _cd(sd["fc.weight"], dev, dt), _f32(sd["fc.bias"], dev)
        self.fc2_w, self.fc2_b = _cd(sd["fc2.weight"], dev, dt), _f32(sd["fc2.bias"], dev)

    def _pack_latent(self, w, b):
        """mu/var Linear(25088 -> 12544): the reference flattens NCHW on both sides
        (input c*196+hw, genconvit_vae.py:53; output c*49+hw, :83).  Activations here are NHWC,
        so permute rows to (hw7, c7) and columns to (hw14, c14) once."""
        dev, dt = self.dev, self.dt
        w = w.detach().to(dev)
        w = w.view(256, 49, 128, 196).permute(1, 0, 3, 2).to(dt).reshape(12544, 25088).contiguous()
        return w, _f32(b.view(256, 49).t().reshape(-1), dev)

    def encode_features(self, x, backend=L.GEMM_AUTO):
        """Encoder.features (genconvit_vae.py:15-31) -> [N, 14*14*128] NHWC-flattened."""
        dev, dt = self.dev, self.dt
        n, _, hh, ww = x.shape
        h, w, c = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1, 16
        e = _empty((n * h * w, c), dt, dev)
        _first_conv(x, e, self.enc0_w, self.enc0_b, 2, L.ACT_LEAKY, False, n, hh, ww, dt)
        for wt, bias in self.enc:
            co = wt.shape[0]
            h2, w2 = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            if (c, co) == (16, 32) and dt != torch.float32:
                nxt = _empty((n * h2 * w2, co), dt, dev)
                L.conv3x3_c16(e, nxt, wt, bias, 2, L.ACT_LEAKY, False, n, h, w)
                e, h, w, c = nxt, h2, w2, co
                continue
            if (c, co) == (32, 64) and dt != torch.float32:
                nxt = _empty((n * h2 * w2, co), dt, dev)
                L.conv3x3_c32(e, nxt, wt, bias, 2, L.ACT_LEAKY, False, n, h, w)
                e, h, w, c = nxt, h2, w2, co
                continue
            if backend == L.GEMM_AUTO and L.conv3x3_tc_supported(dt, c, co):
                nxt = _empty((n * h2 * w2, co), dt, dev)
                L.conv3x3_tc(e, nxt, wt, bias, 2, L.ACT_LEAKY, False, n, h, w, c, co)
                e, h, w, c = nxt, h2, w2, co
                continue
            a = _empty((n * h2 * w2, 9 * c), dt, dev)
            L.im2col3x3(e, a, n, h, w, c, 2)
            e = _empty((n * h2 * w2, co), dt, dev)
            L.gemm(a, wt, e, n * h2 * w2, co, 9 * c, bias=bias, act=L.ACT_LEAKY, backend=backend)
            h, w, c = h2, w2, co
        return e.view(n, h * w * c)

    def latent(self, feat, eps, mu_out=None, backend=L.GEMM_AUTO):
        """mu GEMM with the reparameterisation fused in the epilogue:
        z = eps * exp(0.5*mu) + mu  (genconvit_vae.py:43-49; std comes from ``mu``, evaluated once).
        ``eps`` is [N,12544] fp32 in the reference's latent order; z is NHWC [N*49, 256]."""
        n = feat.shape[0]
        if feat.shape[1] != 25088:
            raise ValueError(f"VAE latent layer takes 25088 features per frame, got {feat.shape[1]}")
        z = _empty((n * 49, 256), self.dt, self.dev)
        L.gemm(feat, self.mu_w, z, n, 12544, 25088, bias=self.mu_b, eps=eps, eps_c=256, eps_hw=49, mu_out=mu_out,
               ldd=12544, backend=backend)
        return z

    def kl(self, feat, mu):
        """The ``Encoder.kl`` side effect (genconvit_vae.py:56,58); needs the ``var`` GEMM, so it is
        computed only on request.  ``mu``: fp32 [N,12544] (column order is irrelevant for the sum)."""
        if self.var_w is None:
            if self._var_src is None:
                raise L.GcvError("kl requested but the var weights were not kept (with_var=False)")
            self.var_w, self.var_b = self._pack_latent(*self._var_src)
        n = feat.shape[0]
        var = _empty((n, 12544), torch.float32, self.dev)
        L.gemm(feat, self.var_w, var, n, 12544, 25088, bias=self.var_b, out_f32=self.dt != torch.float32)
        return 0.5 * torch.mean(-0.5 * torch.sum(1 + var - mu ** 2 - var.exp(), dim=1), dim=0)

    @_on_device
    def forward(self, x, eps, want_xhat=False, want_kl=False, backend=L.GEMM_AUTO):
        """GenConViTVAE.forward (genconvit_vae.py:107-116) -> (fp32 logits [N,2], x_hat224 | None, kl | None).

        The backbone runs on x at 224x224 and on the 112x112 reconstruction; both passes
        share one set of GEMM launches over the concatenated tokens."""
        dev, dt = self.dev, self.dt
        n, _, hh, ww = x.shape
        if (hh, ww) != (224, 224) or x.shape[1] != 3:
            # the encoder's Linear(128*14*14 -> latent) and the decoder's Unflatten(256,7,7) fix the geometry
            # (genconvit_vae.py:34-37,64); the reference fails in nn.Linear with a shape error for anything else
            raise ValueError(f"GenConViTVAE takes [N,3,224,224] frames, got {tuple(x.shape)}")
        feat = self.encode_features(x, backend)
        mu = _empty((n, 12544), torch.float32, dev) if want_kl else None
        z = self.latent(feat, eps, mu, backend)
        xhat, h2, w2 = _run_convt_stack(z, self.dec, n, 7, 7, L.ACT_LEAKY, dt, dev, backend)
        t1 = (hh // 4, ww // 4)
        t2 = (h2 // 4, w2 // 4)
        m1 = n * t1[0] * t1[1]
        if FUSED_STEM and dt != torch.float32 and backend == L.GEMM_AUTO and self.backbone.dims[0] == 96:
            a0 = [(x, True, n, hh, ww), (xhat, False, n, h2, w2)]
        elif isinstance(x, U8Frames):
            raise ValueError("uint8 frames need the 16-bit fused stem (convnext_tiny, bf16 / fp16)")
        else:
            a0 = _empty((m1 + n * t2[0] * t2[1], 48), dt, dev)
            L.stem_patchify_nchw(x, a0, n, hh, ww)
            L.stem_patchify_nhwc(xhat, a0[m1:], n, h2, w2)
        cat = _empty((n, 2000), dt, dev)
        self.backbone.forward_tokens(a0, [(n, *t1), (n, *t2)], [(0, n, cat, 2000), (n, n, cat[:, 1000:], 2000)],
                                     L.ACT_RELU, backend)
        logits = _heads(cat, n, self.fc_w, self.fc_b, self.fc2_w, self.fc2_b, L.ACT_RELU, dt, dev, backend)
        xhat224 = None
        if want_xhat:
            xhat224 = _empty((n, 3, 2 * h2, 2 * w2), torch.float32, dev)
            L.resize2x_to_nchw(xhat, xhat224, n, h2, w2, 3)
        return logits, xhat224, (self.kl(feat, mu) if want_kl else None)


def score_videos(logits, n_nets, n_frames, frames_per_video):
    """Batched pred_vid scoring (model/pred_func.py:111-131): -> (mean [V,2], cls [V] int32, val [V]) on device."""
    v = n_frames // frames_per_video
    dev = logits.device
    mean = torch.empty((v, 2), dtype=torch.float32, device=dev)
    cls = torch.empty((v,), dtype=torch.int32, device=dev)
    val = torch.empty((v,), dtype=torch.float32, device=dev)
    L.require_cuda_tensor(logits, "score_videos")
    with torch.cuda.device(dev):
        L.score_videos(logits, n_nets, n_frames, frames_per_video, mean, cls, val)
    return mean, cls, val
