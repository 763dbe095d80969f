"""Parameter containers with timm==0.6.5 module names for ConvNeXt-T and Swin-T.

The reference builds these through ``timm.create_model`` (model/genconvit_ed.py:68-69,
model/genconvit_vae.py:96-97).  Here they only own parameters/buffers under the same
``state_dict`` keys; the arithmetic runs in the sm_100a kernels (genconvit_b200.engine).
No timm, no pretrained download.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from . import engine
from . import lib as L

DEPTHS = engine.DEPTHS            # convnext_tiny (the reference's default, model/config.yaml:2)
DIMS = engine.DIMS
# timm model name -> (dims, depths) / (embed, depths, heads): the tiny defaults and the '--s large' variants of
# reference prediction.py:314-318
CONVNEXT_VARIANTS = {"convnext_tiny": ((96, 192, 384, 768), (3, 3, 9, 3)),
                     "convnext_large": ((192, 384, 768, 1536), (3, 3, 27, 3))}
SWIN_VARIANTS = {"swin_tiny_patch4_window7_224": (96, (2, 2, 6, 2), (3, 6, 12, 24)),
                 "swin_large_patch4_window7_224": (192, (2, 2, 18, 2), (6, 12, 24, 48))}


def weights_fingerprint(module):
    """Cheap change detector for the kernel-layout weight copies: every in-place write to a parameter or buffer
    (``load_state_dict`` on this module, a parent or a child, ``copy_``, optimiser steps) bumps its ``_version``;
    re-allocation (``.to()``, ``.half()``) goes through ``_apply``, which drops the copies anyway."""
    v = 0
    n = 0
    for t in module.parameters():
        v += t._version
        n += 1
    for t in module.buffers():
        v += t._version
        n += 1
    return (n, v)


def compute_dtype_of(module: nn.Module, override=None) -> torch.dtype:
    """fp32 parameters -> fp32 kernels (exact mode); .half() -> fp16; .bfloat16() -> bf16.
    ``override`` ('fp32' | 'bf16' | 'fp16' | torch dtype) selects the kernel precision
    independently of how the parameters are stored."""
    if override is not None:
        if isinstance(override, str):
            return {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16,
                    "bfloat16": torch.bfloat16, "fp16": torch.float16, "float16": torch.float16,
                    "half": torch.float16}[override]
        return override
    p = next(module.parameters())
    return p.dtype if p.dtype in (torch.float16, torch.bfloat16) else torch.float32


class _Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c, 4 * c)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(4 * c, c)


class _ConvNeXtBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_dw = nn.Conv2d(c, c, 7, padding=3, groups=c)
        self.norm = nn.LayerNorm(c, eps=1e-6)
        self.mlp = _Mlp(c)
        self.gamma = nn.Parameter(torch.full((c,), 1e-6))


class _ConvNeXtStage(nn.Module):
    def __init__(self, cin, cout, depth):
        super().__init__()
        if cin != cout:
            self.downsample = nn.Sequential(nn.LayerNorm(cin, eps=1e-6), nn.Conv2d(cin, cout, 2, stride=2))
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.Sequential(*[_ConvNeXtBlock(cout) for _ in range(depth)])


class ConvNeXt(nn.Module):
    """convnext_tiny: forward = head(norm_pre(stages(stem(x)))).  A ``patch_embed`` attribute
    attached by the caller (reference genconvit_ed.py:70) is carried in the state_dict and never read."""

    def __init__(self, dims=DIMS, depths=DEPTHS):
        super().__init__()
        DIMS, DEPTHS = tuple(dims), tuple(depths)
        self.dims, self.depths = DIMS, DEPTHS
        self.stem = nn.Sequential(nn.Conv2d(3, DIMS[0], 4, stride=4), nn.LayerNorm(DIMS[0], eps=1e-6))
        stages, cin = [], DIMS[0]
        for c, d in zip(DIMS, DEPTHS):
            stages.append(_ConvNeXtStage(cin, c, d))
            cin = c
        self.stages = nn.Sequential(*stages)
        self.norm_pre = nn.Identity()
        self.head = nn.Sequential(OrderedDict([
            ("global_pool", nn.AdaptiveAvgPool2d(1)), ("norm", nn.LayerNorm(DIMS[3], eps=1e-6)),
            ("flatten", nn.Flatten(1)), ("drop", nn.Dropout(0.0)), ("fc", nn.Linear(DIMS[3], 1000))]))
        self.num_features = DIMS[3]
        self._packed = None
        self.compute_dtype = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def forward(self, x):
        """fp32 NCHW frames -> fp32 [N,1000] logits, on the CUDA kernels."""
        L.require_cuda_tensor(x, "ConvNeXt.forward")
        dt = compute_dtype_of(self, self.compute_dtype)
        fp = weights_fingerprint(self)
        if self._packed is None or self._packed.dt != dt or self._packed.dev != x.device or self._packed_fp != fp:
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith("patch_embed.")}
            self._packed, self._packed_fp = engine.PackedConvNeXt(sd, x.device, dt), fp
        return self._packed.forward_images(x.float().contiguous())


# ---- Swin-T (parameters only; see SURVEY.md section 0: never executed by the reference forward) ----
class _WindowAttention(nn.Module):
    def __init__(self, c, heads, ws=7):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(c, 3 * c)
        self.proj = nn.Linear(c, c)


def _shift_mask(res, ws, shift):
    img = torch.zeros(res, res)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(res // ws, ws, res // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    m = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(m != 0, torch.full_like(m, -100.0), torch.zeros_like(m))


class _SwinBlock(nn.Module):
    def __init__(self, c, heads, res, shift, ws=7):
        super().__init__()
        self.norm1 = nn.LayerNorm(c)
        self.attn = _WindowAttention(c, heads, ws)
        self.norm2 = nn.LayerNorm(c)
        self.mlp = _Mlp(c)
        if shift > 0:
            self.register_buffer("attn_mask", _shift_mask(res, ws, shift))
        else:
            self.attn_mask = None


class _PatchMerging(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.reduction = nn.Linear(4 * c, 2 * c, bias=False)
        self.norm = nn.LayerNorm(4 * c)


class _SwinLayer(nn.Module):
    def __init__(self, c, depth, heads, res, downsample, ws=7):
        super().__init__()
        self.blocks = nn.ModuleList(
            [_SwinBlock(c, heads, res, 0 if (k % 2 == 0 or res <= ws) else ws // 2, ws) for k in range(depth)])
        if downsample:
            self.downsample = _PatchMerging(c)


class _PatchEmbed(nn.Module):
    def __init__(self, embed=96):
        super().__init__()
        self.proj = nn.Conv2d(3, embed, 4, stride=4)
        self.norm = nn.LayerNorm(embed)


class SwinTransformer(nn.Module):
    """swin_tiny_patch4_window7_224 (28.3 M parameters + index/mask buffers) with timm's state_dict keys.

    The reference constructs it as ``embedder`` and wraps it in ``HybridEmbed``, but its GenConViT
    forward never reaches either, so no logit depends on it.  ``embedder(x)``
    is a standalone callable -> [N,1000] running on the CUDA kernels (engine.PackedSwin: tcgen05 GEMMs,
    row LayerNorm, the window-attention / patch-merge kernels of csrc/swin_ops.cu).  The
    ``relative_position_index`` / ``attn_mask`` buffers are kept for the state_dict; the kernels
    recompute both from (window, shift, resolution), which is what defines them.
    """

    def __init__(self, embed=96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24)):
        super().__init__()
        self.embed, self.depths, self.heads = embed, tuple(depths), tuple(heads)
        self.patch_embed = _PatchEmbed(embed)
        layers, res = [], 56
        for l, (d, h) in enumerate(zip(depths, heads)):
            layers.append(_SwinLayer(embed * 2 ** l, d, h, res, downsample=l < 3))
            if l < 3:
                res //= 2
        self.layers = nn.Sequential(*layers)
        self.norm = nn.LayerNorm(8 * embed)
        self.head = nn.Linear(8 * embed, 1000)
        self.num_features = 8 * embed
        self._packed = None
        self.compute_dtype = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def forward(self, x):
        """fp32 NCHW 224x224 frames -> fp32 [N,1000] logits, on the CUDA kernels."""
        L.require_cuda_tensor(x, "SwinTransformer.forward")
        dt = compute_dtype_of(self, self.compute_dtype)
        fp = weights_fingerprint(self)
        if self._packed is None or self._packed.dt != dt or self._packed.dev != x.device or self._packed_fp != fp:
            self._packed, self._packed_fp = engine.PackedSwin(self.state_dict(), x.device, dt), fp
        return self._packed.forward_images(x.float().contiguous())


def create_model(name, pretrained=False, num_classes=1000, drop_path_rate=0.0, head_init_scale=1.0, **_):
    """Drop-in for the two ``timm.create_model`` calls of the reference; never downloads weights.  Builds the models the
    reference can select: the tiny defaults (model/config.yaml:2-3) and the ``--s large`` pair (prediction.py:314-318)."""
    if num_classes != 1000:
        raise NotImplementedError("only the 1000-class heads the reference uses are supported")
    if name in CONVNEXT_VARIANTS:
        return ConvNeXt(*CONVNEXT_VARIANTS[name])
    if name in SWIN_VARIANTS:
        return SwinTransformer(*SWIN_VARIANTS[name])
    raise NotImplementedError(
        f"backbone {name!r}: the reference only ever selects convnext_{{tiny,large}} / "
        "swin_{tiny,large}_patch4_window7_224 (model/config.yaml:2-3, prediction.py:314-318)")
