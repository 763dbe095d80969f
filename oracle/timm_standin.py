"""Oracle (test infrastructure): a ``timm`` stand-in so the UNMODIFIED reference imports.

The reference does ``timm.create_model('convnext_tiny' | 'swin_tiny_patch4_window7_224', ...)``
(reference model/genconvit_ed.py:4-5,68-69; model/genconvit_vae.py:4,96-97).
timm==0.6.5 is not installable offline, so ``install()`` registers this module as
``timm`` in ``sys.modules``.  The shells below own parameters/buffers under
timm-0.6.5 module names (so the reference ``state_dict`` layout is reproduced
key-for-key) and run the functional arithmetic in ``oracle.backbones``.

Used only by ``oracle/make_golden.py`` (authoring container) and the CPU tests.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

from . import backbones
from .weights import _convnext_spec, _swin_spec, make_tensor


class _Shell(nn.Module):
    """Generic container: nested children created on demand from dotted names."""

    def put(self, dotted, tensor, buffer=False):
        mod, parts = self, dotted.split(".")
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, _Shell())
            mod = mod._modules[part]
        if buffer:
            mod.register_buffer(parts[-1], tensor)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(tensor))


class _Net(_Shell):
    _spec = staticmethod(lambda p, cfg: [])
    _fwd = None
    _cfg = None

    def __init__(self):
        super().__init__()
        for name, shape, kind in self._spec("", self._cfg):
            t = make_tensor(name, shape, kind, seed=12345)
            self.put(name, t, buffer=kind == "rpi" or kind.startswith("attn_mask"))

    def forward(self, x):
        sd = self.state_dict(keep_vars=True)
        return type(self)._fwd(sd, "", x)


class ConvNeXt(_Net):
    """forward = head(norm_pre(stages(stem(x)))); a ``patch_embed`` attribute attached later is never read."""
    _spec = staticmethod(_convnext_spec)
    _fwd = staticmethod(backbones.convnext_forward)

    _cfg = backbones.CONVNEXT_TINY

    def __init__(self):
        super().__init__()
        self.head.fc.out_features = 1000   # read at reference genconvit_ed.py:72 / genconvit_vae.py:99
        self.num_features = self._cfg["dims"][3]


class SwinTransformer(_Net):
    _spec = staticmethod(_swin_spec)
    _fwd = staticmethod(backbones.swin_forward)
    _cfg = backbones.SWIN_TINY

    def __init__(self):
        super().__init__()
        self.num_features = self._cfg["embed"] * 8


class ConvNeXtLarge(ConvNeXt):
    _cfg = backbones.CONVNEXT_LARGE


class SwinLarge(SwinTransformer):
    _cfg = backbones.SWIN_LARGE


_MODELS = {"convnext_tiny": ConvNeXt, "swin_tiny_patch4_window7_224": SwinTransformer,
           "convnext_large": ConvNeXtLarge, "swin_large_patch4_window7_224": SwinLarge}


def create_model(name, pretrained=False, num_classes=1000, drop_path_rate=0.0, head_init_scale=1.0, **kw):
    """Same call signature the reference uses; ``pretrained`` is ignored (no network)."""
    if name not in _MODELS:
        raise RuntimeError(f"timm stand-in has no model {name!r}")
    if num_classes != 1000:
        raise RuntimeError("timm stand-in only builds the 1000-class heads the reference uses")
    return _MODELS[name]()


def install():
    """Register stand-ins for timm and the four host-only libraries model/pred_func.py imports."""
    timm = types.ModuleType("timm")
    timm.create_model = create_model
    timm.__version__ = "0.6.5-standin"
    sys.modules["timm"] = timm

    dlib = types.ModuleType("dlib")
    dlib.DLIB_USE_CUDA = False
    sys.modules.setdefault("dlib", dlib)
    sys.modules.setdefault("face_recognition", types.ModuleType("face_recognition"))
    decord = types.ModuleType("decord")
    decord.VideoReader = object
    decord.cpu = lambda *_a, **_k: None
    sys.modules.setdefault("decord", decord)
    alb = types.ModuleType("albumentations")
    for n in ("HorizontalFlip VerticalFlip ShiftScaleRotate CLAHE RandomRotate90 Transpose HueSaturationValue "
              "GaussNoise Sharpen Emboss RandomBrightnessContrast OneOf Compose").split():
        setattr(alb, n, object)
    sys.modules.setdefault("albumentations", alb)
