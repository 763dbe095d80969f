"""Network A -- drop-in for reference model/genconvit_ed.py (Encoder :8-36, Decoder :38-61,
GenConViTED :63-89).  The nn layers below only own the parameters (same ``state_dict``
keys as the reference); every forward runs on the sm_100a kernels via genconvit_b200.engine.
"""
import torch
import torch.nn as nn

from genconvit_b200 import engine, lib as L
from genconvit_b200.modules import compute_dtype_of, create_model, weights_fingerprint

from .model_embedder import HybridEmbed


class _Packable(nn.Module):
    """Invalidate the kernel-layout weight copies whenever parameters may have changed: re-allocation (``.to()``,
    ``.half()``) through ``_apply``, in-place writes (``load_state_dict`` on this module, a PARENT or a child, direct
    ``copy_``) through ``weights_fingerprint`` checked on every forward."""
    _packed = None
    _packed_fp = None
    compute_dtype = None        # None: follow the parameter dtype; or 'fp32' | 'bf16' | 'fp16'

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def set_compute_dtype(self, dt):
        self.compute_dtype, self._packed = dt, None
        return self


def _chain(channels, make):
    layers = []
    for ci, co in zip(channels[:-1], channels[1:]):
        layers += make(ci, co)
    return nn.Sequential(*layers)


class Encoder(_Packable):
    """5 x (Conv3x3 s1 p1 -> ReLU -> MaxPool2): 224 -> 7, 3 -> 256 channels."""

    def __init__(self):
        super().__init__()
        self.features = _chain((3, 16, 32, 64, 128, 256),
                               lambda ci, co: [nn.Conv2d(ci, co, 3, 1, 1), nn.ReLU(inplace=True), nn.MaxPool2d(2, 2)])


class Decoder(_Packable):
    """5 x (ConvTranspose k2 s2 -> ReLU): 7 -> 224, 256 -> 3 channels."""

    def __init__(self):
        super().__init__()
        self.features = _chain((256, 128, 64, 32, 16, 3),
                               lambda ci, co: [nn.ConvTranspose2d(ci, co, 2, 2), nn.ReLU(inplace=True)])


class GenConViTED(_Packable):
    def __init__(self, config, pretrained=True):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.backbone = create_model(config["model"]["backbone"], pretrained=pretrained)
        self.embedder = create_model(config["model"]["embedder"], pretrained=pretrained)
        self.backbone.patch_embed = HybridEmbed(self.embedder, img_size=config["img_size"], embed_dim=768)
        self.num_features = self.backbone.head.fc.out_features * 2
        self.fc = nn.Linear(self.num_features, self.num_features // 4)
        self.fc2 = nn.Linear(self.num_features // 4, 2)
        self.relu = nn.GELU()       # the reference's attribute really is a GELU (genconvit_ed.py:75)

    def _engine(self, device):
        dt = compute_dtype_of(self, self.compute_dtype)
        fp = weights_fingerprint(self)
        if self._packed is None or self._packed.dt != dt or self._packed.dev != device or self._packed_fp != fp:
            self._packed, self._packed_fp = engine.PackedED(self.state_dict(), device, dt), fp
        return self._packed

    def _logits_f32(self, x):
        """fp32 contiguous NCHW frames on the GPU -> the engine's fp32 logits [N,2] (no dtype round trip)."""
        return self._engine(x.device).forward(x)

    def forward(self, images):
        """[N,3,224,224] -> logits [N,2] (dtype of the parameters, like the reference)."""
        L.require_cuda_tensor(images, "GenConViTED.forward")
        x = images.float().contiguous()
        logits = self._engine(x.device).forward(x)
        return logits.to(next(self.parameters()).dtype)
