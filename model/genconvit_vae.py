"""Network B -- drop-in for reference model/genconvit_vae.py (Encoder :11-60, Decoder :62-88,
GenConViTVAE :90-116).  Parameter containers with the reference ``state_dict`` layout
(including the never-used ``encoder.fc1/fc2`` and ``fc3``); the forward runs on the
sm_100a kernels via genconvit_b200.engine.
"""
import torch
import torch.nn as nn

from genconvit_b200 import engine, lib as L
from genconvit_b200.modules import compute_dtype_of, create_model, weights_fingerprint

from .config import load_config
from .genconvit_ed import _Packable, _chain
from .model_embedder import HybridEmbed

config = load_config()      # the reference also reads the config at import (genconvit_vae.py:8)


class Encoder(_Packable):
    """4 x (Conv3x3 s2 p1 -> BatchNorm2d -> LeakyReLU), mu/var Linear(25088 -> latent)."""

    def __init__(self, latent_dims=4):
        super().__init__()
        self.features = _chain((3, 16, 32, 64, 128),
                               lambda ci, co: [nn.Conv2d(ci, co, 3, 2, 1), nn.BatchNorm2d(co), nn.LeakyReLU()])
        self.latent_dims = latent_dims
        self.fc1 = nn.Linear(128 * 14 * 14, 256)          # unused by the reference forward; kept for the state_dict
        self.fc2 = nn.Linear(256, 128)
        self.mu = nn.Linear(128 * 14 * 14, latent_dims)
        self.var = nn.Linear(128 * 14 * 14, latent_dims)
        self.kl = 0
        self.kl_weight = 0.5
        self.relu = nn.LeakyReLU()


class Decoder(_Packable):
    """Unflatten(256,7,7) + 4 x (ConvTranspose k2 s2 -> LeakyReLU): 7 -> 112."""

    def __init__(self, latent_dims=4):
        super().__init__()
        self.features = _chain((256, 64, 32, 16, 3),
                               lambda ci, co: [nn.ConvTranspose2d(ci, co, 2, 2), nn.LeakyReLU()])
        self.latent_dims = latent_dims
        self.unflatten = nn.Unflatten(dim=1, unflattened_size=(256, 7, 7))


class GenConViTVAE(_Packable):
    def __init__(self, config, pretrained=True):
        super().__init__()
        self.latent_dims = config["model"]["latent_dims"]
        if self.latent_dims != 12544:
            raise NotImplementedError("the decoder unflattens to (256,7,7): latent_dims must be 12544")
        self.encoder = Encoder(self.latent_dims)
        self.decoder = Decoder(self.latent_dims)
        self.embedder = create_model(config["model"]["embedder"], pretrained=True)
        self.convnext_backbone = create_model(config["model"]["backbone"], pretrained=True, num_classes=1000,
                                              drop_path_rate=0, head_init_scale=1.0)
        self.convnext_backbone.patch_embed = HybridEmbed(self.embedder, img_size=config["img_size"], embed_dim=768)
        self.num_feature = self.convnext_backbone.head.fc.out_features * 2
        self.fc = nn.Linear(self.num_feature, self.num_feature // 4)
        self.fc3 = nn.Linear(self.num_feature // 2, self.num_feature // 4)   # unused by the reference forward
        self.fc2 = nn.Linear(self.num_feature // 4, config["num_classes"])
        self.relu = nn.ReLU()
        self.compute_kl = False     # the reference's encoder.kl side effect needs a second 25088x12544 GEMM
        self._eps = None

    def set_epsilon(self, eps):
        """Inject the reparameterisation noise ([N,12544], reference latent order) for the next
        forward calls; ``None`` restores the reference behaviour (fresh randn every call,
        even in eval mode -- genconvit_vae.py:46)."""
        self._eps = eps
        return self

    def _engine(self, device):
        dt = compute_dtype_of(self, self.compute_dtype)
        fp = (weights_fingerprint(self), bool(self.compute_kl))      # toggling compute_kl re-packs (the var weights)
        if self._packed is None or self._packed.dt != dt or self._packed.dev != device or self._packed_fp != fp:
            self._packed, self._packed_fp = engine.PackedVAE(self.state_dict(), device, dt, with_var=self.compute_kl), fp
        return self._packed

    def _logits_f32(self, x, eps=None):
        """fp32 contiguous NCHW frames (+ fp32 eps [N,latent] or None) on the GPU -> the engine's fp32 logits [N,2]."""
        eps = eps if eps is not None else self._eps
        if eps is None:
            eps = torch.randn(x.shape[0], self.latent_dims, device=x.device, dtype=torch.float32)
        return self._engine(x.device).forward(x, eps, want_xhat=False, want_kl=False)[0]

    def _forward(self, x, eps=None, want_xhat=True):
        L.require_cuda_tensor(x, "GenConViTVAE.forward")
        x = x.float().contiguous()
        eps = eps if eps is not None else self._eps
        if eps is None:
            eps = torch.randn(x.shape[0], self.latent_dims, device=x.device, dtype=torch.float32)
        eps = eps.to(device=x.device, dtype=torch.float32).contiguous()
        if eps.shape != (x.shape[0], self.latent_dims):
            raise ValueError(f"eps must be [{x.shape[0]}, {self.latent_dims}], got {tuple(eps.shape)}")
        logits, xhat, kl = self._engine(x.device).forward(x, eps, want_xhat=want_xhat, want_kl=self.compute_kl)
        if kl is not None:
            self.encoder.kl = kl
        pd = next(self.parameters()).dtype
        return logits.to(pd), (xhat.to(pd) if xhat is not None else None)

    def forward(self, x, eps=None):
        """[N,3,224,224] -> (logits [N,2], x_hat resized to [N,3,224,224])."""
        return self._forward(x, eps, want_xhat=True)
