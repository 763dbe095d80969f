"""Oracle (test infrastructure): reference state_dict inventory + seeded weights.

``ed_spec()`` / ``vae_spec()`` list every ``state_dict`` entry of the
reference's ``GenConViTED`` / ``GenConViTVAE`` (reference
model/genconvit_ed.py:8-75, model/genconvit_vae.py:11-105,
model/model_embedder.py:8-37) with its shape and role; timm-owned names follow
timm==0.6.5.  The real checkpoints are not available offline (reference
weight/.gitkeep, README.md:116-125), so parity runs on ``make_state_dict``:
a deterministic, per-tensor-seeded "trained-like" randomiser.  timm's default
init (layer-scale gamma = 1e-6) would make the backbone ~identity and hide
kernel bugs, hence random gamma, biases, BN statistics and head scaling
(SURVEY.md section 7 "Weak random-init parity").
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

from .backbones import CONVNEXT, CONVNEXT_TINY, SWIN, SWIN_TINY, swin_attn_mask, swin_relative_position_index

LATENT = 12544          # reference model/config.yaml:4
ENC_FLAT = 128 * 14 * 14  # reference model/genconvit_vae.py:34-37


def _convnext_spec(p, cfg=CONVNEXT_TINY):
    d = cfg["dims"]
    out = [(p + "stem.0.weight", (d[0], 3, 4, 4), "w"), (p + "stem.0.bias", (d[0],), "b"),
           (p + "stem.1.weight", (d[0],), "ln_w"), (p + "stem.1.bias", (d[0],), "b")]
    for s, depth in enumerate(cfg["depths"]):
        q, c = f"{p}stages.{s}.", d[s]
        if s > 0:
            out += [(q + "downsample.0.weight", (d[s - 1],), "ln_w"), (q + "downsample.0.bias", (d[s - 1],), "b"),
                    (q + "downsample.1.weight", (c, d[s - 1], 2, 2), "w"), (q + "downsample.1.bias", (c,), "b")]
        for k in range(depth):
            b = f"{q}blocks.{k}."
            out += [(b + "gamma", (c,), "gamma"),
                    (b + "conv_dw.weight", (c, 1, 7, 7), "w"), (b + "conv_dw.bias", (c,), "b"),
                    (b + "norm.weight", (c,), "ln_w"), (b + "norm.bias", (c,), "b"),
                    (b + "mlp.fc1.weight", (4 * c, c), "w"), (b + "mlp.fc1.bias", (4 * c,), "b"),
                    (b + "mlp.fc2.weight", (c, 4 * c), "w"), (b + "mlp.fc2.bias", (c,), "b")]
    out += [(p + "head.norm.weight", (d[3],), "ln_w"), (p + "head.norm.bias", (d[3],), "b"),
            (p + "head.fc.weight", (1000, d[3]), "w"), (p + "head.fc.bias", (1000,), "b")]
    return out


def _swin_spec(p, cfg=SWIN_TINY):
    ws, e = cfg["window"], cfg["embed"]
    out = [(p + "patch_embed.proj.weight", (e, 3, 4, 4), "w"), (p + "patch_embed.proj.bias", (e,), "b"),
           (p + "patch_embed.norm.weight", (e,), "ln_w"), (p + "patch_embed.norm.bias", (e,), "b")]
    res = cfg["img"] // cfg["patch"]
    for l, depth in enumerate(cfg["depths"]):
        c, h = e * 2 ** l, cfg["heads"][l]
        for k in range(depth):
            b = f"{p}layers.{l}.blocks.{k}."
            shift = 0 if (k % 2 == 0 or res <= ws) else ws // 2
            if shift:
                out.append((b + "attn_mask", ((res // ws) ** 2, ws * ws, ws * ws), f"attn_mask:{res}"))
            out += [(b + "norm1.weight", (c,), "ln_w"), (b + "norm1.bias", (c,), "b"),
                    (b + "attn.relative_position_bias_table", ((2 * ws - 1) ** 2, h), "rpb"),
                    (b + "attn.relative_position_index", (ws * ws, ws * ws), "rpi"),
                    (b + "attn.qkv.weight", (3 * c, c), "w"), (b + "attn.qkv.bias", (3 * c,), "b"),
                    (b + "attn.proj.weight", (c, c), "w"), (b + "attn.proj.bias", (c,), "b"),
                    (b + "norm2.weight", (c,), "ln_w"), (b + "norm2.bias", (c,), "b"),
                    (b + "mlp.fc1.weight", (4 * c, c), "w"), (b + "mlp.fc1.bias", (4 * c,), "b"),
                    (b + "mlp.fc2.weight", (c, 4 * c), "w"), (b + "mlp.fc2.bias", (c,), "b")]
        if l < len(cfg["depths"]) - 1:
            q = f"{p}layers.{l}.downsample."
            out += [(q + "reduction.weight", (2 * c, 4 * c), "w"),
                    (q + "norm.weight", (4 * c,), "ln_w"), (q + "norm.bias", (4 * c,), "b")]
            res //= 2
    c = e * 8
    out += [(p + "norm.weight", (c,), "ln_w"), (p + "norm.bias", (c,), "b"),
            (p + "head.weight", (1000, c), "w"), (p + "head.bias", (1000,), "b")]
    return out


def _hybrid_spec(backbone_prefix, swin_cfg=SWIN_TINY):
    """HybridEmbed attached as <backbone>.patch_embed (reference model_embedder.py:16-37):
    proj = Conv2d(1000, 768, 1) because the probe output of the Swin classifier is [1,1000]."""
    p = backbone_prefix + "patch_embed."
    return ([(p + "proj.weight", (768, 1000, 1, 1), "w"), (p + "proj.bias", (768,), "b")]
            + [(n, s, "alias:" + n.replace(p + "backbone.", "embedder.", 1)) for n, s, _ in _swin_spec(p + "backbone.", swin_cfg)])


def _variant(size):
    """'tiny' | 'large' -> (ConvNeXt cfg, Swin cfg), reference prediction.py:314-318."""
    return CONVNEXT[f"convnext_{size}"], SWIN[f"swin_{size}_patch4_window7_224"]


def ed_spec(size="tiny"):
    """reference model/genconvit_ed.py: Encoder 13-33, Decoder 43-58, GenConViTED 66-75."""
    cn, sw = _variant(size)
    out = []
    for i, (ci, co) in zip((0, 3, 6, 9, 12), ((3, 16), (16, 32), (32, 64), (64, 128), (128, 256))):
        out += [(f"encoder.features.{i}.weight", (co, ci, 3, 3), "w_relu"), (f"encoder.features.{i}.bias", (co,), "b")]
    for i, (ci, co) in zip((0, 2, 4, 6, 8), ((256, 128), (128, 64), (64, 32), (32, 16), (16, 3))):
        out += [(f"decoder.features.{i}.weight", (ci, co, 2, 2), "w_convt"), (f"decoder.features.{i}.bias", (co,), "b")]
    out += _convnext_spec("backbone.", cn) + _hybrid_spec("backbone.", sw) + _swin_spec("embedder.", sw)
    out += [("fc.weight", (500, 2000), "w_head1"), ("fc.bias", (500,), "b"),
            ("fc2.weight", (2, 500), "w_head2"), ("fc2.bias", (2,), "b_head2")]
    return out


def vae_spec(latent=LATENT, size="tiny"):
    """reference model/genconvit_vae.py: Encoder 15-37, Decoder 67-83, GenConViTVAE 93-105."""
    cn, sw = _variant(size)
    out = []
    for i, (ci, co) in zip((0, 3, 6, 9), ((3, 16), (16, 32), (32, 64), (64, 128))):
        out += [(f"encoder.features.{i}.weight", (co, ci, 3, 3), "w_relu"), (f"encoder.features.{i}.bias", (co,), "b"),
                (f"encoder.features.{i + 1}.weight", (co,), "ln_w"), (f"encoder.features.{i + 1}.bias", (co,), "b"),
                (f"encoder.features.{i + 1}.running_mean", (co,), "bn_mean"),
                (f"encoder.features.{i + 1}.running_var", (co,), "bn_var"),
                (f"encoder.features.{i + 1}.num_batches_tracked", (), "bn_nbt")]
    out += [("encoder.fc1.weight", (256, ENC_FLAT), "w"), ("encoder.fc1.bias", (256,), "b"),
            ("encoder.fc2.weight", (128, 256), "w"), ("encoder.fc2.bias", (128,), "b"),
            ("encoder.mu.weight", (latent, ENC_FLAT), "w_mu"), ("encoder.mu.bias", (latent,), "b"),
            ("encoder.var.weight", (latent, ENC_FLAT), "w_mu"), ("encoder.var.bias", (latent,), "b")]
    for i, (ci, co) in zip((0, 2, 4, 6), ((256, 64), (64, 32), (32, 16), (16, 3))):
        out += [(f"decoder.features.{i}.weight", (ci, co, 2, 2), "w_convt"), (f"decoder.features.{i}.bias", (co,), "b")]
    out += _swin_spec("embedder.", sw) + _convnext_spec("convnext_backbone.", cn) + _hybrid_spec("convnext_backbone.", sw)
    out += [("fc.weight", (500, 2000), "w_head1"), ("fc.bias", (500,), "b"),
            ("fc3.weight", (500, 1000), "w"), ("fc3.bias", (500,), "b"),
            ("fc2.weight", (2, 500), "w_head2"), ("fc2.bias", (2,), "b_head2")]
    return out


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def _fan_in(shape, kind):
    if kind == "w_convt":            # ConvTranspose2d weight [Cin, Cout, kh, kw]; k2s2 -> one tap per output pixel
        return shape[0]
    return int(math.prod(shape[1:])) if len(shape) > 1 else 1


def make_tensor(name, shape, kind, seed=0, skip_big=False):
    """One deterministic tensor.  Values depend only on (name, seed), never on generation order."""
    g = _gen(name, seed)
    if kind.startswith("attn_mask"):
        res = int(kind.split(":")[1])
        return swin_attn_mask(res, SWIN_TINY["window"], SWIN_TINY["window"] // 2)
    if kind == "rpi":
        return swin_relative_position_index(SWIN_TINY["window"])
    if kind == "bn_nbt":
        return torch.tensor(100, dtype=torch.int64)
    if kind in ("w", "w_relu", "w_convt", "w_mu", "w_head1", "w_head2"):
        if skip_big and math.prod(shape) > 50_000_000:
            return torch.zeros(shape)
        gain = {"w": 1.0, "w_relu": 2.0 ** 0.5, "w_convt": 2.0 ** 0.5, "w_mu": 0.6,
                "w_head1": 0.5, "w_head2": 3.0}[kind]
        std = gain / math.sqrt(_fan_in(shape, kind))
        return torch.empty(shape).normal_(0.0, std, generator=g)
    if kind == "b":
        return torch.empty(shape).normal_(0.0, 0.02, generator=g)
    if kind == "b_head2":
        return torch.empty(shape).normal_(0.0, 0.1, generator=g)
    if kind == "ln_w":
        return torch.empty(shape).uniform_(0.5, 1.5, generator=g)
    if kind == "gamma":
        return torch.empty(shape).uniform_(0.0, 0.5, generator=g)
    if kind == "bn_mean":
        return torch.empty(shape).normal_(0.0, 0.1, generator=g)
    if kind == "bn_var":
        return torch.empty(shape).uniform_(0.5, 1.5, generator=g)
    if kind == "rpb":
        return torch.empty(shape).normal_(0.0, 0.02, generator=g)
    raise ValueError(kind)


def make_state_dict(net: str, seed: int = 0, latent: int = LATENT, skip_var: bool = False, size: str = "tiny"):
    """Seeded state_dict with the reference layout.  ``net`` in {'ed','vae'}.

    Swin tensors appear under both ``embedder.*`` and
    ``<backbone>.patch_embed.backbone.*`` and share storage, as in the reference
    (same module object, reference genconvit_ed.py:69-70).  ``skip_var`` zero-fills
    ``encoder.var.weight`` (a 1.2 GiB tensor that only feeds the ``kl`` side effect).
    """
    spec = ed_spec(size) if net == "ed" else vae_spec(latent, size)
    sd = OrderedDict()
    aliases = []
    for name, shape, kind in spec:
        if kind.startswith("alias:"):
            aliases.append((name, kind[6:]))
            continue
        big_skip = skip_var and name == "encoder.var.weight"
        sd[name] = make_tensor(name, shape, kind, seed, skip_big=big_skip)
    for name, target in aliases:
        sd[name] = sd[target]
    return OrderedDict((n, sd[n]) for n, _, _ in spec)


def synthetic_frames(n: int, seed: int = 0, size: int = 224) -> torch.Tensor:
    """Normalised-image-like frames: N(0,1) clamped to the ImageNet-normalised range."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(n, 3, size, size, generator=g).clamp_(-2.1179, 2.64)


def synthetic_eps(n: int, seed: int = 7, latent: int = LATENT) -> torch.Tensor:
    """Injected VAE epsilon in the reference's (NCHW-flatten) latent order."""
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(n, latent, generator=g)
