"""Oracle (test infrastructure): GenConViT ED / VAE forward and pred_vid scoring.

Functional fp32 torch-CPU restatement of the reference hot path, operating on a
``state_dict`` with the reference key layout.  Every function cites the
reference lines it follows.  Swin / HybridEmbed parameters are ignored here
exactly as the reference's forward ignores them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .backbones import convnext_forward, swin_forward


# ---- Network A (reference model/genconvit_ed.py) --------------------------
def ed_encoder(sd, x):
    """5 x (conv3x3 s1 p1 -> ReLU -> maxpool2): [N,3,224,224] -> [N,256,7,7]  (genconvit_ed.py:13-36)."""
    for i in (0, 3, 6, 9, 12):
        x = F.conv2d(x, sd[f"encoder.features.{i}.weight"], sd[f"encoder.features.{i}.bias"], padding=1)
        x = F.max_pool2d(F.relu(x), 2)
    return x


def ed_decoder(sd, x):
    """5 x (convT k2 s2 -> ReLU): [N,256,7,7] -> [N,3,224,224]  (genconvit_ed.py:43-61)."""
    for i in (0, 2, 4, 6, 8):
        x = F.relu(F.conv_transpose2d(x, sd[f"decoder.features.{i}.weight"], sd[f"decoder.features.{i}.bias"], stride=2))
    return x


def ed_forward(sd, images, taps=None):
    """GenConViTED.forward (genconvit_ed.py:77-89): logits [N,2].

    cat order is (backbone(decoded), backbone(images)); the attribute named
    ``relu`` is GELU and is applied to the 2000-d ImageNet logits before ``fc``.
    """
    dec = ed_decoder(sd, ed_encoder(sd, images))
    x1 = convnext_forward(sd, "backbone.", dec)
    x2 = convnext_forward(sd, "backbone.", images)
    x = torch.cat((x1, x2), dim=1)
    if taps is not None:
        taps.update(decoded=dec, x1=x1, x2=x2)
    x = F.linear(F.gelu(x), sd["fc.weight"], sd["fc.bias"])
    return F.linear(F.gelu(x), sd["fc2.weight"], sd["fc2.bias"])


# ---- Network B (reference model/genconvit_vae.py) -------------------------
def vae_encoder_features(sd, x):
    """4 x (conv3x3 s2 p1 -> BatchNorm(eval, eps 1e-5) -> LeakyReLU(0.01)) -> flatten NCHW  (genconvit_vae.py:15-31,52-53)."""
    for i in (0, 3, 6, 9):
        x = F.conv2d(x, sd[f"encoder.features.{i}.weight"], sd[f"encoder.features.{i}.bias"], stride=2, padding=1)
        b = f"encoder.features.{i + 1}."
        x = F.batch_norm(x, sd[b + "running_mean"], sd[b + "running_var"], sd[b + "weight"], sd[b + "bias"], False, 0.1, 1e-5)
        x = F.leaky_relu(x, 0.01)
    return torch.flatten(x, 1)


def vae_latent(sd, feat, eps):
    """Encoder.forward/reparameterize (genconvit_vae.py:43-60).

    z = eps * exp(0.5 * mu(x)) + mu(x): the standard deviation is taken from
    ``mu`` (not ``var``); the reference evaluates mu three times, once is enough.
    Returns (z, mu).
    """
    mu = F.linear(feat, sd["encoder.mu.weight"], sd["encoder.mu.bias"])
    return eps * torch.exp(0.5 * mu) + mu, mu


def vae_kl(sd, feat, mu):
    """``Encoder.kl`` side effect (genconvit_vae.py:56,58); not part of the logits."""
    var = F.linear(feat, sd["encoder.var.weight"], sd["encoder.var.bias"])
    return 0.5 * torch.mean(-0.5 * torch.sum(1 + var - mu ** 2 - var.exp(), dim=1), dim=0)


def vae_decoder(sd, z):
    """unflatten(256,7,7) + 4 x (convT k2 s2 -> LeakyReLU): -> [N,3,112,112]  (genconvit_vae.py:67-88)."""
    x = z.unflatten(1, (256, 7, 7))
    for i in (0, 2, 4, 6):
        x = F.leaky_relu(F.conv_transpose2d(x, sd[f"decoder.features.{i}.weight"], sd[f"decoder.features.{i}.bias"], stride=2), 0.01)
    return x


def vae_forward(sd, x, eps, taps=None, resize=True):
    """GenConViTVAE.forward (genconvit_vae.py:107-116): (logits [N,2], x_hat resized to 224).

    The backbone sees the 112x112 reconstruction; the 224 resize (antialiased
    bilinear; a pure upscale, so antialiasing is a no-op) only touches the
    returned image.  cat order is (backbone(x), backbone(x_hat)); head act is ReLU.
    """
    feat = vae_encoder_features(sd, x)
    z, mu = vae_latent(sd, feat, eps)
    x_hat = vae_decoder(sd, z)
    x1 = convnext_forward(sd, "convnext_backbone.", x)
    x2 = convnext_forward(sd, "convnext_backbone.", x_hat)
    h = torch.cat((x1, x2), dim=1)
    if taps is not None:
        taps.update(feat=feat, mu=mu, z=z, x_hat=x_hat, x1=x1, x2=x2)
    h = F.linear(F.relu(h), sd["fc.weight"], sd["fc.bias"])
    logits = F.linear(F.relu(h), sd["fc2.weight"], sd["fc2.bias"])
    if not resize:
        return logits, x_hat
    return logits, F.interpolate(x_hat, size=(224, 224), mode="bilinear", align_corners=False, antialias=True)


# ---- wrapper + scoring ------------------------------------------------------
def genconvit_forward(sd_ed, sd_vae, x, eps, net="genconvit"):
    """GenConViT.forward (model/genconvit.py:66-75): 'ed' -> [N,2]; 'vae' -> [N,2]; else rows of ED then VAE -> [2N,2]."""
    if net == "ed":
        return ed_forward(sd_ed, x)
    if net == "vae":
        return vae_forward(sd_vae, x, eps, resize=False)[0]
    return torch.cat((ed_forward(sd_ed, x), vae_forward(sd_vae, x, eps, resize=False)[0]), dim=0)


def max_prediction_value(y_pred):
    """model/pred_func.py:123-131: mean over rows, argmax, and the reported score
    (m[0] if m[0] > m[1] else |1 - m[1]|; ties take the else branch)."""
    m = y_pred.mean(dim=0)
    cls = int(torch.argmax(m))
    val = float(m[0]) if bool(m[0] > m[1]) else float(abs(1 - m[1]))
    return cls, val


def pred_vid(logits):
    """model/pred_func.py:111-120 after the forward: sigmoid of the [rows,2] logits, then max_prediction_value."""
    return max_prediction_value(torch.sigmoid(logits.reshape(-1, 2)))


def embedder_forward(sd, x, prefix="embedder."):
    """``model.embedder(x)``: standalone Swin-T classifier (never reached by the reference forward)."""
    return swin_forward(sd, prefix, x)
