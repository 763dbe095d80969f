"""Oracle (test infrastructure): timm==0.6.5 ConvNeXt-T / Swin-T arithmetic.

The reference reaches these networks through ``timm.create_model``
(reference model/genconvit_ed.py:68-69, model/genconvit_vae.py:96-97,
model/config.yaml:2-3; timm pinned at requirements.txt:5).  timm is not
vendored in the reference and not installable offline, so this file restates
the published architectures functionally, addressed by timm-0.6.5 state_dict
key names (``stem.0.weight`` ... / ``layers.0.blocks.0.attn.qkv.weight`` ...).
Pinned bit-exactly against torchvision's independent implementations in
tests/test_oracle.py.

All functions take ``sd`` (a mapping name -> fp32 CPU tensor) and a key prefix.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CONVNEXT_TINY = dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768))
SWIN_TINY = dict(depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), embed=96, window=7,
                 patch=4, img=224)
# the "--s large" variants of reference prediction.py:314-318 (timm convnext_large / swin_large_patch4_window7_224)
CONVNEXT_LARGE = dict(depths=(3, 3, 27, 3), dims=(192, 384, 768, 1536))
SWIN_LARGE = dict(depths=(2, 2, 18, 2), heads=(6, 12, 24, 48), embed=192, window=7,
                  patch=4, img=224)
CONVNEXT = {"convnext_tiny": CONVNEXT_TINY, "convnext_large": CONVNEXT_LARGE}
SWIN = {"swin_tiny_patch4_window7_224": SWIN_TINY, "swin_large_patch4_window7_224": SWIN_LARGE}


def convnext_cfg_of(sd, p=""):
    """Variant (tiny / large) of the ConvNeXt stored under prefix ``p``, read off the stem width."""
    c0 = sd[p + "stem.0.weight"].shape[0]
    for cfg in CONVNEXT.values():
        if cfg["dims"][0] == c0:
            return cfg
    raise ValueError(f"no ConvNeXt variant with stem width {c0}")


def swin_cfg_of(sd, p=""):
    e = sd[p + "patch_embed.proj.weight"].shape[0]
    for cfg in SWIN.values():
        if cfg["embed"] == e:
            return cfg
    raise ValueError(f"no Swin variant with embedding width {e}")


# --------------------------------------------------------------------------
# ConvNeXt (timm 0.6.5 models/convnext.py: stem -> stages -> norm_pre -> head)
# --------------------------------------------------------------------------
def _ln2d(x, w, b, eps):
    """timm LayerNorm2d: LayerNorm over C of an NCHW tensor."""
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def convnext_block(sd, p, x):
    """x + gamma * fc2(GELU(fc1(LN(dwconv7x7(x)))))  (ConvNeXtBlock, MLP in NHWC)."""
    c = x.shape[1]
    y = F.conv2d(x, sd[p + "conv_dw.weight"], sd[p + "conv_dw.bias"], padding=3, groups=c)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    y = y.permute(0, 3, 1, 2)
    return x + y * sd[p + "gamma"].reshape(1, -1, 1, 1)


def convnext_features(sd, p, x, taps=None):
    """stem + 4 stages.  x: [N,3,H,W] -> [N,768,H/32,W/32]."""
    x = F.conv2d(x, sd[p + "stem.0.weight"], sd[p + "stem.0.bias"], stride=4)
    x = _ln2d(x, sd[p + "stem.1.weight"], sd[p + "stem.1.bias"], 1e-6)
    if taps is not None:
        taps["stem"] = x
    for s, depth in enumerate(convnext_cfg_of(sd, p)["depths"]):
        q = f"{p}stages.{s}."
        if s > 0:
            x = _ln2d(x, sd[q + "downsample.0.weight"], sd[q + "downsample.0.bias"], 1e-6)
            x = F.conv2d(x, sd[q + "downsample.1.weight"], sd[q + "downsample.1.bias"], stride=2)
        for k in range(depth):
            x = convnext_block(sd, f"{q}blocks.{k}.", x)
        if taps is not None:
            taps[f"stage{s}"] = x
    return x


def convnext_forward(sd, p, x, taps=None):
    """Full classifier: features -> global avg pool -> LayerNorm2d -> fc.  -> [N,1000]."""
    x = convnext_features(sd, p, x, taps)
    x = x.mean((2, 3), keepdim=True)
    x = _ln2d(x, sd[p + "head.norm.weight"], sd[p + "head.norm.bias"], 1e-6)
    x = x.flatten(1)
    return F.linear(x, sd[p + "head.fc.weight"], sd[p + "head.fc.bias"])


# --------------------------------------------------------------------------
# Swin (timm 0.6.5 models/swin_transformer.py)
# --------------------------------------------------------------------------
def swin_relative_position_index(ws: int) -> torch.Tensor:
    """[ws*ws, ws*ws] int64 index into the (2ws-1)^2 bias table."""
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij"))
    flat = coords.flatten(1)
    rel = (flat[:, :, None] - flat[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def swin_attn_mask(res: int, ws: int, shift: int) -> torch.Tensor:
    """[nW, ws*ws, ws*ws] additive mask in {0,-100} for the cyclically shifted windows."""
    img = torch.zeros(1, res, res, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    win = _window_partition(img, ws).view(-1, ws * ws)
    m = win.unsqueeze(1) - win.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def _window_partition(x, ws):
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def _window_reverse(win, ws, h, w):
    b = win.shape[0] // ((h // ws) * (w // ws))
    x = win.view(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(b, h, w, -1)


def swin_window_attention(sd, p, xw, heads, mask):
    """xw: [B*nW, 49, C] -> same shape.  q scaled by head_dim^-0.5, rel-pos bias, optional mask."""
    bw, n, c = xw.shape
    hd = c // heads
    qkv = F.linear(xw, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(bw, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    table = sd[p + "relative_position_bias_table"]
    index = sd[p + "relative_position_index"]
    bias = table[index.view(-1)].view(n, n, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(bw // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = attn.softmax(-1)
    out = (attn @ v).transpose(1, 2).reshape(bw, n, c)
    return F.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def swin_block(sd, p, x, res, heads, ws, shift):
    b, l, c = x.shape
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(b, res, res, c)
    if shift > 0:
        h = torch.roll(h, (-shift, -shift), (1, 2))
    hw = _window_partition(h, ws).view(-1, ws * ws, c)
    mask = sd.get(p + "attn_mask") if shift > 0 else None
    hw = swin_window_attention(sd, p + "attn.", hw, heads, mask).view(-1, ws, ws, c)
    h = _window_reverse(hw, ws, res, res)
    if shift > 0:
        h = torch.roll(h, (shift, shift), (1, 2))
    x = x + h.view(b, l, c)
    y = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + y


def swin_patch_merging(sd, p, x, res):
    b, l, c = x.shape
    x = x.view(b, res, res, c)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.view(b, -1, 4 * c)
    x = F.layer_norm(x, (4 * c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def swin_forward(sd, p, x):
    """swin_{tiny,large}_patch4_window7_224 classifier: [N,3,224,224] -> [N,1000]."""
    cfg = swin_cfg_of(sd, p)
    ws = cfg["window"]
    x = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=cfg["patch"])
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (cfg["embed"],), sd[p + "patch_embed.norm.weight"], sd[p + "patch_embed.norm.bias"], 1e-5)
    res = cfg["img"] // cfg["patch"]
    for l, depth in enumerate(cfg["depths"]):
        for k in range(depth):
            shift = 0 if (k % 2 == 0 or res <= ws) else ws // 2
            x = swin_block(sd, f"{p}layers.{l}.blocks.{k}.", x, res, cfg["heads"][l], ws, shift)
        if l < len(cfg["depths"]) - 1:
            x = swin_patch_merging(sd, f"{p}layers.{l}.downsample.", x, res)
            res //= 2
    c = x.shape[-1]
    x = F.layer_norm(x, (c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    x = x.mean(1)
    return F.linear(x, sd[p + "head.weight"], sd[p + "head.bias"])
