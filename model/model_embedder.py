"""``HybridEmbed`` -- reference model/model_embedder.py:4-44.

In the reference this wraps the Swin classifier, probes it with a zero image at
construction (so ``feature_dim`` = 1000 and ``proj`` = Conv2d(1000, 768, 1)), is attached to
the ConvNeXt as ``patch_embed`` -- and is then never executed, because timm's ConvNeXt
forward does not read ``patch_embed`` (its own forward would raise on the 2-D [B,1000]
input).  It therefore only contributes parameters to the ``state_dict``:
``<backbone>.patch_embed.proj.{weight,bias}`` and a second view of the Swin tensors under
``<backbone>.patch_embed.backbone.*``.  No probe forward is run here.
"""
import torch.nn as nn


class HybridEmbed(nn.Module):
    def __init__(self, backbone, img_size=224, patch_size=1, feature_size=None, in_chans=3, embed_dim=768):
        super().__init__()
        assert isinstance(backbone, nn.Module)
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.backbone = backbone
        feature_dim = 1000                  # the Swin classifier's output width (reference :22-25)
        self.grid_size = (1, feature_dim)   # what the reference derives from o.shape[-2:] of a [1,1000] output
        self.num_patches = feature_dim
        self.proj = nn.Conv2d(feature_dim, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)

    def forward(self, x):
        raise RuntimeError("HybridEmbed.forward is unreachable in GenConViT (the reference's own forward raises "
                           "'Expected 3D or 4D input to conv2d' here); it exists for state_dict compatibility")
