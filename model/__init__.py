"""Drop-in ``model`` package: the reference's import paths and class names
(``model.genconvit.GenConViT``, ``model.genconvit_ed.GenConViTED``,
``model.genconvit_vae.GenConViTVAE``, ``model.model_embedder.HybridEmbed``,
``model.config.load_config``, ``model.pred_func``), backed by the sm_100a kernel
library in ``genconvit_b200``.  See INTEGRATION.md.
"""
