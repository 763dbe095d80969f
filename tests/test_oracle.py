"""CPU: the oracle against its pins -- torchvision's independent ConvNeXt-T / Swin-T
(bit-exact under an explicit key map) and the golden vectors produced by running the
unmodified reference in the authoring container (oracle/make_golden.py)."""
import torch
import torchvision

from oracle import backbones as B
from oracle import nets
from oracle import weights as W


def _convnext_to_torchvision(sd, cfg=B.CONVNEXT_TINY):
    m = {}
    for wb in ("weight", "bias"):
        m[f"features.0.0.{wb}"] = sd[f"stem.0.{wb}"]
        m[f"features.0.1.{wb}"] = sd[f"stem.1.{wb}"]
        m[f"classifier.0.{wb}"] = sd[f"head.norm.{wb}"]
        m[f"classifier.2.{wb}"] = sd[f"head.fc.{wb}"]
    for s, d in enumerate(cfg["depths"]):
        if s > 0:
            for j in (0, 1):
                for wb in ("weight", "bias"):
                    m[f"features.{2 * s}.{j}.{wb}"] = sd[f"stages.{s}.downsample.{j}.{wb}"]
        for k in range(d):
            p, q = f"stages.{s}.blocks.{k}.", f"features.{1 + 2 * s}.{k}."
            m[q + "layer_scale"] = sd[p + "gamma"].reshape(-1, 1, 1)
            for a, b in (("conv_dw", "block.0"), ("norm", "block.2"), ("mlp.fc1", "block.3"), ("mlp.fc2", "block.5")):
                for wb in ("weight", "bias"):
                    m[f"{q}{b}.{wb}"] = sd[f"{p}{a}.{wb}"]
    return m


def _swin_to_torchvision(sd, cfg=B.SWIN_TINY):
    m = {}
    for wb in ("weight", "bias"):
        m[f"features.0.0.{wb}"] = sd[f"patch_embed.proj.{wb}"]
        m[f"features.0.2.{wb}"] = sd[f"patch_embed.norm.{wb}"]
        m[f"norm.{wb}"] = sd[f"norm.{wb}"]
        m[f"head.{wb}"] = sd[f"head.{wb}"]
    for l, d in enumerate(cfg["depths"]):
        for k in range(d):
            p, q = f"layers.{l}.blocks.{k}.", f"features.{1 + 2 * l}.{k}."
            for a, b in (("norm1", "norm1"), ("norm2", "norm2"), ("attn.qkv", "attn.qkv"), ("attn.proj", "attn.proj"),
                         ("mlp.fc1", "mlp.0"), ("mlp.fc2", "mlp.3")):
                for wb in ("weight", "bias"):
                    m[f"{q}{b}.{wb}"] = sd[f"{p}{a}.{wb}"]
            m[q + "attn.relative_position_bias_table"] = sd[p + "attn.relative_position_bias_table"]
            m[q + "attn.relative_position_index"] = sd[p + "attn.relative_position_index"].flatten()
        if l < 3:
            p, q = f"layers.{l}.downsample.", f"features.{2 + 2 * l}."
            m[q + "reduction.weight"] = sd[p + "reduction.weight"]
            for wb in ("weight", "bias"):
                m[f"{q}norm.{wb}"] = sd[f"{p}norm.{wb}"]
    return m


def test_convnext_matches_torchvision_bit_exact():
    sd = {n: W.make_tensor(n, s, k, 3) for n, s, k in W._convnext_spec("")}
    tv = torchvision.models.convnext_tiny(weights=None).eval()
    tv.load_state_dict(_convnext_to_torchvision(sd), strict=True)
    for size in (224, 112):
        x = W.synthetic_frames(2, 5, size)
        with torch.no_grad():
            assert torch.equal(B.convnext_forward(sd, "", x), tv(x))


def test_swin_matches_torchvision_bit_exact():
    sd = {n: W.make_tensor(n, s, k, 3) for n, s, k in W._swin_spec("")}
    tv = torchvision.models.swin_t(weights=None).eval()
    tv.load_state_dict(_swin_to_torchvision(sd), strict=True)
    x = W.synthetic_frames(2, 6)
    with torch.no_grad():
        assert torch.equal(B.swin_forward(sd, "", x), tv(x))


def test_large_backbones_match_torchvision_bit_exact():
    """The '--s large' variants (reference prediction.py:314-318): convnext_large against torchvision's own
    convnext_large, swin_large_patch4_window7_224 against torchvision's SwinTransformer built with the large
    hyper-parameters (embed 192, depths 2/2/18/2, heads 6/12/24/48; torchvision ships no swin_l constructor)."""
    sd = {n: W.make_tensor(n, s, k, 3) for n, s, k in W._convnext_spec("", B.CONVNEXT_LARGE)}
    tv = torchvision.models.convnext_large(weights=None).eval()
    tv.load_state_dict(_convnext_to_torchvision(sd, B.CONVNEXT_LARGE), strict=True)
    x = W.synthetic_frames(1, 5)
    with torch.no_grad():
        assert torch.equal(B.convnext_forward(sd, "", x), tv(x))
    del tv, sd
    from torchvision.models.swin_transformer import SwinTransformer
    sd = {n: W.make_tensor(n, s, k, 3) for n, s, k in W._swin_spec("", B.SWIN_LARGE)}
    tv = SwinTransformer(patch_size=[4, 4], embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48],
                         window_size=[7, 7], stochastic_depth_prob=0.0).eval()
    tv.load_state_dict(_swin_to_torchvision(sd, B.SWIN_LARGE), strict=True)
    with torch.no_grad():
        assert torch.equal(B.swin_forward(sd, "", x), tv(x))


def test_state_dict_inventory_counts(golden):
    # counts observed when the REAL reference modules were instantiated (make_golden.py)
    assert len(W.ed_spec()) == golden("ed")["n_entries"] == 588
    assert len(W.vae_spec()) == golden("vae")["n_entries"] == 614


def test_ed_oracle_reproduces_reference_golden(sd_ed, golden):
    g = golden("ed")
    assert g["oracle_vs_reference"] == 0.0
    x = W.synthetic_frames(g["n"], g["meta"]["seed"])
    taps = {}
    with torch.no_grad():
        logits = nets.ed_forward(sd_ed, x, taps)
    # same torch build + same seeds: bit-exact here; other hosts may pick other oneDNN kernels
    assert torch.allclose(logits, g["logits"], atol=2e-5, rtol=0)
    assert torch.allclose(taps["decoded"][:, :, ::37, ::41], g["decoded_sample"], atol=2e-5)
    assert torch.allclose(taps["x1"][:, ::50], g["x1_sample"], atol=2e-5)
    assert torch.allclose(taps["x2"][:, ::50], g["x2_sample"], atol=2e-5)
    with torch.no_grad():
        emb = nets.embedder_forward(sd_ed, x[:1])
    assert torch.allclose(emb[:, ::50], g["embedder_logits_sample"], atol=2e-5)


def test_vae_oracle_reproduces_reference_golden(sd_vae, golden):
    g = golden("vae")
    assert g["oracle_vs_reference"] == 0.0
    x = W.synthetic_frames(g["n"], g["frames_seed"])
    eps = W.synthetic_eps(g["n"], g["eps_seed"])
    taps = {}
    with torch.no_grad():
        logits, xhat = nets.vae_forward(sd_vae, x, eps, taps)
        kl = nets.vae_kl(sd_vae, taps["feat"], taps["mu"])
    assert torch.allclose(logits, g["logits"], atol=2e-5, rtol=0)
    assert xhat.shape == (g["n"], 3, 224, 224) and taps["x_hat"].shape == (g["n"], 3, 112, 112)
    assert torch.allclose(xhat[:, :, ::37, ::41], g["xhat224_sample"], atol=2e-5)
    assert torch.allclose(taps["mu"][:, ::1001], g["mu_sample"], atol=2e-5)
    assert torch.allclose(taps["z"][:, ::1001], g["z_sample"], atol=2e-5)
    assert torch.allclose(kl, g["kl"], rtol=1e-5)


def test_full_model_and_pred_vid_golden(sd_ed, sd_vae, golden):
    g = golden("genconvit")
    x = W.synthetic_frames(g["n"], g["frames_seed"])
    eps = W.synthetic_eps(g["n"], g["eps_seed"])
    with torch.no_grad():
        rows = nets.genconvit_forward(sd_ed, sd_vae, x, eps)
    assert rows.shape == (2 * g["n"], 2)          # ED rows first, then VAE rows
    assert torch.allclose(rows, g["rows"], atol=2e-5, rtol=0)
    cls, val = nets.pred_vid(rows)
    assert cls == g["pred_vid"][0] and abs(val - g["pred_vid"][1]) < 1e-5
    assert {0: "REAL", 1: "FAKE"}[cls ^ 1] == g["real_or_fake"]


def test_max_prediction_value_edge_cases():
    # ties take the else branch: |1 - m1| (reference pred_func.py:128-130)
    assert nets.max_prediction_value(torch.tensor([[0.5, 0.5]])) == (0, 0.5)
    cls, val = nets.max_prediction_value(torch.tensor([[0.9, 0.2], [0.7, 0.4]]))
    assert cls == 0 and abs(val - 0.8) < 1e-6
    cls, val = nets.max_prediction_value(torch.tensor([[0.1, 0.8]]))
    assert cls == 1 and abs(val - 0.2) < 1e-6
