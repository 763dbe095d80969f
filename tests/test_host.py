"""CPU: host-side logic of the drop-in package -- state_dict layout against the reference inventory,
config, pred_func utilities, video sharding and the world-size-2 score gather (gloo)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _config():
    from model.config import load_config
    return load_config()


def test_config_keys_match_reference_yaml():
    cfg = _config()
    assert cfg["model"] == {"backbone": "convnext_tiny", "embedder": "swin_tiny_patch4_window7_224", "latent_dims": 12544}
    for k, v in (("batch_size", 32), ("num_classes", 2), ("img_size", 224), ("learning_rate", 0.0001),
                 ("weight_decay", 0.0001), ("epoch", 1), ("min_val_loss", 10000)):
        assert cfg[k] == v
    cwd = os.getcwd()
    os.chdir("/")                       # the reference's cwd-relative load would fail here
    try:
        assert _config()["img_size"] == 224
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("net", ["ed", "vae"])
def test_state_dict_layout_is_the_reference_layout(net):
    from model.genconvit_ed import GenConViTED
    from model.genconvit_vae import GenConViTVAE
    from oracle import weights as W
    m = (GenConViTED if net == "ed" else GenConViTVAE)(_config())
    spec = W.ed_spec() if net == "ed" else W.vae_spec()
    sd = m.state_dict()
    assert set(sd) == {n for n, _, _ in spec} and len(sd) == (588 if net == "ed" else 614)
    for n, shape, _ in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
    # Swin tensors appear twice (embedder.* and <backbone>.patch_embed.backbone.*) and share storage
    bb = "backbone" if net == "ed" else "convnext_backbone"
    assert sd["embedder.head.weight"].data_ptr() == sd[f"{bb}.patch_embed.backbone.head.weight"].data_ptr()
    assert sd["embedder.layers.0.blocks.1.attn_mask"].shape == (64, 49, 49)
    assert sd["embedder.layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64


def test_large_variant_state_dict_layout():
    """--s large (reference prediction.py:314-318 rewrites config['model'] to convnext_large / swin_large_...)."""
    from model.genconvit_ed import GenConViTED
    from oracle import weights as W
    cfg = _config()
    cfg["model"] = dict(cfg["model"], backbone="convnext_large", embedder="swin_large_patch4_window7_224", type="large")
    with torch.device("meta"):
        m = GenConViTED(cfg)
    spec = W.ed_spec("large")
    sd = m.state_dict()
    assert set(sd) == {n for n, _, _ in spec}
    for n, shape, _ in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
    assert sd["backbone.stages.2.blocks.26.mlp.fc1.weight"].shape == (3072, 768)
    assert sd["embedder.layers.2.blocks.17.attn.qkv.weight"].shape == (2304, 768)


def test_strict_load_of_reference_shaped_checkpoint_and_dtype_switch(sd_ed):
    from model.genconvit_ed import GenConViTED
    from genconvit_b200.modules import compute_dtype_of
    m = GenConViTED(_config())
    assert m.load_state_dict(sd_ed, strict=True).missing_keys == []
    assert compute_dtype_of(m) == torch.float32
    assert compute_dtype_of(m.half()) == torch.float16
    assert compute_dtype_of(m, "bf16") == torch.bfloat16
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd_ed.items() if k != "fc2.bias"}, strict=True)


def test_unsupported_backbones_and_missing_weights_raise():
    from genconvit_b200.modules import create_model
    from model.genconvit import GenConViT
    with pytest.raises(NotImplementedError):
        create_model("convnext_base")             # the reference only ever selects tiny / large (prediction.py:314-318)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        with pytest.raises(Exception, match="weight/nope.pth file not found"):
            GenConViT(_config(), ed="nope", vae="nope", net="ed", fp16=False)
    finally:
        os.chdir(cwd)


def test_forward_on_cpu_fails_loudly_no_fallback(sd_ed):
    from genconvit_b200.lib import GcvError
    from model.genconvit_ed import GenConViTED
    m = GenConViTED(_config()).eval()
    with pytest.raises(GcvError):
        m(torch.zeros(1, 3, 224, 224))


def test_pred_func_host_helpers():
    import numpy as np
    from model import pred_func
    assert pred_func.real_or_fake(0) == "FAKE" and pred_func.real_or_fake(1) == "REAL"
    assert pred_func.max_prediction_value(torch.tensor([[0.5, 0.5]])) == (0, 0.5)
    frames = np.full((2, 224, 224, 3), 128, dtype=np.uint8)
    t = pred_func.preprocess_frame(frames).cpu()
    assert t.shape == (2, 3, 224, 224)
    want = (128 / 255.0 - 0.485) / 0.229
    assert abs(float(t[0, 0, 0, 0]) - want) < 1e-6
    r = pred_func.store_result(pred_func.set_result(), "a.mp4", 1, 0.9, "FAKE", "FAKE")
    assert r["video"]["pred_label"] == ["REAL"] and r["video"]["klass"] == ["fake"]
    assert not pred_func.is_video("/nonexistent.mp4")


def test_shard_videos_partitions_exactly():
    from genconvit_b200.runtime import shard_videos
    for n, world in ((17, 8), (1000, 8), (5, 2), (3, 4)):
        spans = [shard_videos(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from genconvit_b200.runtime import gather_scores, shard_videos
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n_videos = 6
    lo, hi = shard_videos(n_videos, rank, world)
    local = torch.stack((torch.arange(lo, hi).float() % 2, torch.arange(lo, hi).float() / 10))   # [2, V_local]
    out = gather_scores(local)
    q.put((rank, out))
    dist.destroy_process_group()


def test_gather_scores_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        out = results[r]
        assert out.shape == (2, 2, 3)
        flat = out.permute(1, 0, 2).reshape(2, 6)          # ranks are contiguous video shards
        assert torch.equal(flat[0], torch.arange(6).float() % 2)
        assert torch.allclose(flat[1], torch.arange(6).float() / 10)


def test_audit_checkpoint_tool_reports_renamed_and_misshapen_keys(sd_ed, tmp_path, capsys):
    """tools/audit_checkpoint.py: first-contact check for real weights (timm-owned key names are restated, not verified)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("audit_checkpoint", os.path.join(ROOT, "tools", "audit_checkpoint.py"))
    audit = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(audit)
    torch.save(sd_ed, tmp_path / "good.pth")
    assert audit.main([str(tmp_path / "good.pth"), "--net", "ed"]) == 0
    bad = {k.replace("mlp.fc1", "mlp.linear1"): v for k, v in sd_ed.items()}
    bad["fc.weight"] = torch.zeros(500, 1000)
    torch.save({"state_dict": bad, "epoch": 3}, tmp_path / "bad.pth")
    capsys.readouterr()
    assert audit.main([str(tmp_path / "bad.pth"), "--net", "ed", "--max-list", "2"]) == 1
    out = capsys.readouterr().out
    assert "wrapped {state_dict: ...}" in out and "~ backbone.stages.0.blocks.0.mlp.linear1.weight" in out
    assert "fc.weight: file (500, 1000), expected (500, 2000)" in out and "would FAIL" in out


def test_u8_frames_wrapper_validates_its_input():
    """engine.U8Frames: the raw-crop input form of the 16-bit forward (uint8 NHWC, contiguous); stands for NCHW frames."""
    from genconvit_b200.engine import U8Frames
    f = U8Frames(torch.zeros(2, 224, 224, 3, dtype=torch.uint8))
    assert f.shape == (2, 3, 224, 224) and f.mean == (0.485, 0.456, 0.406) and not f.is_cuda
    for bad in (torch.zeros(2, 224, 224, 3), torch.zeros(2, 3, 224, 224, dtype=torch.uint8),
                torch.zeros(2, 224, 224, 6, dtype=torch.uint8)[..., ::2]):
        with pytest.raises(ValueError):
            U8Frames(bad)
